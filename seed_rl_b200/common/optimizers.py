"""Optimizer on the hot path: tf.keras.optimizers.Adam semantics (TF 2.4.1) as built by
the reference at dmlab/vtrace_main.py:46-51 and applied at
agents/vtrace/learner.py:272-273 -- here ONE fused kernel over the flat arena
(seedrl_adam_apply).  `iterations` is the resumable step counter (learner.py:243)."""
import math

import numpy as np
import torch

from seed_rl_b200 import _lib


class PolynomialDecay(object):
  """tf.keras.optimizers.schedules.PolynomialDecay (cycle=False)."""

  def __init__(self, initial_learning_rate, decay_steps, end_learning_rate=0.0001, power=1.0):
    self.initial_learning_rate = initial_learning_rate
    self.decay_steps = decay_steps
    self.end_learning_rate = end_learning_rate
    self.power = power

  def __call__(self, step):
    step = min(float(step), float(self.decay_steps))
    return ((self.initial_learning_rate - self.end_learning_rate) *
            (1.0 - step / float(self.decay_steps)) ** self.power + self.end_learning_rate)


class Adam(object):
  """Keras Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps)."""

  def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
    self.learning_rate = learning_rate
    self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
    self.iterations = 0
    self.m = None
    self.v = None

  def _lr(self):
    lr = self.learning_rate
    return float(lr(self.iterations)) if callable(lr) else float(lr)

  def _create_slots(self, params):
    if self.m is None:
      self.m = torch.zeros_like(params)
      self.v = torch.zeros_like(params)

  def apply_gradients(self, params, grads, grad_scale=1.0, clamp_index=-1,
                      clamp_lo=0.0, clamp_hi=0.0):
    """params/grads: flat fp32 CUDA arenas (updated in place)."""
    self._create_slots(params)
    # Keras `_prepare_local` computes these in fp32 tensors
    f = np.float32
    t = f(self.iterations + 1)
    lr_t = float(f(self._lr()) * (np.sqrt(f(1) - np.power(f(self.beta_2), t, dtype=f), dtype=f) /
                                  (f(1) - np.power(f(self.beta_1), t, dtype=f))))
    _lib.check(_lib.lib().seedrl_adam_apply(
        params.numel(), _lib.ptr(params), _lib.ptr(grads), _lib.ptr(self.m), _lib.ptr(self.v),
        lr_t, self.beta_1, self.beta_2, self.epsilon, grad_scale, clamp_index, clamp_lo,
        clamp_hi, _lib.stream_ptr()))
    self.iterations += 1

  def state_dict(self):
    return {'iterations': self.iterations,
            'm': None if self.m is None else self.m.cpu(),
            'v': None if self.v is None else self.v.cpu()}

  def load_state_dict(self, d, device='cuda'):
    self.iterations = int(d['iterations'])
    self.m = None if d['m'] is None else d['m'].to(device)
    self.v = None if d['v'] is None else d['v'].to(device)
