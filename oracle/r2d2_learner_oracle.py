"""CPU ORACLE (test infrastructure, NOT product code) -- one whole R2D2 learner step on the CPU:

  compute_loss_and_priorities   /root/reference/agents/r2d2/learner.py:333-386 (burn-in prefix
                                unrolled by both networks without gradient, suffix by both,
                                n-step double-DQN loss on the suffix)
  minimize                      /root/reference/agents/r2d2/learner.py:581-634 (mean of the
                                importance-weighted loss, tape.gradient, global-norm clip with the
                                norm before clipping, Keras Adam)
composed from r2d2_net_oracle (the network, torch autograd standing in for tf.GradientTape),
r2d2_oracle (the pinned post-network arithmetic: the Bellman target is a stop_gradient constant,
so it is taken from the numpy restatement) and optim_oracle (Keras Adam).  Parity reference of
the CUDA cfg-5 step and, timed, bench.py's cpu_baseline for `--agent r2d2`.
"""
import collections

import numpy as np
import torch

from . import optim_oracle, r2d2_net_oracle as N, r2d2_oracle as R


def synthetic_replay_batch(T, B, A, obs_shape=(84, 84, 1), seed=1234, done_p=0.01):
  """A sampled replay batch, time-major with T = burn_in + unroll_length + 1 rows (SURVEY 8d
  conventions: uint8 frames U{0..255}, N(0,1) rewards, Bernoulli done)."""
  rng = np.random.default_rng(seed)
  return dict(
      observation=rng.integers(0, 256, (T, B) + tuple(obs_shape), dtype=np.uint8),
      reward=rng.normal(size=(T, B)).astype(np.float32),
      done=rng.random((T, B)) < done_p,
      prev_actions=rng.integers(0, A, (T, B)).astype(np.int32),
      action=rng.integers(0, A, (T, B)).astype(np.int32),
      h0=(0.1 * rng.normal(size=(B, N.LSTM_UNITS))).astype(np.float32),
      c0=(0.1 * rng.normal(size=(B, N.LSTM_UNITS))).astype(np.float32),
      frame_state=rng.integers(0, 1 << 24, (B, int(np.prod(obs_shape)))).astype(np.int32),
      importance_weights=(rng.random(B) * 0.9 + 0.1).astype(np.float32),
      indices=rng.integers(0, 100, B).astype(np.int64))


def _split(batch, burn_in):
  keys = ('observation', 'reward', 'done', 'prev_actions', 'action')
  pre = {k: batch[k][:burn_in] for k in keys}
  suf = {k: batch[k][burn_in:] for k in keys}
  return pre, suf


def _unroll(p, part, state, A, stack_size):
  return N.unroll(p, part['prev_actions'], part['reward'], part['done'], part['observation'], state, A, stack_size)


def compute_loss_and_priorities(p_train, p_target, batch, A, stack_size, gamma, burn_in, n_steps=5, eps=1e-3):
  """learner.py:333-386.  p_*: dicts of torch tensors.  Returns (loss [B] torch (differentiable
  w.r.t. p_train), priorities [B] numpy, aux)."""
  fs = batch['frame_state'] if stack_size > 1 else ()
  state = N.AgentState((torch.as_tensor(batch['h0']), torch.as_tensor(batch['c0'])), fs)
  if burn_in:
    pre, suf = _split(batch, burn_in)
    with torch.no_grad():                                                      # stop_gradient :369
      _, train_state = _unroll(p_train, pre, state, A, stack_size)
      _, target_state = _unroll(p_target, pre, state, A, stack_size)
  else:
    suf = batch
    train_state = target_state = state
  train_out, _ = _unroll(p_train, suf, train_state, A, stack_size)
  with torch.no_grad():
    target_out, _ = _unroll(p_target, suf, target_state, A, stack_size)
  q = train_out.q_values
  T, B = q.shape[0], q.shape[1]
  tq = q.detach().numpy()
  # the (stop_gradient) Bellman target and the priorities from the pinned numpy restatement
  loss_np, prio, abs_td = R.loss_and_priorities(tq, tq.argmax(-1), target_out.q_values.numpy(), suf['action'],
                                                suf['reward'], suf['done'], gamma, n_steps=n_steps, eps=eps)
  tt, bb = np.meshgrid(np.arange(T), np.arange(B), indexing='ij')
  qtarget_max = R.inverse_value_function_rescaling(target_out.q_values.numpy()[tt, bb, tq.argmax(-1)], eps)
  target = R.value_function_rescaling(R.n_step_bellman_target(suf['reward'], suf['done'], qtarget_max, gamma,
                                                              n_steps)[1:], eps)
  replay_q = torch.gather(q, 2, torch.as_tensor(np.asarray(suf['action'])).long()[..., None])[..., 0][:-1]
  td = torch.as_tensor(target) - replay_q
  loss = 0.5 * (td * td).sum(dim=0)                                            # :329
  return loss, prio, dict(q=q, target_q=target_out.q_values, loss_np=loss_np, abs_td=abs_td)


class CpuR2D2Learner(object):
  """Holds online + target params and Adam slots; step() = one `minimize` (:581-634)."""

  def __init__(self, num_actions, obs_shape, stack_size, gamma=0.997, burn_in=40, n_steps=5, clip_norm=40.0,
               lr=0.00048, beta1=0.9, beta2=0.999, eps=1e-3, params=None, target_params=None, seed=0):
    self.A, self.obs_shape, self.stack = num_actions, tuple(obs_shape), stack_size
    self.gamma, self.burn_in, self.n_steps, self.clip_norm = gamma, burn_in, n_steps, clip_norm
    init = params if params is not None else N.init_params(num_actions, obs_shape, stack_size, seed)
    self.params = collections.OrderedDict((k, torch.tensor(np.asarray(v, np.float32), requires_grad=True))
                                          for k, v in init.items())
    tinit = target_params if target_params is not None else init
    self.target = collections.OrderedDict((k, torch.tensor(np.asarray(v, np.float32))) for k, v in tinit.items())
    self.m = {k: np.zeros(tuple(v.shape), np.float32) for k, v in self.params.items()}
    self.v = {k: np.zeros(tuple(v.shape), np.float32) for k, v in self.params.items()}
    self.iterations = 0
    self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps

  def grads(self, batch):
    for t in self.params.values():
      t.grad = None
    loss, prio, aux = compute_loss_and_priorities(self.params, self.target, batch, self.A, self.stack, self.gamma,
                                                  self.burn_in, self.n_steps)
    total = (loss * torch.as_tensor(batch['importance_weights'])).mean()       # :604
    total.backward()
    g = collections.OrderedDict((k, v.grad.numpy().copy()) for k, v in self.params.items())
    norm = float(np.sqrt(sum(float((x.astype(np.float64) ** 2).sum()) for x in g.values())))   # :605
    return float(total.detach()), loss.detach().numpy(), prio, g, norm, aux

  def step(self, batch):
    total, loss, prio, g, norm, _ = self.grads(batch)
    scale = np.float32(self.clip_norm / max(norm, self.clip_norm)) if self.clip_norm else np.float32(1)   # :606-609
    with torch.no_grad():
      for k, p in self.params.items():
        p2, self.m[k], self.v[k] = optim_oracle.keras_adam_step(p.detach().numpy(), g[k] * scale, self.m[k],
                                                                self.v[k], self.iterations, self.lr, self.b1,
                                                                self.b2, self.eps)
        p.copy_(torch.from_numpy(np.asarray(p2)))
    self.iterations += 1
    return total, prio, norm
