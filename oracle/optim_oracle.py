"""CPU ORACLE (test infrastructure, NOT product code) -- numpy restatement of the
optimizer on the hot path.

The reference applies `tf.keras.optimizers.Adam` (TF 2.4.1, NOT vendored under
/root/reference; pinned by docker/Dockerfile.atari:14) built at
dmlab/vtrace_main.py:46-51 and applied at agents/vtrace/learner.py:272-273.
Its dense update rule (`_resource_apply_dense`, non-amsgrad), as published:

    t      = iterations + 1
    lr_t   = lr(iterations) * sqrt(1 - beta2^t) / (1 - beta1^t)
    m      = beta1*m + (1-beta1)*g
    v      = beta2*v + (1-beta2)*g^2
    theta -= lr_t * m / (sqrt(v) + eps)          # eps OUTSIDE the bias correction

No reference test exercises Adam on this path => PARITY UNPINNED;
the only pinned optimizer-side semantic is cross-replica SUM of gradients
(tests/utils_test.py:609-650).
"""
import numpy as np


def polynomial_decay(initial_lr, step, decay_steps, end_lr=0.0, power=1.0):
  """tf.keras.optimizers.schedules.PolynomialDecay (cycle=False), used at
  dmlab/vtrace_main.py:47-48."""
  step = min(float(step), float(decay_steps))
  return (initial_lr - end_lr) * (1.0 - step / float(decay_steps)) ** power + end_lr


def keras_adam_step(p, g, m, v, iterations, lr, beta1=0.9, beta2=0.999, eps=1e-7):
  """One dense Adam step in fp32.  `iterations` = optimizer.iterations BEFORE the
  step (0 for the first step).  Returns new (p, m, v)."""
  f = np.float32
  t = iterations + 1
  lr_t = f(lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t))
  p, g, m, v = (np.asarray(x, f) for x in (p, g, m, v))
  m2 = f(beta1) * m + f(1.0 - beta1) * g
  v2 = f(beta2) * v + f(1.0 - beta2) * g * g
  p2 = p - lr_t * m2 / (np.sqrt(v2) + f(eps))
  return p2.astype(f), m2.astype(f), v2.astype(f)
