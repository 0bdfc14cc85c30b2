"""Generates tests/golden/r2d2_golden.npz.  Run ONLY in the build container (where
/root/reference exists):   python tests/golden/make_golden_r2d2.py

Executes the UNMODIFIED reference functions of agents/r2d2/learner.py (pulled out of the file
by AST, because its module-level imports need TensorFlow / the seed_rl package) over
tf_numpy_shim's numpy-fp32 stand-ins; nothing is copied into this repo.  Inputs are stored
next to the outputs so the GPU box (no /root/reference) can replay them."""
import collections
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402
from make_golden import _extract_function  # noqa: E402


def extend(tf):
  T, raw = tf_numpy_shim.Tensor, tf_numpy_shim._raw
  f32 = np.float32
  tf.math.sign = lambda x: T(np.sign(raw(x)))
  tf.math.sqrt = lambda x: T(np.sqrt(raw(x)))
  tf.math.abs = lambda x: T(np.abs(raw(x)))
  tf.math.square = lambda x: T(np.square(raw(x)))
  tf.math.pow = lambda x, y: T(np.power(f32(raw(x)), raw(y)).astype(f32))
  tf.abs = tf.math.abs
  tf.shape = lambda x: np.asarray(raw(x)).shape
  tf.one_hot = lambda idx, depth, on, off: T(np.where(
      np.arange(depth) == np.asarray(raw(idx))[..., None], f32(on), f32(off)).astype(f32))
  tf.reduce_sum = lambda x, axis=None: T(np.sum(raw(x), axis=axis, dtype=f32))
  tf.reduce_max = lambda x, axis=None: T(np.max(raw(x), axis=axis))
  tf.reduce_mean = lambda x, axis=None: T(np.mean(raw(x), axis=axis, dtype=f32))
  tf.linspace = lambda a, b, num: T(np.linspace(a, b, num, dtype=f32))
  tf.constant = lambda v, dtype=None: T(np.asarray(v, f32 if dtype is None else dtype))
  tf.gather = lambda p, i: T(np.asarray(raw(p))[np.asarray(raw(i))])
  # ---- ops of atari/networks.py:57-173 (stack_frames) ----
  class _Sh(list):                      # TensorShape-like: list with rank / num_elements
    @property
    def rank(self): return len(self)
    def num_elements(self): return int(np.prod(self)) if len(self) else 1
    def __getitem__(self, i):
      r = list.__getitem__(self, i)
      return _Sh(r) if isinstance(i, slice) else r
    def __add__(self, o): return _Sh(list(self) + list(o))
    def __radd__(self, o): return _Sh(list(o) + list(self))
  class T2(T):
    @property
    def shape(self): return _Sh(self.a.shape)
    def __getitem__(self, idx): return T2(self.a[idx])
  tf._T2 = T2
  tf.reshape = lambda x, shape: T2(np.reshape(raw(x), [int(v) for v in shape]))
  tf.bitwise = types.ModuleType('bitwise')
  tf.bitwise.right_shift = lambda x, n: T2(np.right_shift(raw(x), n))
  tf.bitwise.left_shift = lambda x, n: T2(np.left_shift(raw(x), np.asarray(n, np.int32)))
  tf.bitwise.bitwise_and = lambda x, m: T2(np.bitwise_and(raw(x), m))
  tf.zeros = lambda shape, dtype=np.float32: T2(np.zeros([int(v) for v in shape], dtype))
  tf.math.logical_or = lambda a, b: T2(np.logical_or(raw(a), raw(b)))
  tf.pad = lambda x, pads: T2(np.pad(raw(x), pads))
  tf.where = lambda c, a, b: T2(np.where(raw(c), raw(a), raw(b)))
  _old_concat = tf.concat
  tf.concat = lambda values, axis, name=None: T2(np.concatenate([np.asarray(raw(v)) for v in values], axis=axis))
  tf.zeros_like = lambda x, dtype=None, name=None: T2(np.zeros_like(raw(x), dtype=dtype))
  _old_cast = tf.cast
  tf.cast = lambda x, dtype, name=None: T2(np.asarray(raw(x)).astype(dtype))
  tf.reduce_sum = lambda x, axis=None: T2(np.sum(raw(x), axis=axis, dtype=np.asarray(raw(x)).dtype))
  tf.reduce_max = lambda x, axis=None: T2(np.max(raw(x), axis=axis))
  tf.reduce_mean = lambda x, axis=None: T2(np.mean(raw(x), axis=axis, dtype=np.float32))
  return tf


def main():
  tf = extend(tf_numpy_shim.install())
  flags = types.SimpleNamespace(value_function_rescaling_epsilon=1e-3, n_steps=5)
  ns = {'tf': tf, 'FLAGS': flags, 'np': np}
  path = os.path.join(REF, 'agents/r2d2/learner.py')
  for fn in ('value_function_rescaling', 'inverse_value_function_rescaling', 'n_step_bellman_target',
             'compute_loss_and_priorities_from_agent_outputs', 'get_envs_epsilon'):
    _extract_function(path, fn, ns)
  T = tf_numpy_shim.Tensor
  out = {}
  rng = np.random.default_rng(7)
  f32 = np.float32

  x = np.concatenate([np.linspace(-100., 100., 50), [0., 3., -3., 1000., -1000., 1e-4]]).astype(f32)
  out['resc_x'] = x
  out['resc_h'] = ns['value_function_rescaling'](T(x)).a
  out['resc_hinv'] = ns['inverse_value_function_rescaling'](T(x)).a

  for name, (Tn, B, n, gamma) in {'a': (7, 1, 3, 0.9), 'b': (20, 5, 5, 0.997), 'c': (12, 3, 1, 0.99),
                                 'd': (4, 2, 5, 0.997)}.items():
    r = rng.normal(size=(Tn, B)).astype(f32); d = rng.random((Tn, B)) < 0.2
    q = (rng.normal(size=(Tn, B)) * 10).astype(f32)
    out['nstep_%s_in' % name] = np.stack([r, d.astype(f32), q])
    out['nstep_%s_cfg' % name] = np.asarray([n, gamma])
    out['nstep_%s_out' % name] = ns['n_step_bellman_target'](T(r), T(d), T(q), gamma, n).a

  AgentOutput = collections.namedtuple('AgentOutput', 'action q_values')
  EnvOutput = collections.namedtuple('EnvOutput', 'reward done')
  Tn, B, A = 16, 6, 18
  tq = rng.normal(size=(Tn, B, A)).astype(f32); gq = rng.normal(size=(Tn, B, A)).astype(f32)
  ta = tq.argmax(-1); ra = rng.integers(0, A, (Tn, B))
  r = rng.normal(size=(Tn, B)).astype(f32); d = rng.random((Tn, B)) < 0.1
  loss, prio = ns['compute_loss_and_priorities_from_agent_outputs'](
      AgentOutput(T(ta), T(tq)), AgentOutput(None, T(gq)), EnvOutput(T(r), T(d)), AgentOutput(T(ra), None),
      0.997)
  out.update(loss_train_q=tq, loss_target_q=gq, loss_train_action=ta, loss_replay_action=ra, loss_reward=r,
             loss_done=d, loss_out=loss.a, loss_priorities=prio.a)

  out['eps_out'] = ns['get_envs_epsilon'](T(np.arange(20)), 10, 10, 1e-3).a

  # ---- stack_frames (atari/networks.py:57-173): bit-packed frame stacking -----------------
  ns2 = {'tf': tf, 'STACKING_STATE_DTYPE': np.int32}
  stack = _extract_function(os.path.join(REF, 'atari/networks.py'), 'stack_frames', ns2)
  T2 = tf._T2
  for name, (Tn, B, H, W, S) in {'a': (6, 2, 3, 4, 4), 'b': (9, 3, 5, 2, 3), 'c': (4, 1, 2, 2, 2)}.items():
    fr = rng.integers(0, 256, (Tn, B, H, W, 1)).astype(f32)
    dn = rng.random((Tn, B)) < 0.3
    st = rng.integers(0, 1 << (8 * (S - 1)), (B, H * W)).astype(np.int32)
    stacked, new_state = stack(T2(fr), T2(st), T2(dn), S)
    out['stack_%s_frames' % name] = fr; out['stack_%s_done' % name] = dn; out['stack_%s_state' % name] = st
    out['stack_%s_size' % name] = np.asarray(S)
    out['stack_%s_out' % name] = stacked.a; out['stack_%s_new_state' % name] = new_state.a
  np.savez_compressed(os.path.join(HERE, 'r2d2_golden.npz'), **out)
  print('wrote r2d2_golden.npz:', sorted(out))


if __name__ == '__main__':
  main()
