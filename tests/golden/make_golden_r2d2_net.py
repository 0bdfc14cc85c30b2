"""Generates tests/golden/r2d2_net_golden.npz.  Run ONLY in the build container:
    python tests/golden/make_golden_r2d2_net.py

Executes the UNMODIFIED reference atari/networks.py (stack_frames, _unroll_cell,
DuelingLSTMDQNNet, imported whole) and common/utils.batch_apply over a Keras-layer shim whose
layers take oracle/r2d2_net_oracle.py's weights and compute with torch-CPU fp32.  Pins the
WIRING of the R2D2 network (frame stacking -> /255 -> conv body -> [conv_out, reward, one-hot]
-> LSTM with reset-before-step -> dueling head with mean-free advantages -> argmax)."""
import collections
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, HERE); sys.path.insert(0, ROOT)
from make_golden import _extract_function, _load  # noqa: E402
import make_golden_net as M  # noqa: E402
from oracle import net_oracle, r2d2_net_oracle as N  # noqa: E402

EnvOutput = collections.namedtuple('EnvOutput', 'reward done observation abandoned episode_step')
T, raw = M.T, M.raw
A, OBS, S, T1, B = 5, [36, 36, 1], 4, 6, 2


def make_params():
  p = net_oracle.to_torch(N.init_params(A, OBS, S, seed=2))
  rng = np.random.default_rng(5)
  for k in p:
    if k.endswith('bias'):
      p[k] = p[k] + torch.as_tensor(rng.normal(size=tuple(p[k].shape)).astype(np.float32)) * 0.1
  return p


def make_inputs():
  rng = np.random.default_rng(6)
  return dict(obs=rng.integers(0, 256, [T1, B] + OBS, dtype=np.uint8), rew=rng.normal(size=(T1, B)).astype(np.float32),
              done=rng.random((T1, B)) < 0.3, prev=rng.integers(0, A, (T1, B)).astype(np.int32),
              h0=rng.normal(size=(B, 512)).astype(np.float32), c0=rng.normal(size=(B, 512)).astype(np.float32),
              fstate=rng.integers(0, 1 << 24, (B, 36 * 36)).astype(np.int32))


def build_tf(p):
  tf, order = M.build_tf(p, lambda logits: logits.argmax(-1))
  tf.int32, tf.bool, tf.uint8 = torch.int32, torch.bool, torch.uint8
  act = lambda a, x: torch.relu(x) if a == 'relu' else x

  class Sequential(object):
    def __init__(self, layers): self.layers = layers
    def __call__(self, x):
      for l in self.layers:
        x = l(x)
      return x

  names = {'conv': ['body/conv0', 'body/conv1', 'body/conv2'],
           'dense': ['body/dense', 'value/hidden', 'value/head', 'advantage/hidden', 'advantage/head']}
  cnt = {'conv': 0, 'dense': 0}

  class Conv2D(object):
    def __init__(self, ch, k, strides, padding='valid', activation=None):
      self.n = names['conv'][cnt['conv']]; cnt['conv'] += 1
      self.s, self.same, self.act = strides, padding == 'same', activation
      assert tuple(p[self.n + '/kernel'].shape[:2]) == tuple(k) and p[self.n + '/kernel'].shape[3] == ch
    def __call__(self, x):
      return T(act(self.act, net_oracle._conv_nhwc(raw(x), p[self.n + '/kernel'], p[self.n + '/bias'], self.s, self.same)))

  class Dense(object):
    def __init__(self, units, activation=None, use_bias=True, name=None):
      self.n = names['dense'][cnt['dense']]; cnt['dense'] += 1
      self.act, self.use_bias = activation, use_bias
      assert p[self.n + '/kernel'].shape[1] == units and (use_bias == ((self.n + '/bias') in p))
    def __call__(self, x):
      y = raw(x) @ p[self.n + '/kernel']
      if self.use_bias:
        y = y + p[self.n + '/bias']
      return T(act(self.act, y))

  class LSTMCell(object):
    def __init__(self, units): self.units = units
    def get_initial_state(self, batch_size, dtype):
      return [T(torch.zeros(int(batch_size), self.units)), T(torch.zeros(int(batch_size), self.units))]
    def __call__(self, x, state):
      h, c = N.lstm_cell(p, raw(x), raw(state[0]), raw(state[1]))
      return T(h), [T(h), T(c)]

  tf.keras.Sequential = Sequential
  tf.keras.layers.Conv2D, tf.keras.layers.Dense, tf.keras.layers.LSTMCell = Conv2D, Dense, LSTMCell
  # ops of stack_frames / _unroll_cell / _head beyond make_golden_net's set
  tf.zeros = lambda shape, dtype=torch.float32: T(torch.zeros([int(raw(v)) for v in (raw(shape).tolist() if isinstance(raw(shape), torch.Tensor) else shape)], dtype=dtype))
  tf.zeros_like = lambda x, dtype=None: T(torch.zeros_like(raw(x), dtype=dtype))
  tf.math = types.ModuleType('math')
  tf.math.reduce_prod = lambda x: int(np.prod(x))
  tf.math.logical_or = lambda a, b: T(raw(a) | raw(b))
  tf.bitwise = types.ModuleType('bitwise')
  tf.bitwise.right_shift = lambda x, n: T(raw(x) >> n)
  tf.bitwise.left_shift = lambda x, n: T(raw(x) << torch.as_tensor(n, dtype=torch.int32))
  tf.bitwise.bitwise_and = lambda x, m: T(raw(x) & m)
  tf.pad = lambda x, pads: T(torch.as_tensor(np.pad(raw(x).numpy(), pads)))
  tf.reduce_sum = lambda x, axis=None: T(raw(x).sum(dim=axis, dtype=raw(x).dtype))
  tf.reduce_mean = lambda x, axis=None, keepdims=False: T(raw(x).mean(dim=axis, keepdim=keepdims))
  tf.argmax = lambda x, axis: T(raw(x).argmax(dim=axis))
  tf.convert_to_tensor = lambda x, dtype=None: x if isinstance(x, T) else T(x)
  old_concat = tf.concat
  tf.concat = lambda xs, axis: (M.Sh(sum([list(x) for x in xs], [])) if not isinstance(xs[0], T) and not isinstance(xs[0], torch.Tensor)
                                else old_concat(xs, axis))
  # list + TensorShape arithmetic used by stack_frames ([batch_size] + obs_shape, shape[0:2] + [1] * n)
  M.Sh.__add__ = lambda self, o: M.Sh(list(self) + list(o))
  M.Sh.__radd__ = lambda self, o: M.Sh(list(o) + list(self))
  M.Sh.num_elements = lambda self: int(np.prod(self)) if len(self) else 1
  M.T.__sub__ = lambda self, o: T(raw(self) - raw(o))
  M.T.__isub__ = lambda self, o: T(raw(self) - raw(o))
  M.T.dtype = property(lambda self: self.a.dtype)
  return tf


def main():
  p = make_params()
  tf = build_tf(p)
  sys.modules['tensorflow'] = tf
  seed_rl = types.ModuleType('seed_rl'); common = types.ModuleType('seed_rl.common'); utils = types.ModuleType('seed_rl.common.utils')
  utils.batch_apply = _extract_function(os.path.join(REF, 'common/utils.py'), 'batch_apply', {'tf': tf})
  seed_rl.common = common; common.utils = utils
  sys.modules.update({'seed_rl': seed_rl, 'seed_rl.common': common, 'seed_rl.common.utils': utils})
  ref = _load(os.path.join(REF, 'atari/networks.py'), 'ref_atari_networks')
  agent = ref.DuelingLSTMDQNNet(A, OBS, stack_size=S)
  i = make_inputs()
  env = EnvOutput(T(i['rew']), T(i['done']), T(i['obs']), T(np.zeros((T1, B), bool)), T(np.zeros((T1, B), np.int32)))
  state = ref.AgentState([T(i['h0']), T(i['c0'])], T(i['fstate']))
  with torch.no_grad():
    out, st = agent((T(i['prev']), env), state, unroll=True)
  np.savez_compressed(os.path.join(HERE, 'r2d2_net_golden.npz'), q=raw(out.q_values).numpy(), action=raw(out.action).numpy(),
                      h=raw(st.core_state[0]).numpy(), c=raw(st.core_state[1]).numpy(),
                      fstate=raw(st.frame_stacking_state).numpy())
  print('wrote r2d2_net_golden.npz; q', tuple(raw(out.q_values).shape))


if __name__ == '__main__':
  main()
