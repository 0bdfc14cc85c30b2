"""Generates tests/golden/net_golden.npz.  Run ONLY in the build container (where
/root/reference exists):   python tests/golden/make_golden_net.py

Executes the UNMODIFIED reference network source dmlab/networks.py (classes _Stack and
ImpalaDeep, imported whole) and the unmodified `batch_apply` of common/utils.py over a
Keras-layer shim: tf.keras.layers.{Conv2D, MaxPool2D, Dense, Flatten, LSTMCell} are stand-ins
that take their weights from a given parameter dictionary and compute with torch-CPU fp32
(the layer arithmetic itself is what oracle/net_oracle.py restates and
tests/test_oracle_layers_independent.py cross-checks); every other tf.* op is a thin numpy /
torch call.  What this pins is the WIRING of the reference network -- stack order, conv ->
pool -> two residual blocks with their ReLU placement and skip adds, the ReLU before Flatten,
NHWC flatten order, Dense+ReLU, reward clip, one-hot, concat order, LSTM gate/state order,
reset-on-done BEFORE the step, heads -- to the reference source rather than our reading."""
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
from oracle import net_oracle  # noqa: E402

EnvOutput = collections.namedtuple('EnvOutput', 'reward done observation abandoned episode_step')


class Sh(list):                              # TensorShape stand-in
  @property
  def rank(self): return len(self)
  def as_list(self): return list(self)
  def __getitem__(self, i):
    r = list.__getitem__(self, i)
    return Sh(r) if isinstance(i, slice) else r


class T(object):                             # tensor stand-in over torch fp32 / integer tensors
  def __init__(self, a): self.a = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
  @property
  def shape(self): return Sh(self.a.shape)
  def __getitem__(self, i): return T(self.a[i])
  def __add__(self, o): return T(self.a + raw(o))
  def __iadd__(self, o): return T(self.a + raw(o))
  def __truediv__(self, o): return T(self.a / raw(o))
  def __itruediv__(self, o): return T(self.a / raw(o))


def raw(x): return x.a if isinstance(x, T) else x


def map_structure(fn, *structs):
  s0 = structs[0]
  if isinstance(s0, tuple) and hasattr(s0, '_fields'):
    return type(s0)(*[map_structure(fn, *[getattr(s, f) for s in structs]) for f in s0._fields])
  if isinstance(s0, (tuple, list)) and not isinstance(s0, Sh):
    return type(s0)(map_structure(fn, *xs) for xs in zip(*structs))
  return fn(*structs)


def flatten(s):
  if isinstance(s, (tuple, list)) and not isinstance(s, Sh):
    out = []
    for x in s:
      out += flatten(x)
    return out
  return [s]


def build_tf(params, sample_actions):
  """params: oracle-named weights.  Layers take their weights in creation order."""
  tf = types.ModuleType('tensorflow')
  tf.float32, tf.int64 = torch.float32, torch.int64
  tf.Module = type('Module', (object,), {'__init__': lambda self, name=None: None})
  tf.function = lambda f: f
  tf.nest = types.ModuleType('nest'); tf.nest.map_structure = map_structure; tf.nest.flatten = flatten
  tf.nn = types.ModuleType('nn'); tf.nn.relu = lambda x: T(torch.relu(raw(x)))
  tf.cast = lambda x, dt: T(raw(x).to(dt))
  tf.reshape = lambda x, shape: T(raw(x).reshape([int(v) for v in shape]))
  tf.expand_dims = lambda x, axis: T(raw(x).unsqueeze(axis))
  tf.squeeze = lambda x, axis=None, name=None: T(raw(x).squeeze(axis))
  tf.clip_by_value = lambda x, lo, hi: T(torch.clamp(raw(x), lo, hi))
  tf.one_hot = lambda idx, depth: T(torch.nn.functional.one_hot(raw(idx).long(), depth).to(torch.float32))
  tf.concat = lambda xs, axis: T(torch.cat([raw(x) for x in xs], dim=axis))
  tf.unstack = lambda x: [T(v) for v in raw(x)]
  tf.stack = lambda xs: T(torch.stack([raw(x) for x in xs]))
  tf.where = lambda c, a, b: T(torch.where(raw(c), raw(a), raw(b)))
  tf.shape = lambda x: list(raw(x).shape)
  tf.random = types.ModuleType('random')
  tf.random.categorical = lambda logits, n, dtype=None: T(sample_actions(raw(logits))[:, None])

  order = {'conv': [], 'dense': []}           # creation order -> oracle names (filled by the caller)
  counters = {'conv': 0, 'dense': 0}

  class Conv2D(object):
    def __init__(self, ch, k, strides=1, padding='same', name=None):
      self.name = order['conv'][counters['conv']]; counters['conv'] += 1
      self.s, self.same = strides, padding == 'same'
      assert tuple(params[self.name + '/kernel'].shape[:2]) == (k, k) and params[self.name + '/kernel'].shape[3] == ch
    def __call__(self, x):
      return T(net_oracle._conv_nhwc(raw(x), params[self.name + '/kernel'], params[self.name + '/bias'], self.s, self.same))

  class MaxPool2D(object):
    def __init__(self, pool_size, padding, strides):
      assert (pool_size, padding, strides) == (3, 'same', 2)
    def __call__(self, x): return T(net_oracle._maxpool_same_nhwc(raw(x)))

  class Dense(object):
    def __init__(self, units, name=None):
      self.name = order['dense'][counters['dense']]; counters['dense'] += 1
      assert params[self.name + '/kernel'].shape[1] == units
    def __call__(self, x): return T(raw(x) @ params[self.name + '/kernel'] + params[self.name + '/bias'])

  class Flatten(object):
    def __call__(self, x): return T(raw(x).reshape(raw(x).shape[0], -1))

  class LSTMCell(object):
    def __init__(self, units): self.units = units
    def get_initial_state(self, batch_size, dtype):
      return [T(torch.zeros(batch_size, self.units)), T(torch.zeros(batch_size, self.units))]   # Keras: [h, c]
    def __call__(self, x, state):
      h, c = net_oracle.lstm_cell(params, raw(x), raw(state[0]), raw(state[1]))
      return T(h), [T(h), T(c)]

  tf.keras = types.ModuleType('keras'); tf.keras.layers = types.ModuleType('layers')
  for cls in (Conv2D, MaxPool2D, Dense, Flatten, LSTMCell):
    setattr(tf.keras.layers, cls.__name__, cls)
  return tf, order


def main():
  import net_golden_params as G
  A, OBS, T1, B = G.A, G.OBS, G.T1, G.B
  p = G.make_params()
  tf, order = build_tf(p, lambda logits: logits.argmax(-1))
  # creation order inside the reference constructors (dmlab/networks.py:29-44, 74-89)
  for s in range(3):
    order['conv'] += ['stack%d/conv' % s, 'stack%d/res_0/conv2d_0' % s, 'stack%d/res_1/conv2d_0' % s,
                      'stack%d/res_0/conv2d_1' % s, 'stack%d/res_1/conv2d_1' % s]
  order['dense'] += ['conv_to_linear', 'policy_logits', 'baseline']
  sys.modules['tensorflow'] = tf
  seed_rl = types.ModuleType('seed_rl'); common = types.ModuleType('seed_rl.common'); utils = types.ModuleType('seed_rl.common.utils')
  ns = {'tf': tf}
  utils.batch_apply = _extract_function(os.path.join(REF, 'common/utils.py'), 'batch_apply', ns)
  seed_rl.common = common; common.utils = utils
  sys.modules.update({'seed_rl': seed_rl, 'seed_rl.common': common, 'seed_rl.common.utils': utils})
  ref = _load(os.path.join(REF, 'dmlab/networks.py'), 'ref_networks')
  agent = ref.ImpalaDeep(A)

  inp = G.make_inputs()
  obs, rew, done, prev, h0, c0 = (inp[k] for k in ('obs', 'rew', 'done', 'prev', 'h0', 'c0'))
  env = EnvOutput(T(rew), T(done), T(obs), T(np.zeros((T1, B), bool)), T(np.zeros((T1, B), np.int32)))
  with torch.no_grad():
    out, state = agent(T(prev), env, [T(h0), T(c0)], unroll=True)
    # single step (unroll=False) from the same state, like inference
    env1 = EnvOutput(T(rew[0]), T(done[0]), T(obs[0]), T(np.zeros(B, bool)), T(np.zeros(B, np.int32)))
    out1, state1 = agent(T(prev[0]), env1, [T(h0), T(c0)], unroll=False)
  np.savez_compressed(
      os.path.join(HERE, 'net_golden.npz'),
      logits=raw(out.policy_logits).numpy(), baseline=raw(out.baseline).numpy(), action=raw(out.action).numpy(),
      h=raw(state[0]).numpy(), c=raw(state[1]).numpy(),
      logits1=raw(out1.policy_logits).numpy(), baseline1=raw(out1.baseline).numpy(), h1=raw(state1[0]).numpy())
  print('wrote net_golden.npz; logits', tuple(raw(out.policy_logits).shape), 'step logits', tuple(raw(out1.policy_logits).shape))


if __name__ == '__main__':
  main()
