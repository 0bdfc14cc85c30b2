"""Generates tests/golden/replay_golden.npz.  Run ONLY in the build container:
    python tests/golden/make_golden_replay.py

Executes the UNMODIFIED reference class common/utils.py::PrioritizedReplay (:260-370, pulled
out by AST) over a small numpy stand-in for tf.Variable / scatter / gather, with
tf.random.categorical replaced by an inverse-CDF draw from recorded uniforms (the reference
samples with Philox, which cannot be reproduced; the DISTRIBUTION and everything downstream --
probabilities, importance weights, FIFO wrap-around insertion, priority updates -- is pinned)."""
import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/common/utils.py'
f32 = np.float32


class V(object):                                   # tf.Variable / tensor stand-in
  def __init__(self, a, dtype=None): self.a = np.array(a.a if isinstance(a, V) else a, dtype=dtype)
  @property
  def shape(self): return list(self.a.shape)
  @property
  def dtype(self): return self.a.dtype
  def __getitem__(self, i):
    if isinstance(i, slice):
      i = slice(*[int(raw(x)) if x is not None else None for x in (i.start, i.stop, i.step)])
    return V(self.a[i])
  def __pow__(self, e): return V(np.power(self.a, f32(e)).astype(f32))
  def __truediv__(self, o): return V((self.a / raw(o)).astype(self.a.dtype))
  def __rtruediv__(self, o): return V((f32(o) / self.a).astype(f32))
  def __itruediv__(self, o): return V((self.a / raw(o)).astype(self.a.dtype))
  def __add__(self, o): return V(self.a + raw(o))
  def __mod__(self, o): return V(self.a % raw(o))
  def __eq__(self, o): return bool(np.all(self.a == raw(o)))
  def __int__(self): return int(self.a)
  def __index__(self): return int(self.a)
  def assign_add(self, n): self.a = self.a + raw(n)
  def batch_scatter_update(self, sl): self.a[np.asarray(raw(sl.indices))] = raw(sl.values)
  def sparse_read(self, idx): return V(self.a[np.asarray(raw(idx))])


def raw(x): return x.a if isinstance(x, V) else x


def build_tf(uniform_source):
  tf = types.ModuleType('tensorflow')
  tf.float32, tf.int64 = f32, np.int64
  class Module(object):
    def __init__(self, name=None): pass
    @staticmethod
    def with_name_scope(fn): return fn
  tf.Module = Module
  tf.function = lambda fn: fn
  tf.Variable = lambda init, dtype=None: V(init, dtype)
  tf.zeros = lambda shape, dtype=f32: V(np.zeros([int(s) for s in shape], dtype))
  tf.constant = lambda v, dtype=None: V(np.asarray(v, dtype))
  tf.convert_to_tensor = lambda x: x if isinstance(x, V) else V(x)
  tf.range = lambda a, b: V(np.arange(int(raw(a)), int(raw(b))))
  tf.IndexedSlices = lambda values, indices: types.SimpleNamespace(values=values, indices=indices)
  tf.cast = lambda x, dt: V(np.asarray(raw(x)).astype(dt))
  tf.minimum = lambda a, b: V(np.minimum(raw(a), raw(b)))
  tf.reduce_sum = lambda x: V(np.sum(raw(x), dtype=f32))
  tf.reduce_max = lambda x: V(np.max(raw(x)))
  tf.gather = lambda p, i: V(raw(p)[np.asarray(raw(i))])
  tf.ones_like = lambda x, dtype=None: V(np.ones_like(raw(x), dtype=dtype))
  tf.math = types.ModuleType('math'); tf.math.log = lambda x: V(np.log(raw(x)))
  tf.debugging = types.ModuleType('debugging'); tf.debugging.assert_greater_equal = lambda a, b, message=None: None
  nest = types.ModuleType('nest')
  nest.map_structure = lambda fn, *s: (type(s[0])(nest.map_structure(fn, *xs) for xs in zip(*s))
                                      if isinstance(s[0], (list, tuple)) else fn(*s))
  nest.flatten = lambda s: sum((nest.flatten(x) for x in s), []) if isinstance(s, (list, tuple)) else [s]
  nest.assert_same_structure = lambda a, b: None
  tf.nest = nest
  rnd = types.ModuleType('random')
  def categorical(logits, num_samples):             # inverse CDF of recorded uniforms over softmax(logits)
    lg = np.asarray(raw(logits[0]), np.float64)
    p = np.exp(lg - lg.max()); p /= p.sum()
    u = uniform_source(num_samples)
    return [V(np.minimum(np.searchsorted(np.cumsum(p), u, side='right'), len(p) - 1).astype(np.int64))]
  rnd.categorical = categorical
  rnd.uniform = lambda shape, maxval, dtype: V((uniform_source(shape[0]) * int(raw(maxval))).astype(np.int64))
  tf.random = rnd
  return tf


def extract_class(path, name, ns):
  tree = ast.parse(open(path).read())
  for node in tree.body:
    if isinstance(node, ast.ClassDef) and node.name == name:
      exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
      return ns[name]
  raise KeyError(name)


def main():
  rng = np.random.default_rng(21)
  draws = []
  def uniform_source(n):
    u = rng.random(int(n)); draws.append(u); return u
  tf = build_tf(uniform_source)
  PR = extract_class(REF, 'PrioritizedReplay', {'tf': tf})
  spec = types.SimpleNamespace(shape=[3], dtype=f32)
  rb = PR(7, [spec], 0.6)
  out = {}
  step = 0
  for n_ins in (3, 3, 4):                            # third insert wraps around the ring of 7
    vals = rng.normal(size=(n_ins, 3)).astype(f32); pr = (rng.random(n_ins) + 0.05).astype(f32)
    idx = rb.insert([V(vals)], V(pr))
    out['ins%d_vals' % step] = vals; out['ins%d_prio' % step] = pr; out['ins%d_idx' % step] = raw(idx)
    indices, weights, sampled = rb.sample(5, 0.9)
    out['smp%d_u' % step] = draws[-1]; out['smp%d_idx' % step] = raw(indices); out['smp%d_w' % step] = raw(weights)
    out['smp%d_vals' % step] = raw(sampled[0]); out['smp%d_prio_table' % step] = raw(rb._priorities).copy()
    out['smp%d_num_inserted' % step] = np.asarray(int(rb.num_inserted))
    newp = (rng.random(5) + 0.05).astype(f32)
    rb.update_priorities(indices, V(newp))
    out['upd%d_prio' % step] = newp; out['upd%d_table' % step] = raw(rb._priorities).copy()
    step += 1
  indices, weights, _ = rb.sample(4, 0)                # priority_exp == 0: uniform, unit weights
  out['uni_u'] = draws[-1]; out['uni_idx'] = raw(indices); out['uni_w'] = raw(weights)
  np.savez_compressed(os.path.join(HERE, 'replay_golden.npz'), **out)
  print('wrote replay_golden.npz:', len(out), 'arrays')


if __name__ == '__main__':
  main()
