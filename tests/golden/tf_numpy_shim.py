"""A numpy-fp32 stand-in for the handful of TensorFlow ops the reference's
V-trace sources use.  TEST INFRASTRUCTURE ONLY (used by make_golden.py in the
build container, where /root/reference exists but TensorFlow does not).

With this module installed as `sys.modules['tensorflow']`, the UNMODIFIED
reference files

    /root/reference/common/vtrace.py                               (from_importance_weights)
    /root/reference/agents/policy_gradient/modules/advantages.py  (vtrace, 2nd impl)

can be imported and executed line by line; every tf.* call they make lands on
the same-named numpy float32 operation below.  This is how the golden vectors in
tests/golden/*.npz were produced ("the reference's own control flow over numpy
fp32 arithmetic") -- it pins the oracle's structure (slicing, clip order, scan
direction, bootstrap handling) to the reference source rather than to our
reading of it.  What it can NOT pin is TF's Eigen `exp` rounding (<= 1 ulp).
"""
import contextlib
import sys
import types

import numpy as np


class Shape(tuple):
  @property
  def ndims(self):
    return len(self)

  rank = ndims

  def assert_has_rank(self, r):
    if len(self) != r:
      raise ValueError('Shape %s must have rank %d' % (tuple(self), r))


def _raw(x):
  if isinstance(x, Tensor):
    return x.a
  if isinstance(x, (list, tuple)):
    return np.stack([_raw(e) for e in x]) if len(x) else np.zeros((0,), np.float32)
  return x


class Tensor(object):
  """Immutable value wrapper; arithmetic stays in the array's dtype (fp32)."""
  __array_priority__ = 1000

  def __init__(self, a):
    self.a = np.asarray(a)

  # -- structure ---------------------------------------------------------
  @property
  def shape(self):
    return Shape(self.a.shape)

  @property
  def dtype(self):
    return self.a.dtype

  def __len__(self):
    return self.a.shape[0]

  def __iter__(self):
    return (Tensor(self.a[i]) for i in range(self.a.shape[0]))

  def __getitem__(self, idx):
    return Tensor(self.a[idx])

  def __array__(self, dtype=None, copy=None):
    return self.a if dtype is None else self.a.astype(dtype)

  def numpy(self):
    return self.a

  # -- arithmetic --------------------------------------------------------
  def _bin(self, other, fn, rev=False):
    o = _raw(other)
    if isinstance(o, (float, int)) and self.a.dtype == np.float32:
      o = np.float32(o)
    return Tensor(fn(o, self.a) if rev else fn(self.a, o))

  def __add__(self, o): return self._bin(o, np.add)
  def __radd__(self, o): return self._bin(o, np.add, True)
  def __sub__(self, o): return self._bin(o, np.subtract)
  def __rsub__(self, o): return self._bin(o, np.subtract, True)
  def __mul__(self, o): return self._bin(o, np.multiply)
  def __rmul__(self, o): return self._bin(o, np.multiply, True)
  def __truediv__(self, o): return self._bin(o, np.divide)
  def __neg__(self): return Tensor(-self.a)
  def __invert__(self): return Tensor(~self.a)


def _wrap1(fn):
  def f(x, name=None):
    return Tensor(fn(_raw(x)))
  return f


def build_module():
  tf = types.ModuleType('tensorflow')
  tf.float32 = np.float32
  tf.bool = np.bool_
  tf.int32 = np.int32
  tf.int64 = np.int64

  def convert_to_tensor(x, dtype=None, name=None):
    a = np.asarray(_raw(x))
    if dtype is not None:
      a = a.astype(dtype)
    return Tensor(a)

  def cast(x, dtype, name=None):
    return Tensor(np.asarray(_raw(x)).astype(dtype))

  def minimum(x, y, name=None):
    x, y = _raw(x), _raw(y)
    if isinstance(x, float): x = np.float32(x)
    if isinstance(y, float): y = np.float32(y)
    return Tensor(np.minimum(x, y))

  def concat(values, axis, name=None):
    return Tensor(np.concatenate([_raw(v) for v in values], axis=axis))

  def expand_dims(x, axis, name=None):
    return Tensor(np.expand_dims(_raw(x), axis))

  def add(x, y, name=None):
    return Tensor(np.add(_raw(x), _raw(y)))

  tf.convert_to_tensor = convert_to_tensor
  tf.cast = cast
  tf.exp = _wrap1(np.exp)
  tf.minimum = minimum
  tf.concat = concat
  tf.expand_dims = expand_dims
  tf.zeros_like = lambda x, dtype=None, name=None: Tensor(
      np.zeros_like(_raw(x), dtype=dtype))
  tf.add = add
  tf.stop_gradient = lambda x, name=None: Tensor(_raw(x))
  tf.name_scope = lambda name: contextlib.nullcontext()
  tf.math = types.ModuleType('tensorflow.math')
  tf.math.log = lambda x, name=None: Tensor(np.log(np.float32(_raw(x))))
  tf.Module = object   # base class of unrelated estimator classes in advantages.py
  return tf


def install():
  """Installs the shim as `tensorflow` (+ a no-op `gin`). Returns the module."""
  tf = build_module()
  sys.modules['tensorflow'] = tf
  gin = types.ModuleType('gin')
  gin.configurable = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
  sys.modules['gin'] = gin
  return tf
