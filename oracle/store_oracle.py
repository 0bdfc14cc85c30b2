"""CPU ORACLE (test infrastructure, NOT product code) -- numpy restatement of the
per-environment state tables on the inference path:

  UnrollStore   /root/reference/common/utils.py:119-257
  Aggregator    /root/reference/common/utils.py:461-543

Pinned by the scripted sequences of reference tests/utils_test.py:70-301 (replayed
in tests/test_store.py).  Fields are a flat list of arrays (the tf.nest is
flattened by the caller).
"""
import numpy as np


class UnrollStore(object):

  def __init__(self, num_envs, unroll_length, field_specs, num_overlapping_steps=0):
    """field_specs: list of (shape_tuple, np.dtype)."""
    self._full_length = num_overlapping_steps + unroll_length + 1     # :129
    self._unroll_length = unroll_length
    self._num_overlapping_steps = num_overlapping_steps
    self._state = [np.zeros((num_envs, self._full_length) + tuple(s), d)
                   for s, d in field_specs]                            # :131-139
    self._index = np.full([num_envs], num_overlapping_steps, np.int32)  # :142-145

  def append(self, env_ids, values):
    env_ids = np.asarray(env_ids)
    if len(np.unique(env_ids)) != len(env_ids):                       # :173-176
      raise ValueError('Duplicate environment ids in store')
    for v in values:                                                  # :178-184
      if np.asarray(v).shape[0] != env_ids.shape[0]:
        raise ValueError('Batch dimension must equal the number of environments')
    cur = self._index[env_ids]                                        # :187
    for s, v in zip(self._state, values):                             # :188-190
      s[env_ids, cur] = v
    self._index[env_ids] += 1                                         # :194
    return self._complete_unrolls(env_ids)

  def reset(self, env_ids):
    env_ids = np.asarray(env_ids, np.int64)
    self._index[env_ids] = self._num_overlapping_steps                # :207-208
    j = self._num_overlapping_steps                                   # :212-225
    for s in self._state:
      s[env_ids, :j] = 0

  def _complete_unrolls(self, env_ids):
    idx = self._index[env_ids]                                        # :229
    done_ids = env_ids[idx == self._full_length].astype(np.int64)     # :230-233
    unrolls = [s[done_ids].copy() for s in self._state]               # :234-235
    j = self._num_overlapping_steps + 1                               # :240-252
    for s in self._state:
      s[done_ids, :j] = s[done_ids, self._full_length - j:]
    self._index[done_ids] = 1 + self._num_overlapping_steps           # :254-255
    return done_ids, unrolls


class Aggregator(object):

  def __init__(self, num_envs, field_specs):
    self._state = [np.zeros((num_envs,) + tuple(s), d) for s, d in field_specs]

  def reset(self, env_ids):                                           # :481-485
    for s in self._state:
      s[np.asarray(env_ids, np.int64)] = 0

  def add(self, env_ids, values):                                     # :488-501
    for s, v in zip(self._state, values):
      np.add.at(s, np.asarray(env_ids, np.int64), v)

  def read(self, env_ids):                                            # :504-516
    return [s[np.asarray(env_ids, np.int64)] for s in self._state]

  def replace(self, env_ids, values):                                 # :519-543
    env_ids = np.asarray(env_ids, np.int64)
    if env_ids.ndim != 1:
      raise ValueError('Invalid rank for aggregator')
    if len(np.unique(env_ids)) != len(env_ids):
      raise ValueError('Duplicate environment ids in Aggregator')
    for s, v in zip(self._state, values):
      s[env_ids] = v
