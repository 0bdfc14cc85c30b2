"""utils.BatchAssembler host logic (zero-copy minibatch assembly) on CPU tensors:
column claims across slot boundaries, publication only when a slot is full, capacity back-pressure
(the reference's capacity-1 unroll queue blocks the inference thread the same way,
agents/vtrace/learner.py:336), release / reuse, close."""
import threading
import time

import pytest
import torch

from seed_rl_b200.common import utils

TS = utils.TensorSpec


def _mk(batch, slots=2):
  specs = (TS([], 'int64', 'a'), (TS([3], 'float32', 'x'), TS([], 'bool', 'd')))
  state = (TS([5], 'float32', 'h'), TS([5], 'float32', 'c'))
  return utils.BatchAssembler(specs, state, full_length=4, batch_size=batch, slots=slots, device='cpu')


def test_claims_fill_columns_in_order_and_split_across_slots():
  asm = _mk(4)
  s0, c0, n0 = asm.claim(3)
  assert (c0, n0) == (0, 3)
  asm.commit()
  with pytest.raises(TimeoutError):
    asm.get(timeout=0.05)                        # 3 of 4 columns: not published
  s1, c1, n1 = asm.claim(3)                      # only one column left in this slot
  assert (s1, c1, n1) == (s0, 3, 1)
  asm.commit()                                   # full -> published
  s2, c2, n2 = asm.claim(2)                      # the remainder goes to the next slot
  assert s2 != s0 and (c2, n2) == (0, 2)
  slot, state, (a, (x, d)) = asm.get(timeout=1.0)
  assert slot == s0
  assert tuple(a.shape) == (4, 4) and tuple(x.shape) == (4, 4, 3) and d.dtype == torch.bool
  assert tuple(state[0].shape) == (4, 5)
  assert a.data_ptr() == asm.field(slot, 0).data_ptr()        # views of the slot, no copy


def test_back_pressure_blocks_until_release_and_close_unblocks():
  asm = _mk(2, slots=2)
  for _ in range(2):                             # fill both slots
    asm.claim(2); asm.commit()
  got = []
  th = threading.Thread(target=lambda: got.append(asm.claim(1)))
  th.start(); time.sleep(0.2)
  assert not got, 'claim must block while every slot is full or in use'
  slot, _, _ = asm.get(timeout=1.0)
  asm.release(slot)                              # the learner is done with it
  th.join(2.0)
  assert got and got[0][0] == slot and got[0][1:] == (0, 1)
  asm.close()
  with pytest.raises(utils.QueueClosedError):
    while True:
      s, _, _ = asm.get()                        # drains the remaining full slot, then raises
      asm.release(s)
