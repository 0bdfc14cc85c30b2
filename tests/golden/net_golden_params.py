"""Deterministic parameters / inputs of tests/golden/net_golden.npz (shared by the generator
and the test so that the 2.3 MB of weights need not be committed)."""
import numpy as np

A, OBS, T1, B = 6, (16, 16, 4), 5, 3


def make_params():
  import torch
  from oracle import net_oracle
  p = net_oracle.to_torch(net_oracle.init_params('deep', A, OBS, seed=4))
  rng = np.random.default_rng(9)
  for k in p:                                   # non-zero biases so that bias wiring is visible
    if k.endswith('bias'):
      p[k] = p[k] + torch.as_tensor(rng.normal(size=tuple(p[k].shape)).astype(np.float32)) * 0.1
  return p


def make_inputs():
  rng = np.random.default_rng(10)
  return dict(obs=rng.integers(0, 256, (T1, B) + OBS, dtype=np.uint8),
              rew=(rng.normal(size=(T1, B)) * 2).astype(np.float32), done=rng.random((T1, B)) < 0.3,
              prev=rng.integers(0, A, (T1, B)), h0=rng.normal(size=(B, 256)).astype(np.float32),
              c0=rng.normal(size=(B, 256)).astype(np.float32))
