"""GPU: R2D2 post-network kernels against oracle/r2d2_oracle.py."""
import numpy as np
import pytest
import torch

from oracle import r2d2_oracle as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('T,B,H,W,S', [(6, 2, 3, 4, 4), (9, 3, 5, 2, 3), (4, 1, 2, 2, 2), (20, 4, 84, 84, 4), (1, 2, 3, 3, 4)])
def test_stack_frames_bit_exact(T, B, H, W, S):
  from seed_rl_b200.atari import networks
  rng = np.random.default_rng(T + B)
  fr = rng.integers(0, 256, (T, B, H, W, 1), dtype=np.uint8)
  dn = rng.random((T, B)) < 0.3
  st = rng.integers(0, 1 << (8 * (S - 1)), (B, H * W)).astype(np.int32)
  want, want_state = R.stack_frames(fr.astype(np.float32), st, dn, S)
  got, got_state = networks.stack_frames(torch.as_tensor(fr).cuda(), torch.as_tensor(st).cuda(),
                                         torch.as_tensor(dn).cuda(), S)
  np.testing.assert_array_equal(got.cpu().numpy(), want.astype(np.uint8))
  np.testing.assert_array_equal(got_state.cpu().numpy(), want_state)
  # the reference's known-answer sequence (atari/networks_test.py:176-247) chained through the state
  z = networks.initial_frame_stacking_state(4, 1, [1])
  f = lambda v: torch.tensor(v, dtype=torch.uint8).reshape(len(v), 1, 1).cuda()
  d = lambda v: torch.tensor(v).reshape(len(v), 1).cuda()
  o, s = networks.stack_frames(f([1]), z, d([False]), 4)
  o, s = networks.stack_frames(f([2]), s, d([True]), 4)
  assert o.flatten().tolist() == [2, 0, 0, 0]
  o, s = networks.stack_frames(f([3, 4, 5, 6, 7, 8]), s, d([False, False, False, False, True, False]), 4)
  assert o[0].flatten().tolist() == [3, 2, 0, 0] and o[5].flatten().tolist() == [8, 7, 0, 0]


@pytest.mark.parametrize('T,B,A,n', [(16, 6, 18, 5), (101, 64, 18, 5), (4, 2, 3, 5), (12, 3, 4, 1)])
def test_loss_and_priorities_vs_oracle(T, B, A, n):
  from seed_rl_b200.agents.r2d2 import learner
  from seed_rl_b200.common import utils
  rng = np.random.default_rng(T + A)
  tq = rng.normal(size=(T, B, A)).astype(np.float32); gq = rng.normal(size=(T, B, A)).astype(np.float32)
  ra = rng.integers(0, A, (T, B)); r = rng.normal(size=(T, B)).astype(np.float32); d = rng.random((T, B)) < 0.1
  w = rng.random(B).astype(np.float32) + 0.1
  loss, prio, abs_td = R.loss_and_priorities(tq, tq.argmax(-1), gq, ra, r, d, 0.997, n_steps=n)
  c = lambda a: torch.as_tensor(a).cuda()
  env = utils.EnvOutput(c(r), c(d), None, None, None)
  got_loss, got_prio, dq = learner.compute_loss_and_priorities_from_agent_outputs(
      learner.AgentOutput(None, c(tq)), learner.AgentOutput(None, c(gq)), env, learner.AgentOutput(c(ra), None),
      0.997, n_steps=n, importance_weights=c(w))
  np.testing.assert_allclose(got_loss.cpu().numpy(), loss, rtol=2e-5, atol=1e-6)
  np.testing.assert_allclose(got_prio.cpu().numpy(), prio, rtol=2e-5, atol=1e-6)
  # gradient of mean_b(w_b loss_b): -(w_b / B) * td on the replayed action, zero elsewhere / last row
  tt, bb = np.meshgrid(np.arange(T - 1), np.arange(B), indexing='ij')
  target = R.value_function_rescaling(R.n_step_bellman_target(
      r, d, R.inverse_value_function_rescaling(gq[np.arange(T)[:, None], np.arange(B)[None], tq.argmax(-1)]), 0.997, n)[1:])
  want = np.zeros((T, B, A), np.float32)
  want[tt, bb, ra[:-1]] = -(w[None] / B) * (target - tq[tt, bb, ra[:-1]])
  np.testing.assert_allclose(dq.cpu().numpy(), want, rtol=2e-4, atol=1e-6)   # td ~ 0 cancels to ~1e-7 abs


def test_replay_sample_and_clip():
  from seed_rl_b200.agents.r2d2 import learner
  rng = np.random.default_rng(0)
  prio = rng.random(100).astype(np.float32) + 0.01
  u = rng.random(4096).astype(np.float32)
  idx, wts, probs = learner.replay_sample(torch.as_tensor(prio).cuda(), 70, 4096, 0.9, 0.6, uniforms=torch.as_tensor(u).cuda())
  p = R.replay_probabilities(prio, 70, 0.9)
  np.testing.assert_allclose(probs.cpu().numpy(), p, rtol=5e-5)   # powf vs numpy power: 1.3e-5 seen
  i = idx.cpu().numpy()
  assert i.min() >= 0 and i.max() < 70
  np.testing.assert_allclose(wts.cpu().numpy(), R.replay_importance_weights(p, i, 0.6), rtol=5e-5)
  # inverse-CDF draw: index = first i with cdf_i > u * total
  cdf = np.cumsum(np.power(prio[:70], np.float32(0.9), dtype=np.float32), dtype=np.float32)
  np.testing.assert_array_equal(i, np.minimum(np.searchsorted(cdf, u * cdf[-1], side='right'), 69))
  freq = np.bincount(i, minlength=70) / len(i)
  assert np.abs(freq - p).max() < 0.02                                   # statistical, like utils_test.py
  g = rng.normal(size=1 << 20).astype(np.float32)
  gc = torch.as_tensor(g).cuda()
  norm = learner.clip_by_global_norm(gc, 40.0)
  n64 = np.sqrt((g.astype(np.float64) ** 2).sum())
  np.testing.assert_allclose(float(norm), n64, rtol=1e-5)
  np.testing.assert_allclose(gc.cpu().numpy(), g * np.float32(40.0 / max(n64, 40.0)), rtol=1e-5)
