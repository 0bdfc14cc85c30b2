"""CPU: the R2D2 oracle (test infrastructure for the next round's CUDA
path) against (i) the known-answer cases of the reference's own tests
(agents/r2d2/learner_test.py:60-70, 114-198) and (ii) tests/golden/r2d2_golden.npz = outputs
of the unmodified reference functions executed over the numpy shim."""
import os

import numpy as np
import pytest

from oracle import r2d2_oracle as R

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'r2d2_golden.npz'))
col = lambda *v: np.array([v], np.float32).T


def test_value_function_rescaling_reference_cases():          # learner_test.py:114-140
  for x in np.linspace(-100., 100.):
    np.testing.assert_allclose(R.inverse_value_function_rescaling(R.value_function_rescaling(x)), x,
                               rtol=1e-6 * 100, atol=2e-4)
  assert R.value_function_rescaling(0.) == 0 and R.inverse_value_function_rescaling(0.) == 0
  assert R.value_function_rescaling(1000.) > 10. and R.value_function_rescaling(-1000.) < -10.
  np.testing.assert_allclose(R.value_function_rescaling([0., 3., -3.]), [0., 1 + 3e-3, -1 - 3e-3], rtol=1e-6)
  np.testing.assert_allclose(R.inverse_value_function_rescaling([0., 1 + 3e-3, -1 - 3e-3]), [0., 3, -3], atol=2e-4)


def test_value_function_rescaling_golden():
  np.testing.assert_array_equal(R.value_function_rescaling(G['resc_x']), G['resc_h'])
  np.testing.assert_allclose(R.inverse_value_function_rescaling(G['resc_x']), G['resc_hinv'], rtol=1e-6, atol=1e-6)


def test_n_step_bellman_target_reference_cases():             # learner_test.py:142-198
  t = R.n_step_bellman_target(col(1., 2., 3.), col(0, 0, 0) > 0, col(100, 200, 300), 0.9, 1)
  np.testing.assert_allclose(t, col(1 + 0.9 * 100, 2 + 0.9 * 200, 3 + 0.9 * 300), rtol=1e-6)
  t = R.n_step_bellman_target(col(1., 2., 3.), col(0, 1, 0) > 0, col(100, 200, 300), 0.9, 1)
  np.testing.assert_allclose(t, col(1 + 0.9 * 100, 2, 3 + 0.9 * 300), rtol=1e-6)
  t = R.n_step_bellman_target(col(1., 2., 3.), col(0, 0, 0) > 0, col(100, 200, 300), 0.9, 2)
  np.testing.assert_allclose(t, col(1 + 0.9 * 2 + 0.9 ** 2 * 200, 2 + 0.9 * 3 + 0.9 ** 2 * 300, 3 + 0.9 * 300),
                             rtol=1e-6)
  t = R.n_step_bellman_target(col(1., 2., 3., 4., 5., 6., 7.), col(0, 0, 0, 1, 0, 0, 0) > 0,
                              col(100, 200, 300, 400, 500, 600, 700), 0.9, 3)
  np.testing.assert_allclose(t, col(
      1 + 0.9 * 2 + 0.9 ** 2 * 3 + 0.9 ** 3 * 300, 2 + 0.9 * 3 + 0.9 ** 2 * 4, 3 + 0.9 * 4, 4,
      5 + 0.9 * 6 + 0.9 ** 2 * 7 + 0.9 ** 3 * 700, 6 + 0.9 * 7 + 0.9 ** 2 * 700, 7 + 0.9 * 700), rtol=1e-6)


@pytest.mark.parametrize('name', ['a', 'b', 'c', 'd'])
def test_n_step_bellman_target_golden(name):
  r, d, q = G['nstep_%s_in' % name]
  n, gamma = G['nstep_%s_cfg' % name]
  got = R.n_step_bellman_target(r, d > 0, q, float(gamma), int(n))
  np.testing.assert_allclose(got, G['nstep_%s_out' % name], rtol=2e-6, atol=1e-6)


def test_loss_and_priorities_golden():
  loss, prio, abs_td = R.loss_and_priorities(G['loss_train_q'], G['loss_train_action'], G['loss_target_q'],
                                             G['loss_replay_action'], G['loss_reward'], G['loss_done'], 0.997)
  np.testing.assert_allclose(loss, G['loss_out'], rtol=1e-5)
  np.testing.assert_allclose(prio, G['loss_priorities'], rtol=1e-5)
  assert abs_td.shape == (15, 6)


def test_get_envs_epsilon():                                   # learner_test.py:60-70
  e = R.get_envs_epsilon(np.arange(20), 10, 10, 1e-3)
  np.testing.assert_allclose(e[10:], [1e-3] * 10, rtol=1e-6)
  np.testing.assert_allclose(e[0], 0.4, rtol=1e-6)
  np.testing.assert_allclose(e[9], 0.4 ** 8, rtol=1e-5)
  np.testing.assert_allclose(e, G['eps_out'], rtol=1e-6)


def test_prioritized_replay_probabilities_and_weights():       # common/utils.py:327-352
  prio = np.array([1., 2., 4., 0., 0.], np.float32)            # ring of 5, three inserted
  p = R.replay_probabilities(prio, 3, 0.5)
  np.testing.assert_allclose(p, np.sqrt([1., 2., 4.]) / np.sqrt([1., 2., 4.]).sum(), rtol=1e-6)
  w = R.replay_importance_weights(p, [0, 2, 2, 1], 0.6)
  raw = ((1. / 3) / p[[0, 2, 2, 1]]) ** 0.6
  np.testing.assert_allclose(w, raw / raw.max(), rtol=1e-6)
  assert w.max() == 1.0
  np.testing.assert_array_equal(R.replay_insert_indices(3, 4, 5), [3, 4, 0, 1])
  # full ring: every slot counts
  assert len(R.replay_probabilities(np.ones(5, np.float32), 9, 1.0)) == 5


def _stack(frames, state, done, S=4):
  return R.stack_frames(np.asarray(frames, np.float32), state, np.asarray(done), S)


def test_stack_frames_reference_cases():                       # atari/networks_test.py:176-247
  zero = R.initial_frame_stacking_state(4, 1, [1])
  assert zero.dtype == np.int32 and zero.shape == (1, 1)
  out, st = _stack([[[1]]], zero, [[False]])
  np.testing.assert_array_equal(out, [[[1, 0, 0, 0]]])
  out, st = _stack([[[2]]], st, [[False]])
  np.testing.assert_array_equal(out, [[[2, 1, 0, 0]]])
  out, st = _stack([[[3]], [[4]], [[5]], [[6]], [[7]], [[8]]], st, [[False]] * 6)
  assert out.shape[0] == 6
  np.testing.assert_array_equal(out[0], [[3, 2, 1, 0]])
  np.testing.assert_array_equal(out[5], [[8, 7, 6, 5]])
  # done resets the stack
  out, st = _stack([[[1]]], zero, [[False]])
  out, st = _stack([[[2]]], st, [[True]])
  np.testing.assert_array_equal(out, [[[2, 0, 0, 0]]])
  out, st = _stack([[[3]], [[4]], [[5]], [[6]], [[7]], [[8]]], st,
                   [[False], [False], [False], [False], [True], [False]])
  np.testing.assert_array_equal(out[0], [[3, 2, 0, 0]])
  np.testing.assert_array_equal(out[5], [[8, 7, 0, 0]])
  # stack_size 1 is the identity with an empty state; errors as in the reference
  f = np.zeros((2, 1, 3, 3, 1), np.float32)
  o, s = R.stack_frames(f, (), np.zeros((2, 1), bool), 1)
  assert o is not None and s == ()
  with pytest.raises(ValueError):
    R.stack_frames(f, zero, np.zeros((3, 1), bool), 4)
  with pytest.raises(ValueError):
    R.stack_frames(f, zero, np.zeros((2, 1), bool), 5)


@pytest.mark.parametrize('name', ['a', 'b', 'c'])
def test_stack_frames_golden_bit_exact(name):
  out, st = R.stack_frames(G['stack_%s_frames' % name], G['stack_%s_state' % name], G['stack_%s_done' % name],
                           int(G['stack_%s_size' % name]))
  np.testing.assert_array_equal(out, G['stack_%s_out' % name])
  np.testing.assert_array_equal(st, G['stack_%s_new_state' % name])
  assert st.dtype == np.int32


def test_dueling_lstm_dqn_net_structure():
  """atari/networks_test.py:78-117: unrolls run with and without frame stacking, the torso sees
  stack_size channels, the core input is 512 + num_actions + 1 wide; plus the invariants the
  code states (dueling advantages are mean-free, greedy action, state reset on done)."""
  import torch
  from oracle import net_oracle, r2d2_net_oracle as N
  rng = np.random.default_rng(0)
  OBS, A, T, B = [84, 84, 1], 37, 5, 2
  for S in (4, 1):
    specs = dict(N.param_specs(A, OBS, S))
    assert specs['body/conv0/kernel'] == (8, 8, S if S > 1 else 1, 32)
    assert specs['core/kernel'] == (512 + A + 1, 4 * 512)                    # networks_test.py:115-117
    assert specs['body/dense/kernel'][0] == 7 * 7 * 64 and 'advantage/head/bias' not in specs
    p = net_oracle.to_torch(N.init_params(A, OBS, S, seed=1))
    obs = rng.integers(0, 256, [T, B] + OBS, dtype=np.uint8)
    done = rng.random((T, B)) < 0.3
    st = N.initial_state(B, OBS, S)
    out, st2 = N.unroll(p, rng.integers(0, A, (T, B)), rng.normal(size=(T, B)), done, obs, st, A, S)
    assert tuple(out.q_values.shape) == (T, B, A) and tuple(out.action.shape) == (T, B)
    assert torch.equal(out.action.long(), out.q_values.argmax(-1))
    assert isinstance(st2.frame_stacking_state, tuple) == (S == 1)
    # splitting the unroll at any step and carrying the state gives the same outputs
    _pa = rng.integers(0, A, (T, B)); _rw = rng.normal(size=(T, B))
    out_a, mid = N.unroll(p, _pa[:2], _rw[:2], done[:2], obs[:2], st, A, S)
    out_b, _ = N.unroll(p, _pa[2:], _rw[2:], done[2:], obs[2:], mid, A, S)
    full, _ = N.unroll(p, _pa, _rw, done, obs, st, A, S)
    np.testing.assert_allclose(torch.cat([out_a.q_values, out_b.q_values]).detach().numpy(),
                               full.q_values.detach().numpy(), rtol=1e-5, atol=1e-6)


def test_dueling_lstm_dqn_net_wiring_against_reference_source():
  """oracle/r2d2_net_oracle.py against tests/golden/r2d2_net_golden.npz = the UNMODIFIED reference
  atari/networks.py (stack_frames, _unroll_cell, DuelingLSTMDQNNet) + common/utils.batch_apply
  executed over a Keras-layer shim (tests/golden/make_golden_r2d2_net.py): same weights and
  inputs -> same Q values, greedy actions, final LSTM state and bit-identical packed frame state."""
  import sys
  import torch
  from oracle import r2d2_net_oracle as N
  sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
  import make_golden_r2d2_net as GEN
  g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'r2d2_net_golden.npz'))
  p = GEN.make_params(); i = GEN.make_inputs()
  with torch.no_grad():
    out, st = N.unroll(p, i['prev'], i['rew'], i['done'], i['obs'],
                       N.AgentState((torch.as_tensor(i['h0']), torch.as_tensor(i['c0'])), i['fstate']), GEN.A, GEN.S)
  np.testing.assert_allclose(out.q_values.numpy(), g['q'], rtol=1e-5, atol=1e-6)
  np.testing.assert_array_equal(out.action.numpy(), g['action'])
  np.testing.assert_allclose(st.core_state[0].numpy(), g['h'], rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(st.core_state[1].numpy(), g['c'], rtol=1e-5, atol=1e-6)
  np.testing.assert_array_equal(st.frame_stacking_state, g['fstate'])
  assert i['done'].any() and not i['done'].all()


def test_prioritized_replay_against_reference_class():
  """oracle replay functions against tests/golden/replay_golden.npz = the UNMODIFIED reference
  PrioritizedReplay (common/utils.py:260-370) executed over a tf.Variable stand-in
  (tests/golden/make_golden_replay.py): FIFO wrap-around insertion, probabilities over the
  filled part of the ring, inverse-CDF draws from the recorded uniforms, importance weights,
  priority updates."""
  g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'replay_golden.npz'))
  size, table, inserted = 7, np.zeros(7, np.float32), 0
  store = np.zeros((7, 3), np.float32)
  for s in range(3):
    pr = g['ins%d_prio' % s]
    idx = R.replay_insert_indices(inserted, len(pr), size)
    np.testing.assert_array_equal(idx, g['ins%d_idx' % s])
    table[idx] = pr; store[idx] = g['ins%d_vals' % s]; inserted += len(pr)
    assert inserted == int(g['smp%d_num_inserted' % s])
    np.testing.assert_array_equal(table, g['smp%d_prio_table' % s])
    p = R.replay_probabilities(table, inserted, 0.9)
    draw = np.minimum(np.searchsorted(np.cumsum(p.astype(np.float64)), g['smp%d_u' % s], side='right'), len(p) - 1)
    np.testing.assert_array_equal(draw, g['smp%d_idx' % s])
    np.testing.assert_allclose(R.replay_importance_weights(p, draw, 0.6), g['smp%d_w' % s], rtol=1e-5)
    np.testing.assert_array_equal(store[draw], g['smp%d_vals' % s])
    table[draw] = g['upd%d_prio' % s]                         # update_priorities
    np.testing.assert_array_equal(table, g['upd%d_table' % s])
  assert (g['ins2_idx'] < g['ins2_idx'][0]).any()             # the third insert wrapped around
  np.testing.assert_array_equal(g['uni_w'], np.ones(4, np.float32))   # priority_exp == 0: unit weights
  np.testing.assert_array_equal(g['uni_idx'], (g['uni_u'] * 7).astype(np.int64))
