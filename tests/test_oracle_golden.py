"""CPU: the oracle against the reference's golden vectors / known-answer tests
(SURVEY 8c).  Fixtures in tests/golden/ were produced by tests/golden/make_golden.py
from the reference's own sources."""
import os

import numpy as np
import pytest

from oracle import loss_oracle, net_oracle, optim_oracle, store_oracle, vtrace_oracle

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'vtrace_golden.npz'))


def _inputs(prefix):
  names = ['target_action_log_probs', 'behaviour_action_log_probs', 'discounts', 'rewards',
           'values', 'bootstrap_value']
  return {n: G['%s_%s' % (prefix, n)] for n in names}


def test_vtrace_known_answer_reference_test():
  """reference tests/vtrace_test.py:118-145, compared with assertAllClose's 1e-6."""
  r = vtrace_oracle.from_importance_weights(
      **_inputs('A'), clip_rho_threshold=3.7, clip_pg_rho_threshold=2.2)
  # the reference test's own O(T^2) ground truth (float64 discounts)
  np.testing.assert_allclose(r.vs, G['A_gt_vs'], rtol=1e-6, atol=1e-5)
  np.testing.assert_allclose(r.pg_advantages, G['A_gt_pg'], rtol=1e-6, atol=1e-5)
  # the reference's common/vtrace.py executed over numpy fp32
  np.testing.assert_array_equal(r.vs, G['A_ref_vs'])
  np.testing.assert_array_equal(r.pg_advantages, G['A_ref_pg'])
  # the numbers printed in SURVEY 8(c)
  np.testing.assert_allclose(r.vs[0], [0.929816, 0.556437, 0.816895, 1.189767, 1.654995], atol=2e-6)
  np.testing.assert_allclose(r.pg_advantages[3],
                             [118.50515, 88.928314, 78.00101, 69.30495, 64.79132], rtol=1e-6)


@pytest.mark.parametrize('case,kw', [
    ('B1', {}),
    ('B2', dict(clip_rho_threshold=None, clip_pg_rho_threshold=None)),
    ('B3', dict(clip_rho_threshold=3.7, clip_pg_rho_threshold=2.2, lambda_=0.9))])
def test_vtrace_cfg1_against_reference_source(case, kw):
  r = vtrace_oracle.from_importance_weights(**_inputs('B'), **kw)
  np.testing.assert_array_equal(r.vs, G[case + '_ref_vs'])
  np.testing.assert_array_equal(r.pg_advantages, G[case + '_ref_pg'])


def test_vtrace_lambda_done_against_second_reference_implementation():
  """agents/policy_gradient/modules/advantages_test.py:129-150."""
  v = G['C_values']
  disc = (0.99 * (~G['C_done'])).astype(np.float32)
  r = vtrace_oracle.from_importance_weights(G['C_tlp'], G['C_blp'], disc, G['C_rewards'],
                                            v[:-1], v[-1], lambda_=0.95)
  np.testing.assert_allclose(r.vs, G['C_adv_targets'], rtol=1e-6, atol=1e-6)
  np.testing.assert_array_equal(r.vs, G['C_ref_vs'])


def test_vtrace_extra_trailing_dims():
  r = vtrace_oracle.from_importance_weights(**_inputs('D'))
  np.testing.assert_array_equal(r.vs, G['D_ref_vs'])
  np.testing.assert_array_equal(r.pg_advantages, G['D_ref_pg'])


def test_vtrace_rank_check():
  i = _inputs('B')
  i['bootstrap_value'] = i['bootstrap_value'][None]
  with pytest.raises(ValueError):
    vtrace_oracle.from_importance_weights(**i)


def test_log_probs_from_logits_and_actions():
  """reference tests/vtrace_test.py:88-115."""
  T, B, A = 7, 2, 3
  logits = np.arange(T * B * A, dtype=np.float32).reshape(T, B, A) + 10
  actions = np.random.default_rng(0).integers(0, A - 1, size=(T, B)).astype(np.int32)
  got = vtrace_oracle.categorical_log_prob(logits, actions)
  sm = np.exp(logits) / np.sum(np.exp(logits), axis=-1, keepdims=True)
  want = np.log(sm)[actions[..., None] == np.arange(A)].reshape(T, B)
  np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)


def test_entropy_uniform():
  np.testing.assert_allclose(vtrace_oracle.categorical_entropy(np.zeros((4, 18), np.float32)),
                             np.log(18.0), rtol=1e-6)


def test_loss_analytic_gradient_formulas_match_autograd():
  """The closed-form gradient the CUDA loss kernel implements (vtrace_kernels.cu phase D)
  restated in numpy, against torch autograd through the oracle's compute_loss."""
  rng = np.random.default_rng(3)
  T1, B, A = 6, 3, 5
  cfg = loss_oracle.default_config(kl_cost=0.1, entropy_cost=0.01, max_abs_reward=1.0,
                                   target_entropy=0.5)
  ll = rng.normal(size=(T1, B, A)).astype(np.float32)
  lb = rng.normal(size=(T1, B)).astype(np.float32)
  bl = rng.normal(size=(T1, B, A)).astype(np.float32)
  act = rng.integers(0, A, (T1, B)); rew = rng.normal(size=(T1, B)).astype(np.float32) * 2
  done = rng.random((T1, B)) < 0.2
  total, logs, dl, db, dep, aux = loss_oracle.loss_and_grads(cfg, ll, lb, bl, act, rew, done)
  assert np.all(dl[-1] == 0) and np.all(db[-1] == 0)
  T = T1 - 1
  N = T * B
  lsm = vtrace_oracle.log_softmax(ll[:-1]); p = np.exp(lsm)
  H = -(p * lsm).sum(-1)
  onehot = np.eye(A, dtype=np.float32)[act[:-1]]
  pg = aux['pg_advantages'].numpy(); verr = (aux['vs'] - lb[:-1]).numpy()
  mul = cfg.entropy_cost_adjustment_speed
  ec = cfg.entropy_cost
  g = (-(pg[..., None] + cfg.kl_cost) / N * (onehot - p) + ec / N * p * (lsm + H[..., None]))
  np.testing.assert_allclose(dl[:-1], g, rtol=2e-4, atol=1e-6)
  np.testing.assert_allclose(db[:-1], -cfg.baseline_cost * verr / N, rtol=2e-4, atol=1e-7)
  np.testing.assert_allclose(dep, mul * ec * (H.mean() - cfg.target_entropy), rtol=1e-4)
  np.testing.assert_allclose(
      logs['losses/total'], logs['losses/policy'] + logs['losses/V'] + logs['losses/entropy'] +
      logs['losses/kl'] + ec * (H.mean() - cfg.target_entropy), rtol=1e-5)


def test_keras_adam_first_step_is_lr_sized():
  p = np.ones(5, np.float32); g = np.full(5, 0.3, np.float32)
  p2, m, v = optim_oracle.keras_adam_step(p, g, np.zeros(5), np.zeros(5), 0, 0.01,
                                          beta1=0.9, beta2=0.999, eps=1e-7)
  np.testing.assert_allclose(p - p2, 0.01, rtol=1e-4)   # |step| ~= lr on the first step
  assert abs(optim_oracle.polynomial_decay(4.8e-4, 50, 100) - 2.4e-4) < 1e-12


def test_impala_deep_has_39_tensors():
  """reference tests/agents_test.py:45 (72x96x3, 9 actions) and SURVEY 8(a4)."""
  assert len(net_oracle.param_specs('deep', 9, (72, 96, 3))) == 39
  specs = net_oracle.param_specs('deep', 18, (84, 84, 4))
  assert len(specs) == 39
  assert sum(int(np.prod(s)) for _, s in specs) == 1638883


def _drive(store, rows, batch):
  out = []
  for i in range(0, len(rows) - len(rows) % batch, batch):
    chunk = rows[i:i + batch]
    rs = np.array([c[0] for c in chunk]); ids = np.array([c[1] for c in chunk], np.int32)
    vals = np.array([c[2] for c in chunk], np.int32)
    store.reset(ids[rs])
    cid, un = store.append(ids, [vals])
    out.append((cid.tolist(), un[0].tolist()))
  return out


FULL_ROWS = [(False, 0, 10), (False, 2, 30), (False, 1, 20), (False, 0, 11), (False, 2, 31),
             (False, 3, 40), (False, 0, 12), (False, 2, 32), (False, 3, 41), (False, 0, 13),
             (False, 1, 21), (True, 2, 33), (False, 0, 14), (False, 2, 34), (False, 3, 42),
             (False, 0, 15), (False, 1, 22), (False, 2, 35), (False, 0, 16), (False, 1, 23),
             (False, 2, 36)]
FULL_EXPECT = [([], []), ([], []), ([], []), ([0], [[10, 11, 12, 13]]), ([], []), ([], []),
               ([0, 1, 2], [[13, 14, 15, 16], [20, 21, 22, 23], [33, 34, 35, 36]])]
OVERLAP_ROWS = [(False, 0, 10), (False, 1, 20), (False, 0, 11), (False, 1, 21), (False, 0, 12),
                (True, 1, 22), (False, 0, 13), (False, 1, 23), (False, 0, 14), (False, 1, 24),
                (True, 0, 15), (False, 1, 25), (False, 0, 16), (False, 1, 26), (False, 0, 17),
                (False, 1, 27)]
OVERLAP_EXPECT = [([], []), ([], []), ([0], [[0, 0, 10, 11, 12]]), ([], []),
                  ([0, 1], [[10, 11, 12, 13, 14], [0, 0, 22, 23, 24]]), ([], []),
                  ([1], [[22, 23, 24, 25, 26]]), ([0], [[0, 0, 15, 16, 17]])]


def test_unroll_store_oracle_full_sequence():
  """reference tests/utils_test.py:80-160."""
  s = store_oracle.UnrollStore(4, 3, [((), np.int32)])
  assert _drive(s, FULL_ROWS, 3) == FULL_EXPECT


def test_unroll_store_oracle_overlap_2():
  """reference tests/utils_test.py:191-271."""
  s = store_oracle.UnrollStore(2, 2, [((), np.int32)], num_overlapping_steps=2)
  assert _drive(s, OVERLAP_ROWS, 2) == OVERLAP_EXPECT


def test_unroll_store_oracle_duplicates():
  s = store_oracle.UnrollStore(2, 3, [((), np.int32)])
  with pytest.raises(ValueError):
    s.append(np.array([1, 1]), [np.array([42, 43])])


def test_aggregator_oracle():
  """reference tests/utils_test.py:276-286."""
  a = store_oracle.Aggregator(4, [((), np.int32)])
  assert a.read([0, 1, 2, 3])[0].tolist() == [0, 0, 0, 0]
  a.add([0, 1], [np.array([42, 43])])
  assert a.read([0, 1, 2, 3])[0].tolist() == [42, 43, 0, 0]
  a.reset([0])
  assert a.read([0, 1, 2, 3])[0].tolist() == [0, 43, 0, 0]
  a.replace([0, 2], [np.array([1, 2])])
  assert a.read([0, 1, 2, 3])[0].tolist() == [1, 43, 2, 0]


@pytest.mark.parametrize('case', ['a', 'b', 'c'])
def test_compute_loss_composition_against_reference_source(case):
  """oracle/loss_oracle.py against tests/golden/loss_golden.npz = the UNMODIFIED reference
  compute_loss (agents/vtrace/learner.py:73-159) and common/vtrace.py executed over the numpy
  shim (tests/golden/make_golden_loss.py): total loss and every logged scalar, by name."""
  import os
  from oracle import loss_oracle
  g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'loss_golden.npz'))
  d, lam, bc, ec, kc, mar, te = (float(x) for x in g[case + '_cfg'])
  cfg = loss_oracle.default_config(discounting=d, lambda_=lam, baseline_cost=bc, entropy_cost=ec, kl_cost=kc,
                                   max_abs_reward=mar, target_entropy=te or None)
  total, logs, _ = loss_oracle.compute_loss_from_outputs(
      cfg, g[case + '_ll'], g[case + '_lb'], g[case + '_bl'], g[case + '_act'], g[case + '_rew'], g[case + '_done'])
  np.testing.assert_allclose(float(total), float(g[case + '_total']), rtol=2e-5, atol=2e-6)
  keys = [k[len(case) + 5:].replace('__', '/') for k in g.files if k.startswith(case + '_log_')]
  assert sorted(keys) == sorted(logs)                     # same scalar names as the reference logs
  for k in keys:
    np.testing.assert_allclose(float(logs[k]), float(g['%s_log_%s' % (case, k.replace('/', '__'))]),
                               rtol=2e-5, atol=2e-6, err_msg=k)


def test_impala_deep_wiring_against_reference_source():
  """oracle/net_oracle.py against tests/golden/net_golden.npz = the UNMODIFIED reference
  dmlab/networks.py (_Stack, ImpalaDeep) + common/utils.batch_apply executed over a Keras-layer
  shim (tests/golden/make_golden_net.py): same weights, same inputs -> same logits, baseline,
  final LSTM state, for a 5-step unroll with done-resets and for a single inference step."""
  import os, sys
  import torch
  from oracle import net_oracle
  sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
  import net_golden_params as G
  g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'net_golden.npz'))
  p = G.make_params(); i = G.make_inputs()
  t = torch.as_tensor
  with torch.no_grad():
    logits, baseline, (h, c) = net_oracle.unroll('deep', p, t(i['prev']), t(i['rew']), t(i['done']), t(i['obs']),
                                                 (t(i['h0']), t(i['c0'])), G.A)
    l1, b1, (h1, _) = net_oracle.unroll('deep', p, t(i['prev'][:1]), t(i['rew'][:1]), t(i['done'][:1]), t(i['obs'][:1]),
                                         (t(i['h0']), t(i['c0'])), G.A)
  for got, want in ((logits, 'logits'), (baseline, 'baseline'), (h, 'h'), (c, 'c'), (l1[0], 'logits1'),
                    (b1[0], 'baseline1'), (h1, 'h1')):
    np.testing.assert_allclose(got.numpy(), g[want], rtol=1e-5, atol=1e-6, err_msg=want)
  assert i['done'].any() and not i['done'].all()          # the reset path is exercised
