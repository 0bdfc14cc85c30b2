"""CPU ORACLE (test infrastructure, NOT product code) -- numpy fp32 restatement of

    /root/reference/common/vtrace.py:34-148        from_importance_weights
    /root/reference/common/parametric_distribution.py:66-74,94-95
                                                   categorical log_prob / entropy
                                                   (arithmetic = TFP 0.11.0
                                                   tfd.Categorical, not vendored)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; the product path
(seed_rl_b200/) never does and fails loudly without its CUDA library.

Pinning: checked by tests/test_oracle_golden.py against
  * the known-answer case of reference tests/vtrace_test.py:118-145, evaluated
    with the reference test's own O(T^2) numpy ground truth (:41-82);
  * the reference's common/vtrace.py source itself, executed over a numpy
    stand-in for the ~10 TF ops it uses (tests/golden/make_golden.py);
  * the second, independent reference implementation
    agents/policy_gradient/modules/advantages.py:28-108 (lambda=0.95 case of
    advantages_test.py:129-150).
"""
import collections

import numpy as np

VTraceReturns = collections.namedtuple('VTraceReturns', 'vs pg_advantages')

_f32 = np.float32


def from_importance_weights(target_action_log_probs, behaviour_action_log_probs,
                            discounts, rewards, values, bootstrap_value,
                            clip_rho_threshold=1.0, clip_pg_rho_threshold=1.0,
                            lambda_=1.0):
  """vtrace.py:84-148. All tensors [T, B, ...] (bootstrap [B, ...]), fp32."""
  log_rhos = (np.asarray(target_action_log_probs, _f32) -
              np.asarray(behaviour_action_log_probs, _f32))        # :84
  discounts = np.asarray(discounts, _f32)
  rewards = np.asarray(rewards, _f32)
  values = np.asarray(values, _f32)
  bootstrap_value = np.asarray(bootstrap_value, _f32)
  # rank checks, :99-107
  if values.ndim != log_rhos.ndim or discounts.ndim != log_rhos.ndim or \
     rewards.ndim != log_rhos.ndim or bootstrap_value.ndim != log_rhos.ndim - 1:
    raise ValueError('inconsistent ranks')

  rhos = np.exp(log_rhos)                                          # :110
  if clip_rho_threshold is not None:                               # :111-114
    clipped_rhos = np.minimum(_f32(clip_rho_threshold), rhos)
  else:
    clipped_rhos = rhos
  cs = np.minimum(_f32(1.0), rhos) * _f32(lambda_)                 # :116-117
  values_t_plus_1 = np.concatenate(
      [values[1:], bootstrap_value[None]], axis=0)                 # :120-121
  deltas = clipped_rhos * (rewards + discounts * values_t_plus_1 - values)  # :122

  acc = np.zeros_like(bootstrap_value)                             # :124
  out = [None] * discounts.shape[0]
  for i in range(discounts.shape[0] - 1, -1, -1):                  # :126-129
    acc = deltas[i] + discounts[i] * cs[i] * acc
    out[i] = acc
  vs_minus_v_xs = np.stack(out, axis=0) if out else np.zeros_like(values)
  vs = vs_minus_v_xs + values                                      # :133
  vs_t_plus_1 = np.concatenate([vs[1:], bootstrap_value[None]], axis=0)  # :136-137
  if clip_pg_rho_threshold is not None:                            # :138-142
    clipped_pg_rhos = np.minimum(_f32(clip_pg_rho_threshold), rhos)
  else:
    clipped_pg_rhos = rhos
  pg_advantages = clipped_pg_rhos * (
      rewards + discounts * vs_t_plus_1 - values)                  # :143-144
  return VTraceReturns(vs=vs.astype(_f32), pg_advantages=pg_advantages.astype(_f32))


def log_softmax(logits):
  logits = np.asarray(logits, _f32)
  m = logits.max(axis=-1, keepdims=True)
  z = logits - m
  return z - np.log(np.exp(z).sum(axis=-1, keepdims=True, dtype=_f32))


def categorical_log_prob(logits, actions):
  """parametric_distribution.py:69-70 -> tfd.Categorical(logits).log_prob(a)
  = log_softmax(logits)[a]  (pinned by reference tests/vtrace_test.py:88-115)."""
  lsm = log_softmax(logits)
  a = np.asarray(actions).astype(np.int64)
  return np.take_along_axis(lsm, a[..., None], axis=-1)[..., 0]


def categorical_entropy(logits):
  """parametric_distribution.py:72-74 -> tfd.Categorical.entropy()
  = -sum softmax * log_softmax."""
  lsm = log_softmax(logits)
  return -(np.exp(lsm) * lsm).sum(axis=-1, dtype=_f32)


def categorical_sample_from_noise(logits, gumbel_noise):
  """dmlab/networks.py:121 tf.random.categorical == Gumbel-max: argmax_k(logit_k
  + g_k). TF's Philox stream is not reproducible, so parity is defined on
  INJECTED noise g (fp32): one fp32 add + first-max argmax is bit-exact."""
  s = np.asarray(logits, _f32) + np.asarray(gumbel_noise, _f32)
  return np.argmax(s, axis=-1).astype(np.int64)
