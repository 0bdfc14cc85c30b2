"""CPU ORACLE (test infrastructure, NOT product code) -- torch-CPU fp32 restatement of
/root/reference/agents/vtrace/learner.py:82-157 (the part of `compute_loss`
after the network unroll) with autograd standing in for tf.GradientTape
(learner.py:261-264).

Parity status: the reference has NO test of compute_loss for V-trace.  The
FORWARD composition (row slicing, reward clip, discounts, the five loss terms and the 11
logged scalars with their names) is pinned against the UNMODIFIED reference compute_loss
executed over tests/golden/tf_numpy_shim.py (tests/golden/make_golden_loss.py ->
loss_golden.npz, tests/test_oracle_golden.py); its V-trace core is pinned separately
(vtrace_oracle.py) and the categorical log_prob by reference tests/vtrace_test.py:88-115.
The gradient (tf.GradientTape) stays "parity unpinned": torch autograd of this forward.
"""
import collections

import numpy as np
import torch

from . import vtrace_oracle

LossConfig = collections.namedtuple(
    'LossConfig',
    'discounting lambda_ baseline_cost entropy_cost kl_cost max_abs_reward '
    'target_entropy entropy_cost_adjustment_speed')


def default_config(**kw):
  """Flag defaults of learner.py:51-62."""
  d = dict(discounting=0.99, lambda_=1.0, baseline_cost=0.5,
           entropy_cost=0.00025, kl_cost=0.0, max_abs_reward=0.0,
           target_entropy=None, entropy_cost_adjustment_speed=10.0)
  d.update(kw)
  return LossConfig(**d)


def compute_loss_from_outputs(cfg, learner_logits, learner_baseline,
                              behaviour_logits, actions, rewards, done,
                              entropy_cost_param=None):
  """All inputs carry T+1 rows (time-major [T+1, B, ...]) exactly as they
  arrive in compute_loss.  `learner_logits`/`learner_baseline` may require
  grad.  Returns (total_loss, logs dict of python floats/tensors)."""
  f32 = torch.float32
  learner_logits = torch.as_tensor(learner_logits, dtype=f32)
  learner_baseline = torch.as_tensor(learner_baseline, dtype=f32)
  behaviour_logits = torch.as_tensor(behaviour_logits, dtype=f32)
  actions = torch.as_tensor(actions).long()
  rewards = torch.as_tensor(rewards, dtype=f32)
  done = torch.as_tensor(done).bool()

  bootstrap_value = learner_baseline[-1]                       # :82
  a = actions[:-1]                                             # :86
  beh_logits = behaviour_logits[:-1]
  rewards = rewards[1:]                                        # :87
  done = done[1:]
  tgt_logits = learner_logits[:-1]                             # :88
  values = learner_baseline[:-1]
  if cfg.max_abs_reward:                                       # :90-92
    rewards = torch.clamp(rewards, -cfg.max_abs_reward, cfg.max_abs_reward)
  discounts = (~done).to(f32) * cfg.discounting                # :93

  tgt_lsm = torch.log_softmax(tgt_logits, -1)
  beh_lsm = torch.log_softmax(beh_logits, -1)
  tgt_logp = tgt_lsm.gather(-1, a[..., None])[..., 0]          # :95-96
  beh_logp = beh_lsm.gather(-1, a[..., None])[..., 0]          # :97-98

  vt = vtrace_oracle.from_importance_weights(                  # :101-108
      tgt_logp.detach().numpy(), beh_logp.detach().numpy(),
      discounts.numpy(), rewards.numpy(), values.detach().numpy(),
      bootstrap_value.detach().numpy(), lambda_=cfg.lambda_)
  vs = torch.from_numpy(vt.vs)
  pg_adv = torch.from_numpy(vt.pg_advantages)

  policy_loss = -torch.mean(tgt_logp * pg_adv)                 # :111-112
  v_error = vs - values                                        # :115
  v_loss = cfg.baseline_cost * 0.5 * torch.mean(v_error ** 2)  # :116
  entropy = torch.mean(-(tgt_lsm.exp() * tgt_lsm).sum(-1))     # :119-120
  mul = cfg.entropy_cost_adjustment_speed
  if entropy_cost_param is None:
    entropy_cost_param = torch.tensor(np.log(cfg.entropy_cost) / mul, dtype=f32)
  entropy_cost = torch.exp(mul * entropy_cost_param)           # :225-234
  entropy_loss = entropy_cost.detach() * -entropy              # :121
  kl = beh_logp - tgt_logp                                     # :124
  kl_loss = cfg.kl_cost * torch.mean(kl)                       # :125
  if cfg.target_entropy:                                       # :128-132
    adj = entropy_cost * (entropy.detach() - cfg.target_entropy)
  else:
    adj = 0. * entropy_cost
  total = policy_loss + v_loss + entropy_loss + kl_loss + adj  # :134-135
  logs = collections.OrderedDict([                             # :138-157
      ('V/value function', values.mean()),
      ('V/L2 error', torch.sqrt(torch.mean(v_error ** 2))),
      ('losses/policy', policy_loss),
      ('losses/V', v_loss),
      ('losses/entropy', entropy_loss),
      ('losses/kl', kl_loss),
      ('losses/total', total),
      ('policy/max_action_abs(before_tanh)', a.abs().max()),
      ('policy/entropy', entropy),
      ('policy/entropy_cost', entropy_cost),
      ('policy/kl(old|new)', kl.mean()),
  ])
  aux = dict(vs=vs, pg_advantages=pg_adv, tgt_logp=tgt_logp, beh_logp=beh_logp)
  return total, logs, aux


def loss_and_grads(cfg, learner_logits, learner_baseline, behaviour_logits,
                   actions, rewards, done, entropy_cost_param=None):
  """Returns (total_loss, logs, dlogits [T+1,B,A], dbaseline [T+1,B],
  d_entropy_cost_param)."""
  ll = torch.tensor(np.asarray(learner_logits), dtype=torch.float32,
                    requires_grad=True)
  lb = torch.tensor(np.asarray(learner_baseline), dtype=torch.float32,
                    requires_grad=True)
  mul = cfg.entropy_cost_adjustment_speed
  if entropy_cost_param is None:
    entropy_cost_param = np.log(cfg.entropy_cost) / mul
  ep = torch.tensor(float(entropy_cost_param), dtype=torch.float32,
                    requires_grad=True)
  total, logs, aux = compute_loss_from_outputs(
      cfg, ll, lb, behaviour_logits, actions, rewards, done, ep)
  total.backward()
  dep = ep.grad if ep.grad is not None else torch.zeros(())
  return (total.detach(), {k: float(v) for k, v in logs.items()},
          ll.grad.numpy(), lb.grad.numpy(), float(dep), aux)
