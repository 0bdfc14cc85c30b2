"""CPU ORACLE (test infrastructure, NOT product code) -- one whole learner step on
the CPU: the `minimize` of /root/reference/agents/vtrace/learner.py:255-280
= compute_loss (:73-159: agent unroll -> log-probs -> V-trace -> losses)
-> tape.gradient (:264) -> Keras Adam apply (:272-273), composed from
net_oracle / loss_oracle / optim_oracle.  Used as the parity reference for the
CUDA learner step and, timed, as bench.py's `cpu_baseline` / `--impl reference`
arm (kind "port": TensorFlow is not installable here).
"""
import collections

import numpy as np
import torch

from . import loss_oracle, net_oracle, optim_oracle


def synthetic_batch(T, B, A, obs_shape=(84, 84, 4), seed=1234):
  """SURVEY 8(d) synthetic unroll batch, time-major with T+1 rows."""
  rng = np.random.default_rng(seed)
  T1 = T + 1
  return dict(
      observation=rng.integers(0, 256, (T1, B) + tuple(obs_shape), dtype=np.uint8),
      reward=rng.normal(size=(T1, B)).astype(np.float32),
      done=rng.random((T1, B)) < 0.02,
      prev_actions=rng.integers(0, A, (T1, B), dtype=np.int64),
      action=rng.integers(0, A, (T1, B), dtype=np.int64),
      behaviour_logits=rng.normal(size=(T1, B, A)).astype(np.float32),
      behaviour_baseline=rng.normal(size=(T1, B)).astype(np.float32),
      h0=np.zeros((B, net_oracle.LSTM_UNITS), np.float32),
      c0=np.zeros((B, net_oracle.LSTM_UNITS), np.float32))


def forward_loss(net, params_t, batch, cfg, num_actions, entropy_cost_param=None):
  """compute_loss (learner.py:73-159).  params_t: dict of torch tensors."""
  logits, baseline, state = net_oracle.unroll(
      net, params_t, torch.as_tensor(batch['prev_actions']),
      torch.as_tensor(batch['reward']), torch.as_tensor(batch['done']),
      torch.as_tensor(batch['observation']),
      (torch.as_tensor(batch['h0']), torch.as_tensor(batch['c0'])), num_actions)
  total, logs, aux = loss_oracle.compute_loss_from_outputs(
      cfg, logits, baseline, batch['behaviour_logits'], batch['action'],
      batch['reward'], batch['done'], entropy_cost_param)
  return total, logs, dict(aux, logits=logits, baseline=baseline, state=state)


class CpuLearner(object):
  """Holds params + Adam slots; step() = one `minimize`."""

  def __init__(self, net, num_actions, obs_shape, cfg, lr=0.00048, beta1=0.0,
               beta2=0.999, eps=3.125e-7, decay_steps=None, params=None, seed=0):
    self.net, self.A, self.cfg = net, num_actions, cfg
    init = params if params is not None else net_oracle.init_params(
        net, num_actions, obs_shape, seed)
    self.params = net_oracle.to_torch(init, requires_grad=True)
    mul = cfg.entropy_cost_adjustment_speed
    self.entropy_cost_param = torch.tensor(
        np.log(cfg.entropy_cost) / mul, dtype=torch.float32, requires_grad=True)
    self.m = {k: np.zeros(tuple(v.shape), np.float32) for k, v in self.params.items()}
    self.v = {k: np.zeros(tuple(v.shape), np.float32) for k, v in self.params.items()}
    self.m['entropy_cost_param'] = np.zeros((), np.float32)
    self.v['entropy_cost_param'] = np.zeros((), np.float32)
    self.iterations = 0
    self.lr, self.b1, self.b2, self.eps, self.decay_steps = lr, beta1, beta2, eps, decay_steps

  def grads(self, batch):
    for t in list(self.params.values()) + [self.entropy_cost_param]:
      t.grad = None
    total, logs, aux = forward_loss(self.net, self.params, batch, self.cfg, self.A,
                                    self.entropy_cost_param)
    total.backward()
    g = collections.OrderedDict((k, v.grad.numpy().copy()) for k, v in self.params.items())
    g['entropy_cost_param'] = (self.entropy_cost_param.grad.numpy().copy()
                               if self.entropy_cost_param.grad is not None
                               else np.zeros((), np.float32))
    return total.detach(), logs, g, aux

  def step(self, batch, grad_scale=1.0):
    total, logs, g, _ = self.grads(batch)
    lr = self.lr if self.decay_steps is None else optim_oracle.polynomial_decay(
        self.lr, self.iterations, self.decay_steps)
    with torch.no_grad():
      for k in g:
        p = self.entropy_cost_param if k == 'entropy_cost_param' else self.params[k]
        p2, self.m[k], self.v[k] = optim_oracle.keras_adam_step(
            p.detach().numpy(), g[k] * np.float32(grad_scale), self.m[k], self.v[k],
            self.iterations, lr, self.b1, self.b2, self.eps)
        if k == 'entropy_cost_param':       # constraint, learner.py:231
          mul = self.cfg.entropy_cost_adjustment_speed
          p2 = np.clip(p2, -20.0 / mul, 20.0 / mul).astype(np.float32)
        p.copy_(torch.from_numpy(np.asarray(p2)))
    self.iterations += 1
    return float(total), logs
