"""Checkpoint / resume (reference agents/vtrace/learner.py:283-296,470-476: tf.train.Checkpoint of
agent + optimizer, restored with assert_consumed) on the flat-arena format of this framework:
train 2 steps, save, train 2 more; a fresh agent restored from the file and trained on the same 2
batches must land on bit-identical parameters and Adam slots (kernels are deterministic)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(seed):
  from seed_rl_b200.agents.vtrace import learner
  from seed_rl_b200.common import optimizers
  from seed_rl_b200.dmlab import networks
  agent = networks.ImpalaDeep(18, (84, 84, 4), seed=seed, conv_mode='tc3p')
  opt = optimizers.Adam(optimizers.PolynomialDecay(4.8e-4, 1000, 0.0), beta_1=0.0, epsilon=3.125e-7)
  return agent, opt, learner.LearnerStep(agent, opt, settings=learner.default_loss_settings())


def test_checkpoint_roundtrip_resumes_bit_identically(tmp_path):
  from oracle import learner_oracle
  from seed_rl_b200.agents.vtrace import learner_loop
  from test_gpu_parity import _batch_to_cuda
  batches = [_batch_to_cuda(learner_oracle.synthetic_batch(4, 3, 18, seed=40 + i)) for i in range(4)]
  agent, opt, step = _mk(seed=3)
  for b in batches[:2]:
    step.minimize(b)
  path = str(tmp_path / 'ckpt.pt')
  learner_loop.save_checkpoint(path, agent, opt)
  for b in batches[2:]:
    step.minimize(b)
  agent2, opt2, step2 = _mk(seed=99)                  # different init: everything must come from the file
  learner_loop.restore_checkpoint(path, agent2, opt2)
  assert opt2.iterations == 2
  for b in batches[2:]:
    step2.minimize(b)
  torch.cuda.synchronize()
  assert torch.equal(agent.params, agent2.params)
  assert torch.equal(opt.m, opt2.m) and torch.equal(opt.v, opt2.v)
  assert opt.iterations == opt2.iterations == 4
  # a checkpoint of another network is refused (assert_consumed analogue)
  from seed_rl_b200.common import optimizers
  from seed_rl_b200.dmlab import networks
  other = networks.ImpalaShallow(18, (84, 84, 4), conv_mode='tc3')
  with pytest.raises(ValueError, match='tensor table mismatch'):
    learner_loop.restore_checkpoint(path, other, optimizers.Adam(1e-3))
