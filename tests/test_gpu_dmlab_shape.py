"""ImpalaDeep on the reference's own DMLab observation shape 72x96x3 (dmlab/env.py:44-54): the
variable structure of tests/agents_test.py:45 (39 trainable tensors, first conv kernel [3,3,3,16])
and forward / gradient parity against the CPU oracle.  The kernels of the first convolution are
built for 4 input channels; 3-channel frames run on a zero-padded copy (csrc/net.cu PadScope)."""
import numpy as np
import pytest
import torch

from oracle import learner_oracle, loss_oracle, net_oracle

pytestmark = pytest.mark.gpu

OBS = (72, 96, 3)


def test_variable_structure_matches_agents_test():
  from seed_rl_b200.dmlab import networks
  agent = networks.ImpalaDeep(9, OBS)
  assert len(agent.trainable_variables) == 39                      # tests/agents_test.py:45
  shapes = {k: tuple(v.shape) for k, v in agent.named_parameters().items()}
  assert shapes['stack0/conv/kernel'] == (3, 3, 3, 16)
  assert shapes['conv_to_linear/kernel'] == (9 * 12 * 32, 256)
  want = net_oracle.init_params('deep', 9, OBS, seed=0)
  assert list(shapes) == list(want) and all(shapes[k] == tuple(v.shape) for k, v in want.items())


@pytest.mark.parametrize('mode,ftol,gtol', [('simt', 2e-5, 2e-3), ('tc3', 2e-4, 1e-2), ('tc3p', 2e-4, 1e-2)])
def test_learner_step_on_dmlab_frames_matches_oracle(mode, ftol, gtol):
  from seed_rl_b200.agents.vtrace import learner
  from seed_rl_b200.common import optimizers
  from seed_rl_b200.dmlab import networks
  from test_gpu_parity import _batch_to_cuda
  A, T, B = 9, 4, 3
  params = net_oracle.init_params('deep', A, OBS, seed=1)
  cpu = learner_oracle.CpuLearner('deep', A, OBS, loss_oracle.default_config(), params=params)
  b = learner_oracle.synthetic_batch(T, B, A, OBS, seed=100)
  total, _, g, aux = cpu.grads(b)
  agent = networks.ImpalaDeep(A, OBS, conv_mode=mode)
  agent.load_named_parameters(params)
  step = learner.LearnerStep(agent, optimizers.Adam(4.8e-4, beta_1=0.0, epsilon=3.125e-7))
  u = _batch_to_cuda(b)
  loss, _ = step.compute_gradients(u)
  agent.check_errors()
  assert abs(float(loss) - float(total)) < 2e-4 * max(1.0, abs(float(total)))
  out, _ = agent(u.prev_actions, u.env_outputs, u.agent_state, unroll=True)
  lg = aux['logits'].detach().numpy()
  assert np.abs(out.policy_logits.cpu().numpy() - lg).max() < ftol * max(1.0, np.abs(lg).max())
  mine = agent.named_gradients()
  errs = {}
  for k in g:
    if k != 'entropy_cost_param':
      a, w = mine[k].cpu().numpy().astype(np.float64), g[k].astype(np.float64)
      assert a.shape == w.shape, k
      errs[k] = float(np.linalg.norm(a - w) / (np.linalg.norm(w) + 1e-30))     # L2-relative, as test_gpu_zz_tc.py
  bad = {k: v for k, v in errs.items() if not v < gtol}
  print('DMLAB_SHAPE %s: max L2-rel grad err %.3g' % (mode, max(errs.values())))
  assert not bad, bad
