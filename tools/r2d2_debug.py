import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from oracle import r2d2_learner_oracle as RL, r2d2_net_oracle as NO
from seed_rl_b200.atari import networks
from test_gpu_r2d2 import _net_case, _to_cuda_inputs
for (T, B, A, obs, S, mode) in [(3, 2, 18, (84, 84, 1), 4, 'simt'), (5, 3, 6, (36, 36, 1), 4, 'tc3'), (5, 3, 6, (36, 36, 1), 4, 'simt')]:
  params, b = _net_case(T, B, A, obs, S, seed=T + A)
  agent = networks.DuelingLSTMDQNNet(A, obs, S, gemm_mode=mode); agent.load_named_parameters(params)
  pa, env, state = _to_cuda_inputs(b, S)
  out, _ = agent((pa, env), state, unroll=True, is_training=True)
  pt = {k: torch.tensor(v, requires_grad=True) for k, v in params.items()}
  want, _ = NO.unroll(pt, b['prev_actions'], b['reward'], b['done'], b['observation'],
                      NO.AgentState((torch.as_tensor(b['h0']), torch.as_tensor(b['c0'])), b['frame_state']), A, S)
  dq = np.random.default_rng(3).normal(size=tuple(want.q_values.shape)).astype(np.float32)
  (want.q_values * torch.as_tensor(dq)).sum().backward()
  agent.backward(torch.as_tensor(dq).cuda())
  mine = agent.named_gradients()
  print('CASE', T, B, A, obs, S, mode, 'fwd err', float(np.abs(out.q_values.cpu().numpy() - want.q_values.detach().numpy()).max()))
  for k, v in pt.items():
    w = v.grad.numpy(); a = mine[k].cpu().numpy()
    print('  %-28s err %.3e  maxabs %.3e' % (k, np.abs(a - w).max() / (np.abs(w).max() + 1e-30), np.abs(w).max()))
  w = pt['body/conv0/kernel'].grad.numpy(); a = mine['body/conv0/kernel'].cpu().numpy()
  e = np.abs(a - w) / np.abs(w).max()
  print('  conv0 err by kh', e.max(axis=(1, 2, 3)).round(5)); print('  by kw', e.max(axis=(0, 2, 3)).round(5))
  print('  by ci', e.max(axis=(0, 1, 3)).round(5)); print('  by co', e.max(axis=(0, 1, 2)).round(5))
  bw = pt['body/conv0/bias'].grad.numpy(); ba = mine['body/conv0/bias'].cpu().numpy()
  print('  bias oracle', bw[:8].round(5), 'mine', ba[:8].round(5))
