#!/usr/bin/env python
"""Times the cfg-5 R2D2 learner step (BASELINE configs[4]) at the reference's default shapes:
batch 64 sampled unrolls of burn_in 40 + unroll_length 100 + 1 steps, 84x84x1 frames, stack 4."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seed_rl_b200 import _lib
from seed_rl_b200.agents.r2d2 import learner
from seed_rl_b200.atari import networks
from seed_rl_b200.common import optimizers, utils

A, obs, S = 18, (84, 84, 1), 4
B = int(os.environ.get('B', 64)); burn = 40; T = burn + 100 + 1
mode = os.environ.get('MODE', 'tc3')
g = torch.Generator(device='cuda').manual_seed(0)
frames = torch.randint(0, 256, (T, B) + obs, dtype=torch.uint8, device='cuda', generator=g)
env = utils.EnvOutput(torch.randn(T, B, device='cuda', generator=g), torch.rand(T, B, device='cuda', generator=g) < 0.01,
                      frames, torch.zeros(T, B, dtype=torch.bool, device='cuda'), torch.zeros(T, B, dtype=torch.int32, device='cuda'))
agent = networks.DuelingLSTMDQNNet(A, obs, S, seed=0, gemm_mode=mode)
target = networks.DuelingLSTMDQNNet(A, obs, S, seed=0, gemm_mode=mode)
state = agent.initial_state(B)
pa = torch.randint(0, A, (T, B), device='cuda', generator=g)
unrolls = learner.Unroll(state, None, pa, env, learner.AgentOutput(torch.randint(0, A, (T, B), device='cuda', generator=g), None))
sampled = learner.SampledUnrolls(unrolls, torch.arange(B, device='cuda'), torch.rand(B, device='cuda', generator=g))
step = learner.R2D2LearnerStep(agent, target, optimizers.Adam(0.00048, epsilon=1e-3), settings=learner.default_settings())
for _ in range(2):
  loss, prio, idx, norm = step.minimize(sampled)
torch.cuda.synchronize(); agent.check_errors()
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
  loss, prio, idx, norm = step.minimize(sampled)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
L = _lib.lib()
import ctypes
ncat = L.seedrl_profile_num_categories()
ms_c = (ctypes.c_double * ncat)(); n_c = (ctypes.c_uint64 * ncat)()
_lib.check(L.seedrl_profile_begin(_lib.stream_ptr()))
step.minimize(sampled)
_lib.check(L.seedrl_profile_end(ms_c, n_c))
cats = {L.seedrl_profile_category_name(i).decode(): round(ms_c[i], 3) for i in range(ncat)}
print(json.dumps(dict(mode=mode, B=B, T=T, ms_per_step=ms, env_frames_per_s=B * 100 / (ms * 1e-3),
                      launches_per_step=(_lib.launch_count() - n0) // (K + 1), loss=float(loss), grad_norm=float(norm),
                      mem_gb=torch.cuda.max_memory_allocated() / 2**30, kernel_ms=cats)))
