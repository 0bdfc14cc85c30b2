"""One configuration of the fused V-trace loss kernel (for ncu): T1 B A [stream]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from seed_rl_b200 import _lib
from seed_rl_b200.agents.vtrace import learner
T1, Bs, A = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
_lib.lib().seedrl_debug_set_loss_stream(int(sys.argv[4]) if len(sys.argv) > 4 else 1)
st = learner.default_loss_settings()
ecp = torch.tensor(np.log(st.entropy_cost) / 10.0, dtype=torch.float32).cuda()
g = torch.Generator(device='cuda').manual_seed(0)
ll = torch.randn(T1, Bs, A, device='cuda', generator=g); lb = torch.randn(T1, Bs, device='cuda', generator=g)
bl = torch.randn(T1, Bs, A, device='cuda', generator=g)
act = torch.randint(0, A, (T1, Bs), device='cuda', generator=g)
rew = torch.randn(T1, Bs, device='cuda', generator=g); dn = torch.rand(T1, Bs, device='cuda', generator=g) < 0.02
for _ in range(3):
  learner.vtrace_loss_fwd_bwd(st, ll, lb, bl, act, rew, dn, ecp)
torch.cuda.synchronize()
