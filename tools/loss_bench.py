import sys, torch, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seed_rl_b200 import _lib
from seed_rl_b200.agents.vtrace import learner
st = learner.default_loss_settings()
ecp = torch.tensor(np.log(st.entropy_cost) / 10.0, dtype=torch.float32).cuda()
for T1, A in ((21, 18), (101, 18)):
  for Bs in (64, 4096, 65536):
    if T1 == 101 and Bs == 65536: Bs = 16384
    g = torch.Generator(device='cuda').manual_seed(0)
    ll = torch.randn(T1, Bs, A, device='cuda', generator=g); lb = torch.randn(T1, Bs, device='cuda', generator=g)
    bl = torch.randn(T1, Bs, A, device='cuda', generator=g)
    act = torch.randint(0, A, (T1, Bs), device='cuda', generator=g)
    rew = torch.randn(T1, Bs, device='cuda', generator=g); dn = torch.rand(T1, Bs, device='cuda', generator=g) < 0.02
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for stream in (0, 1, 4, 8):
      _lib.lib().seedrl_debug_set_loss_stream(stream)
      for _ in range(3):
        learner.vtrace_loss_fwd_bwd(st, ll, lb, bl, act, rew, dn, ecp)
      times = []
      for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); learner.vtrace_loss_fwd_bwd(st, ll, lb, bl, act, rew, dn, ecp); e1.record()
        torch.cuda.synchronize(); times.append(e0.elapsed_time(e1))
      times.sort(); ms = times[5]
      # kernel alone: 20 back-to-back launches (working set >> L2 at the large sizes)
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      torch.cuda.synchronize(); e0.record()
      for _ in range(20):
        learner.vtrace_loss_fwd_bwd(st, ll, lb, bl, act, rew, dn, ecp)
      e1.record(); torch.cuda.synchronize()
      msk = e0.elapsed_time(e1) / 20
      nb = (161 + 76) * (T1 - 1) * Bs + 4 * Bs + 32
      print('T1=%d B=%d stream=%d  single %.4f ms %.0f GB/s | back-to-back %.4f ms %.0f GB/s' % (T1, Bs, stream, ms, nb / ms / 1e6, msk, nb / msk / 1e6), flush=True)
