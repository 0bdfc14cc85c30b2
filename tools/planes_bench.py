#!/usr/bin/env python
"""Single-kernel timings of the plane-tensor conv path (csrc/conv_planes.cu) at the learner's
layer shapes: python tools/planes_bench.py  [SEEDRL_PLANES_CHUNK=16|32|64|128 in the env]."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seed_rl_b200 import _lib
L = _lib.lib()
N = int(os.environ.get('FRAMES', 1344))
peak = 6561.6

def ev(fn, k=10):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(k): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / k

def planes(H, C, relu=0):
  x = torch.randn(N, H, H, C, device='cuda')
  nb = int(L.seedrl_debug_planes_bytes(N, H, H, C))
  p = torch.empty(nb, dtype=torch.uint8, device='cuda')
  _lib.check(L.seedrl_debug_to_planes(N, H, H, C, relu, _lib.ptr(x), _lib.ptr(p), _lib.stream_ptr()))
  return p

out = {'chunk': os.environ.get('SEEDRL_PLANES_CHUNK', 'default'), 'frames': N}
err = torch.zeros(1, dtype=torch.int32, device='cuda')
for (ci, co, H) in [(16, 16, 42), (32, 32, 21), (32, 32, 11), (16, 32, 42), (32, 16, 42)]:
  xin = planes(H, ci, 1)
  w = torch.randn(3, 3, ci, co, device='cuda') * 0.1
  b = torch.zeros(co, device='cuda')
  wq = torch.empty(2 * 9 * ci * co * 2, dtype=torch.uint8, device='cuda')
  o = torch.empty(int(L.seedrl_debug_planes_bytes(N, H, H, co)), dtype=torch.uint8, device='cuda')
  def conv():
    _lib.check(L.seedrl_debug_convp(ci, co, N, H, H, _lib.ptr(xin), _lib.ptr(w), _lib.ptr(b), None, None, 0, None,
                                    _lib.ptr(o), None, _lib.ptr(wq), _lib.ptr(err), _lib.stream_ptr()))
  ms = ev(conv)
  alg = N * H * H * (ci + co) * 4
  out['convp_%d_%d_%d' % (ci, co, H)] = dict(ms=round(ms, 4), GBps=round(alg / ms / 1e6, 1), frac=round(alg / ms / 1e6 / peak, 3))
  if (ci, co) in ((16, 16), (32, 32)):
    dy = planes(H, co)
    dw = torch.empty(3, 3, ci, co, device='cuda'); db = torch.empty(co, device='cuda')
    part = torch.empty(148 * (9 * ci * co + co), device='cuda')
    def wg():
      _lib.check(L.seedrl_debug_wgradp(ci, co, N, H, H, _lib.ptr(xin), _lib.ptr(dy), _lib.ptr(dw), _lib.ptr(db),
                                       _lib.ptr(part), part.numel() * 4, _lib.ptr(err), _lib.stream_ptr()))
    ms = ev(wg)
    out['wgradp_%d_%d_%d' % (ci, co, H)] = dict(ms=round(ms, 4), GBps=round(alg / ms / 1e6, 1), frac=round(alg / ms / 1e6 / peak, 3))
    del dy
  del xin, o
# pools
for (C, H) in [(16, 84), (32, 42)]:
  x = torch.randn(N, H, H, C, device='cuda'); Ho = (H + 1) // 2
  nbp = int(L.seedrl_debug_planes_bytes(N, Ho, Ho, C))
  raw = torch.empty(nbp, dtype=torch.uint8, device='cuda'); rel = torch.empty(nbp, dtype=torch.uint8, device='cuda')
  idx = torch.empty(N, Ho, Ho, C, dtype=torch.uint8, device='cuda')
  f = lambda: _lib.check(L.seedrl_debug_poolp(0, N, H, H, C, _lib.ptr(x), _lib.ptr(raw), _lib.ptr(rel), None, _lib.ptr(idx), _lib.stream_ptr()))
  ms = ev(f); alg = N * H * H * C * 4 + 2 * N * Ho * Ho * C * 4 + N * Ho * Ho * C
  out['poolp_fwd_%d_%d' % (C, H)] = dict(ms=round(ms, 4), frac=round(alg / ms / 1e6 / peak, 3))
  dxn = torch.empty(N, H, H, C, device='cuda')
  f = lambda: _lib.check(L.seedrl_debug_poolp(1, N, H, H, C, _lib.ptr(raw), None, None, _lib.ptr(dxn), _lib.ptr(idx), _lib.stream_ptr()))
  ms = ev(f); alg = N * H * H * C * 4 + N * Ho * Ho * C * 5
  out['poolp_bwd_nhwc_%d_%d' % (C, H)] = dict(ms=round(ms, 4), frac=round(alg / ms / 1e6 / peak, 3))
  del x, raw, rel, dxn
assert int(err.item()) == 0
print(json.dumps(out))
