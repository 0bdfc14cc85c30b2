#!/usr/bin/env python
"""One shape of the plane-tensor conv path, a few launches (for ncu): planes_one.py conv|wgrad CIN COUT H [N]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seed_rl_b200 import _lib
L = _lib.lib()
kind, ci, co, H = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
N = int(sys.argv[5]) if len(sys.argv) > 5 else 1344
def planes(C, relu=0):
  x = torch.randn(N, H, H, C, device='cuda')
  p = torch.empty(int(L.seedrl_debug_planes_bytes(N, H, H, C)), dtype=torch.uint8, device='cuda')
  _lib.check(L.seedrl_debug_to_planes(N, H, H, C, relu, _lib.ptr(x), _lib.ptr(p), _lib.stream_ptr()))
  return p
err = torch.zeros(1, dtype=torch.int32, device='cuda')
xin = planes(ci, 1)
if kind == 'conv':
  w = torch.randn(3, 3, ci, co, device='cuda') * 0.1; b = torch.zeros(co, device='cuda')
  wq = torch.empty(2 * 9 * ci * co * 2, dtype=torch.uint8, device='cuda')
  o = torch.empty(int(L.seedrl_debug_planes_bytes(N, H, H, co)), dtype=torch.uint8, device='cuda')
  for _ in range(4):
    _lib.check(L.seedrl_debug_convp(ci, co, N, H, H, _lib.ptr(xin), _lib.ptr(w), _lib.ptr(b), None, None, 0, None,
                                    _lib.ptr(o), None, _lib.ptr(wq), _lib.ptr(err), _lib.stream_ptr()))
else:
  dy = planes(co)
  dw = torch.empty(3, 3, ci, co, device='cuda'); db = torch.empty(co, device='cuda')
  part = torch.empty(148 * (9 * ci * co + co), device='cuda')
  for _ in range(4):
    _lib.check(L.seedrl_debug_wgradp(ci, co, N, H, H, _lib.ptr(xin), _lib.ptr(dy), _lib.ptr(dw), _lib.ptr(db),
                                     _lib.ptr(part), part.numel() * 4, _lib.ptr(err), _lib.stream_ptr()))
torch.cuda.synchronize()
assert int(err.item()) == 0
