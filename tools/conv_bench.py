"""Per-layer timing of the 3x3 conv kernels at the learner's shapes (N = 21*64 = 1344 frames):
back-to-back launches between CUDA events; prints us and algorithmic GB/s.
  python tools/conv_bench.py [wgrad|fwd|all] [split] [wgrad chunk 128|256]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from seed_rl_b200 import _lib

L = _lib.lib()
what = sys.argv[1] if len(sys.argv) > 1 else 'all'
split = int(sys.argv[2]) if len(sys.argv) > 2 else 0
N = 1344
LAYERS = [(4, 16, 2, 84), (16, 16, 1, 42), (16, 32, 0, 42), (32, 32, 1, 21), (32, 32, 0, 21), (32, 32, 1, 11)]
REP = 10
if len(sys.argv) > 3:
  _lib.check(L.seedrl_debug_set_wgrad_chunk(int(sys.argv[3])))
  _lib.check(L.seedrl_debug_set_conv_tile(int(sys.argv[3])))


def timed(fn):
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(REP):
    fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / REP * 1e3


for cin, cout, mode, H in LAYERS:
  W = H
  if mode == 2:
    x = torch.randint(0, 256, (N, H, W, cin), dtype=torch.uint8, device='cuda')
  else:
    x = torch.randn(N, H, W, cin, device='cuda')
  dy = torch.randn(N, H, W, cout, device='cuda')
  xb = x.numel() * x.element_size(); yb = dy.numel() * 4
  err = torch.zeros(1, dtype=torch.int32, device='cuda')
  if what in ('wgrad', 'all'):
    pb = int(L.seedrl_debug_wgrad_partial_bytes())
    partial = torch.empty(pb // 4, device='cuda'); dw = torch.empty(3, 3, cin, cout, device='cuda'); db = torch.empty(cout, device='cuda')
    for name, fn in (('wgrad_tc', lambda: _lib.check(L.seedrl_debug_conv3x3_wgrad_tc(
        cin, cout, mode, split, N, H, W, _lib.ptr(x), _lib.ptr(dy), _lib.ptr(dw), _lib.ptr(db), _lib.ptr(partial), pb,
        _lib.ptr(err), _lib.stream_ptr()))),
                     ('wgrad_simt', lambda: _lib.check(L.seedrl_debug_conv3x3_wgrad(
                         cin, cout, mode, N, H, W, _lib.ptr(x), _lib.ptr(dy), _lib.ptr(dw), _lib.ptr(db), _lib.ptr(partial), pb,
                         _lib.stream_ptr())))):
      if name == 'wgrad_simt' and cin != 4:
        continue
      us = timed(fn)
      print('%-10s cin=%2d cout=%2d mode=%d %2dx%-2d  %8.1f us  %7.0f GB/s (x %d MB + dy %d MB)' %
            (name, cin, cout, mode, H, W, us, (xb + yb) / us / 1e3, xb >> 20, yb >> 20), flush=True)
  if what in ('fwd', 'all'):
    w = torch.randn(3, 3, cin, cout, device='cuda') * 0.1; b = torch.randn(cout, device='cuda')
    out = torch.empty(N, H, W, cout, device='cuda'); wq = torch.empty(2 * 9 * max(cin, 16) * cout * 2, dtype=torch.uint8, device='cuda')
    us = timed(lambda: _lib.check(L.seedrl_debug_conv3x3_tc(cin, cout, mode, split, N, H, W, _lib.ptr(x), _lib.ptr(w), _lib.ptr(b),
                                                           None, None, _lib.ptr(out), 0, 0, _lib.ptr(wq), _lib.ptr(err), _lib.stream_ptr())))
    print('%-10s cin=%2d cout=%2d mode=%d %2dx%-2d  %8.1f us  %7.0f GB/s (incl. weight pack launch)' %
          ('fwd_tc', cin, cout, mode, H, W, us, (xb + yb) / us / 1e3), flush=True)
  assert int(err.item()) == 0
