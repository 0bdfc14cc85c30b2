"""CPU: the Keras-layer arithmetic of oracle/net_oracle.py (restated from TF's published
semantics -- TF 2.4.1 is not vendored under /root/reference) cross-checked against
a second, independent restatement written as explicit numpy loops straight from the TF
documentation of `padding='SAME'` (out = ceil(n / s); pad_total = max((out-1) s + k - n, 0);
pad_before = pad_total // 2, the odd pixel goes AFTER), Conv2D (cross-correlation, HWIO
kernels), MaxPool2D (padding never wins: -inf) and LSTMCell (gates i, f, c, o; c' = sigma(f) c +
sigma(i) tanh(g); h' = sigma(o) tanh(c')).  Two restatements that agree do not replace a TF
run, but they remove "the oracle misreads the semantics in one place" as a failure mode."""
import numpy as np
import pytest
import torch

from oracle import net_oracle


def _same_pad(n, k, s):
  out = -(-n // s)
  total = max((out - 1) * s + k - n, 0)
  return out, total // 2


def conv2d_loops(x, w, b, stride, same):
  N, H, W, C = x.shape
  kh, kw, _, O = w.shape
  if same:
    Ho, pt = _same_pad(H, kh, stride); Wo, pl = _same_pad(W, kw, stride)
  else:
    Ho, Wo, pt, pl = (H - kh) // stride + 1, (W - kw) // stride + 1, 0, 0
  y = np.zeros((N, Ho, Wo, O), np.float64)
  for n in range(N):
    for i in range(Ho):
      for j in range(Wo):
        for a in range(kh):
          for c in range(kw):
            h, ww = i * stride - pt + a, j * stride - pl + c
            if 0 <= h < H and 0 <= ww < W:
              y[n, i, j] += x[n, h, ww].astype(np.float64) @ w[a, c].astype(np.float64)
  return y + (0 if b is None else b)


def maxpool_loops(x, k=3, s=2):
  N, H, W, C = x.shape
  Ho, pt = _same_pad(H, k, s); Wo, pl = _same_pad(W, k, s)
  y = np.full((N, Ho, Wo, C), -np.inf)
  for i in range(Ho):
    for j in range(Wo):
      for a in range(k):
        for c in range(k):
          h, w = i * s - pt + a, j * s - pl + c
          if 0 <= h < H and 0 <= w < W:
            y[:, i, j] = np.maximum(y[:, i, j], x[:, h, w])
  return y


@pytest.mark.parametrize('H,W,C,O,k,s,same', [(7, 5, 3, 4, 3, 1, True), (10, 9, 4, 8, 3, 1, True), (20, 20, 4, 6, 8, 4, False),
                                              (9, 9, 6, 5, 4, 2, False), (6, 6, 2, 3, 3, 2, True), (5, 8, 2, 2, 2, 2, True)])
def test_conv2d_semantics(H, W, C, O, k, s, same):
  rng = np.random.default_rng(H * W + k)
  x = rng.normal(size=(2, H, W, C)).astype(np.float32); w = rng.normal(size=(k, k, C, O)).astype(np.float32)
  b = rng.normal(size=O).astype(np.float32)
  got = net_oracle._conv_nhwc(torch.as_tensor(x), torch.as_tensor(w), torch.as_tensor(b), s, same).numpy()
  want = conv2d_loops(x, w, b, s, same)
  assert got.shape == want.shape
  np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('H,W', [(84, 84), (42, 42), (21, 21), (7, 5), (4, 4), (1, 1), (2, 3)])
def test_maxpool_3x3_stride2_same(H, W):
  rng = np.random.default_rng(H + W)
  x = rng.normal(size=(2, H, W, 3)).astype(np.float32)
  got = net_oracle._maxpool_same_nhwc(torch.as_tensor(x)).numpy()
  want = maxpool_loops(x)
  assert got.shape == (2, -(-H // 2), -(-W // 2), 3)
  np.testing.assert_array_equal(got, want.astype(np.float32))


def test_lstm_cell_equations():
  rng = np.random.default_rng(0)
  B, I, Hd = 3, 7, net_oracle.LSTM_UNITS
  p = {'core/kernel': rng.normal(size=(I, 4 * Hd)).astype(np.float32) * 0.1,
       'core/recurrent_kernel': rng.normal(size=(Hd, 4 * Hd)).astype(np.float32) * 0.1,
       'core/bias': rng.normal(size=4 * Hd).astype(np.float32) * 0.1}
  x = rng.normal(size=(B, I)).astype(np.float32); h = rng.normal(size=(B, Hd)).astype(np.float32)
  c = rng.normal(size=(B, Hd)).astype(np.float32)
  h2, c2 = net_oracle.lstm_cell({k: torch.as_tensor(v) for k, v in p.items()}, torch.as_tensor(x), torch.as_tensor(h),
                                torch.as_tensor(c))
  z = x.astype(np.float64) @ p['core/kernel'] + h.astype(np.float64) @ p['core/recurrent_kernel'] + p['core/bias']
  sig = lambda v: 1 / (1 + np.exp(-v))
  i, f, g, o = (z[:, k * Hd:(k + 1) * Hd] for k in range(4))          # Keras gate order i, f, c, o
  cw = sig(f) * c + sig(i) * np.tanh(g)
  hw = sig(o) * np.tanh(cw)
  np.testing.assert_allclose(c2.numpy(), cw, rtol=1e-5, atol=1e-5)
  np.testing.assert_allclose(h2.numpy(), hw, rtol=1e-5, atol=1e-5)
