"""CPU: the hand-derived backward schedule the CUDA network implements (net.cu /
gemm_kernels.cu LSTM BPTT with done-resets) restated in numpy and checked against torch
autograd through the oracle's unroll.  Guards the formulas before they reach the GPU."""
import numpy as np
import torch

from oracle import net_oracle


def _sig(x):
  return 1.0 / (1.0 + np.exp(-x))


def test_lstm_bptt_with_done_resets_matches_autograd():
  rng = np.random.default_rng(0)
  T, B, H, X = 5, 3, 256, 20
  Wk = (rng.normal(size=(X, 4 * H)) * 0.1).astype(np.float32)
  U = (rng.normal(size=(H, 4 * H)) * 0.1).astype(np.float32)
  b = (rng.normal(size=(4 * H,)) * 0.1).astype(np.float32)
  x = rng.normal(size=(T, B, X)).astype(np.float32)
  done = rng.random((T, B)) < 0.3
  done[0, 0] = True
  h0 = rng.normal(size=(B, H)).astype(np.float32); c0 = rng.normal(size=(B, H)).astype(np.float32)
  dH = rng.normal(size=(T, B, H)).astype(np.float32)

  # ---- autograd reference (oracle cell) ----
  p = {'core/kernel': torch.tensor(Wk, requires_grad=True),
       'core/recurrent_kernel': torch.tensor(U, requires_grad=True),
       'core/bias': torch.tensor(b, requires_grad=True)}
  xt = torch.tensor(x, requires_grad=True)
  h, c = torch.tensor(h0), torch.tensor(c0)
  outs = []
  for t in range(T):
    d = torch.tensor(done[t])[:, None]
    h = torch.where(d, torch.zeros_like(h), h); c = torch.where(d, torch.zeros_like(c), c)
    h, c = net_oracle.lstm_cell(p, xt[t], h, c)
    outs.append(h)
  (torch.stack(outs) * torch.tensor(dH)).sum().backward()

  # ---- the CUDA schedule in numpy ----
  z = x.reshape(T * B, X) @ Wk + b
  z = z.reshape(T, B, 4 * H)
  hp = np.zeros((T, B, H), np.float32); cs = np.zeros((T, B, H), np.float32)
  hs = np.zeros((T, B, H), np.float32); gates = np.zeros_like(z)
  hp[0] = np.where(done[0][:, None], 0, h0)
  for t in range(T):
    zt = z[t] + hp[t] @ U
    gi, gf, gg, go = _sig(zt[:, :H]), _sig(zt[:, H:2 * H]), np.tanh(zt[:, 2 * H:3 * H]), _sig(zt[:, 3 * H:])
    cprev = np.where(done[t][:, None], 0, c0 if t == 0 else cs[t - 1])
    cs[t] = gf * cprev + gi * gg
    hs[t] = go * np.tanh(cs[t])
    gates[t] = np.concatenate([gi, gf, gg, go], 1)
    if t + 1 < T:
      hp[t + 1] = np.where(done[t + 1][:, None], 0, hs[t])
  dz = np.zeros_like(z)
  dhrec = None; dcn = None
  for t in range(T - 1, -1, -1):
    gi, gf, gg, go = np.split(gates[t], 4, 1)
    cut = done[t + 1][:, None] if t + 1 < T else None
    dh = dH[t].copy()
    if dhrec is not None:
      dh += np.where(cut, 0, dhrec)
    tc = np.tanh(cs[t])
    dc = dh * go * (1 - tc * tc)
    if dcn is not None:
      dc += np.where(cut, 0, dcn)
    cprev = np.where(done[t][:, None], 0, c0 if t == 0 else cs[t - 1])
    dz[t] = np.concatenate([dc * gg * gi * (1 - gi), dc * cprev * gf * (1 - gf),
                            dc * gi * (1 - gg * gg), dh * tc * go * (1 - go)], 1)
    dcn = dc * gf
    dhrec = dz[t] @ U.T
  dZ = dz.reshape(T * B, 4 * H)
  dU = hp.reshape(T * B, H).T @ dZ
  dWk = x.reshape(T * B, X).T @ dZ
  db = dZ.sum(0)
  dX = dZ @ Wk.T
  np.testing.assert_allclose(dU, p['core/recurrent_kernel'].grad.numpy(), rtol=2e-3, atol=2e-4)
  np.testing.assert_allclose(dWk, p['core/kernel'].grad.numpy(), rtol=2e-3, atol=2e-4)
  np.testing.assert_allclose(db, p['core/bias'].grad.numpy(), rtol=2e-3, atol=2e-4)
  np.testing.assert_allclose(dX.reshape(T, B, X), xt.grad.numpy(), rtol=2e-3, atol=2e-4)


def test_conv_dgrad_is_conv_with_flipped_transposed_weights():
  """conv_kernels.cu: dX = conv3x3(dY, W'[8-tap][co][ci]); res-block wiring of net.cu."""
  import torch.nn.functional as F
  rng = np.random.default_rng(1)
  N, Hh, Ww, Ci, Co = 2, 6, 5, 4, 8
  x = torch.tensor(rng.normal(size=(N, Hh, Ww, Ci)).astype(np.float32), requires_grad=True)
  w = torch.tensor(rng.normal(size=(3, 3, Ci, Co)).astype(np.float32))
  dy = torch.tensor(rng.normal(size=(N, Hh, Ww, Co)).astype(np.float32))
  y = net_oracle._conv_nhwc(x, w, None, 1, True)
  (y * dy).sum().backward()
  wt = torch.zeros(3, 3, Co, Ci)
  for tap in range(9):
    wt[(8 - tap) // 3, (8 - tap) % 3] = w[tap // 3, tap % 3].T
  dx = net_oracle._conv_nhwc(dy, wt, None, 1, True)
  np.testing.assert_allclose(dx.numpy(), x.grad.numpy(), rtol=1e-4, atol=1e-5)
