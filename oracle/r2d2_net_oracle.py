"""CPU ORACLE (test infrastructure, NOT product code) -- torch-CPU fp32 restatement of the R2D2
agent network, SURVEY 8(a) row a11:

  DuelingLSTMDQNNet   /root/reference/atari/networks.py:221-340
  _unroll_cell        /root/reference/atari/networks.py:176-218
  stack_frames        oracle/r2d2_oracle.py (pinned bit-exact)

Keras layer arithmetic (TF 2.4.1 Conv2D(padding='valid') / Dense / LSTMCell) is restated from
its published semantics; the reference's tests pin shapes only (atari/networks_test.py:78-117:
unrolls run, the torso sees stack_size channels, the core input is 512 + num_actions + 1 wide)
=> layer NUMERICS parity-unpinned; pinned: variable structure, shapes, and the WIRING -- the
unmodified reference classes executed over a Keras-layer shim with these weights
(tests/golden/make_golden_r2d2_net.py) reproduce this file's Q values, greedy actions, LSTM
state and (bit for bit) packed frame state (tests/test_oracle_r2d2.py).

Weights use the Keras layouts (conv HWIO, dense [in,out], LSTM [in,4H]/[H,4H], gates i,f,c,o).
"""
import collections
import math

import numpy as np
import torch

from oracle import net_oracle, r2d2_oracle

AgentOutput = collections.namedtuple('AgentOutput', 'action q_values')
AgentState = collections.namedtuple('AgentState', 'core_state frame_stacking_state')

LSTM_UNITS = 512
CONVS = ((32, 8, 4), (64, 4, 2), (64, 3, 1))     # (filters, kernel, stride), padding 'valid'


def param_specs(num_actions, obs_shape, stack_size):
  """tf.Module.trainable_variables order is attribute-name order of the module
  (_advantage, _body, _core, _value); listed here in network order, names carry the owner."""
  h, w = obs_shape[0], obs_shape[1]
  cin = stack_size if stack_size > 1 else obs_shape[-1]
  specs = []
  for i, (f, k, s) in enumerate(CONVS):
    specs += [('body/conv%d/kernel' % i, (k, k, cin, f)), ('body/conv%d/bias' % i, (f,))]
    h, w = (h - k) // s + 1, (w - k) // s + 1
    cin = f
  flat = h * w * cin
  core_in = 512 + 1 + num_actions                              # networks.py:266-273
  specs += [('body/dense/kernel', (flat, 512)), ('body/dense/bias', (512,)),
            ('core/kernel', (core_in, 4 * LSTM_UNITS)), ('core/recurrent_kernel', (LSTM_UNITS, 4 * LSTM_UNITS)),
            ('core/bias', (4 * LSTM_UNITS,)),
            ('value/hidden/kernel', (LSTM_UNITS, 512)), ('value/hidden/bias', (512,)),
            ('value/head/kernel', (512, 1)), ('value/head/bias', (1,)),
            ('advantage/hidden/kernel', (LSTM_UNITS, 512)), ('advantage/hidden/bias', (512,)),
            ('advantage/head/kernel', (512, num_actions))]     # use_bias=False, networks.py:249-250
  return specs


def init_params(num_actions, obs_shape, stack_size, seed=0):
  """Keras defaults: glorot_uniform kernels, zero biases, orthogonal recurrent kernel,
  unit_forget_bias."""
  rng = np.random.default_rng(seed)
  out = collections.OrderedDict()
  for name, shape in param_specs(num_actions, obs_shape, stack_size):
    if name.endswith('bias'):
      a = np.zeros(shape, np.float32)
      if name == 'core/bias':
        a[LSTM_UNITS:2 * LSTM_UNITS] = 1.0
    elif name == 'core/recurrent_kernel':
      q, r = np.linalg.qr(rng.normal(size=(shape[1], shape[0])))
      a = (q * np.sign(np.diag(r))).T.astype(np.float32)
    else:
      rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
      lim = math.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
      a = rng.uniform(-lim, lim, shape).astype(np.float32)
    out[name] = a
  return out


def torso(p, prev_action, reward, frames01, num_actions):
  """_torso, networks.py:262-273: body(observation) ++ reward ++ one_hot(prev_action)."""
  x = frames01
  for i, (_, _, s) in enumerate(CONVS):
    x = torch.relu(net_oracle._conv_nhwc(x, p['body/conv%d/kernel' % i], p['body/conv%d/bias' % i], s, False))
  x = x.reshape(x.shape[0], -1)                                 # Flatten (NHWC order)
  x = torch.relu(x @ p['body/dense/kernel'] + p['body/dense/bias'])
  one_hot = torch.nn.functional.one_hot(prev_action.long(), num_actions).to(x.dtype)
  return torch.cat([x, reward[:, None], one_hot], dim=1)


def lstm_cell(p, x, h, c):
  z = x @ p['core/kernel'] + h @ p['core/recurrent_kernel'] + p['core/bias']
  i, f, g, o = z.chunk(4, dim=1)
  c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
  return torch.sigmoid(o) * torch.tanh(c2), c2


def head(p, core):
  """_head, networks.py:275-288: dueling combination and greedy action."""
  value = torch.relu(core @ p['value/hidden/kernel'] + p['value/hidden/bias']) @ p['value/head/kernel'] + \
      p['value/head/bias']
  adv = torch.relu(core @ p['advantage/hidden/kernel'] + p['advantage/hidden/bias']) @ p['advantage/head/kernel']
  adv = adv - adv.mean(dim=-1, keepdim=True)
  q = value + adv
  return q.argmax(dim=-1).to(torch.int32), q


def unroll(p, prev_actions, reward, done, observation_u8, agent_state, num_actions, stack_size):
  """_unroll, networks.py:321-340.  Time-major inputs [T,B,...]; observation uint8
  [T,B,H,W,1] (or C channels when stack_size == 1); agent_state = AgentState((h, c), int32
  frame-stacking state or ()).  Returns (AgentOutput(action [T,B] int32, q [T,B,A]), AgentState)."""
  T, B = prev_actions.shape
  frames = np.asarray(observation_u8).astype(np.float32)                      # :324
  stacked, frame_state = r2d2_oracle.stack_frames(frames, agent_state.frame_stacking_state,
                                                  np.asarray(done, bool), stack_size)   # :328-329
  x = torch.as_tensor(stacked) / 255                                          # :331
  tor = torso(p, torch.as_tensor(np.asarray(prev_actions)).reshape(T * B),
              torch.as_tensor(np.asarray(reward, np.float32)).reshape(T * B),
              x.reshape((T * B,) + tuple(x.shape[2:])), num_actions).reshape(T, B, -1)   # batch_apply
  h, c = agent_state.core_state
  d_all = torch.as_tensor(np.asarray(done, bool))
  outs = []
  for t in range(T):                                                          # _unroll_cell :204-217
    d = d_all[t][:, None]
    h = torch.where(d, torch.zeros_like(h), h)
    c = torch.where(d, torch.zeros_like(c), c)
    h, c = lstm_cell(p, tor[t], h, c)
    outs.append(h)
  action, q = head(p, torch.stack(outs).reshape(T * B, -1))                   # batch_apply(_head)
  return AgentOutput(action.reshape(T, B), q.reshape(T, B, -1)), AgentState((h, c), frame_state)


def initial_state(batch_size, obs_shape, stack_size):
  z = torch.zeros(batch_size, LSTM_UNITS)
  return AgentState((z, z.clone()), r2d2_oracle.initial_frame_stacking_state(stack_size, batch_size, obs_shape))
