"""reference atari/networks.py: the bit-packed frame stacking (`stack_frames`,
`initial_frame_stacking_state`, :33-173) and the R2D2 agent network `DuelingLSTMDQNNet`
(:221-340, with `_unroll_cell` :176-218), same agent protocol as the reference:

    agent.initial_state(batch_size) -> AgentState(core_state=(h, c), frame_stacking_state)
    agent((prev_actions, env_outputs), agent_state, unroll=False)
        -> (AgentOutput(action int32, q_values float32), AgentState)
    agent.trainable_variables     (18 tensors, tf.Module order: _advantage, _body, _core, _value)

All math runs in libseedrl_b200 (seedrl_r2d2_stack_frames, seedrl_r2d2_net_forward /
_backward: csrc/r2d2_kernels.cu, csrc/r2d2_net.cu)."""
import collections
import ctypes
import math
import threading

import numpy as np
import torch

from seed_rl_b200 import _lib

AgentOutput = collections.namedtuple('AgentOutput', 'action q_values')
AgentState = collections.namedtuple('AgentState', 'core_state frame_stacking_state')
LSTM_UNITS = 512

STACKING_STATE_DTYPE = torch.int32


def initial_frame_stacking_state(stack_size, batch_size, observation_shape, device='cuda'):
  """reference :33-54: () when stack_size == 1, else zeros int32 [batch_size, prod(obs_shape)]."""
  if stack_size == 1:
    return ()
  n = 1
  for d in observation_shape:
    n *= int(d)
  return torch.zeros([batch_size, n], dtype=STACKING_STATE_DTYPE, device=device)


def stack_frames(frames, frame_stacking_state, done, stack_size):
  """reference :57-173.  frames: uint8 [time, batch, *obs, 1] (un-normalised); state int32
  [batch, prod(obs)] bit-packed; done bool [time, batch].  Returns (stacked uint8
  [time, batch, *obs, stack_size], newest first, frames across an episode boundary zeroed;
  new int32 state).  The reference returns float32 with the same values; here the stack stays
  uint8 and the /255 lives in the first convolution."""
  if tuple(frames.shape[0:2]) != tuple(done.shape[0:2]):
    raise ValueError('Expected same first 2 dims for frames and dones. Got {} vs {}.'.format(
        tuple(frames.shape[0:2]), tuple(done.shape[0:2])))
  if stack_size > 4:
    raise ValueError('Only up to stack size 4 is supported due to bit-packing.')
  if stack_size > 1 and frames.shape[-1] != 1:
    raise ValueError('Due to frame stacking, we require last observation dimension to be 1. Got {}'.format(
        frames.shape[-1]))
  if stack_size == 1:
    return frames, ()
  if frame_stacking_state.dtype != STACKING_STATE_DTYPE:
    raise ValueError('Expected dtype {} got {}'.format(STACKING_STATE_DTYPE, frame_stacking_state.dtype))
  fr = _lib.require_cuda(frames, torch.uint8, 'frames')
  st = _lib.require_cuda(frame_stacking_state, torch.int32, 'frame_stacking_state')
  dn = _lib.require_cuda(done, torch.bool, 'done')
  T, B = int(fr.shape[0]), int(fr.shape[1])
  obs = tuple(int(x) for x in fr.shape[2:-1])
  P = 1
  for d in obs:
    P *= d
  out = torch.empty((T, B) + obs + (stack_size,), dtype=torch.uint8, device=fr.device)
  new_state = torch.empty_like(st)
  _lib.check(_lib.lib().seedrl_r2d2_stack_frames(T, B, P, stack_size, _lib.ptr(fr), _lib.ptr(st), _lib.ptr(dn),
                                                 _lib.ptr(out), _lib.ptr(new_state), _lib.stream_ptr()))
  return out, new_state


class DuelingLSTMDQNNet(object):
  """reference atari/networks.py:221-340.  Conv 8x8/4 -> 32, 4x4/2 -> 64, 3x3/1 -> 64 ('valid',
  ReLU), Dense(512, ReLU), concat(reward, one_hot(prev_action)), LSTMCell(512) with done-resets,
  dueling value / advantage heads, greedy action.  Parameters live in one flat fp32 HBM arena
  (Keras layouts, tf.Module variable order)."""

  def __init__(self, num_actions, observation_shape, stack_size=1, seed=0, device=None, gemm_mode='tc3'):
    """gemm_mode: 'tc3' = tcgen05 bf16x3 (fp32-faithful) for every contraction (convolutions as
    im2col GEMMs, Dense, LSTM projection, heads); 'simt' = fp32 CUDA cores."""
    L = _lib.lib()
    self._num_actions = int(num_actions)
    self._observation_shape = tuple(int(x) for x in observation_shape)
    self._stack_size = int(stack_size)
    if len(self._observation_shape) != 3:
      raise ValueError('observation_shape must be [height, width, channels]')
    if self._stack_size > 1 and self._observation_shape[-1] != 1:
      raise ValueError('Due to frame stacking, we require last observation dimension to be 1. Got {}'.format(
          self._observation_shape[-1]))
    self._channels = self._stack_size if self._stack_size > 1 else self._observation_shape[-1]
    h = ctypes.c_void_p()
    _lib.check(L.seedrl_r2d2_net_create(self._num_actions, self._observation_shape[0], self._observation_shape[1],
                                        self._channels, ctypes.byref(h)))
    self._h = h
    modes = {'simt': 0, 'tc3': 2}
    if gemm_mode not in modes:
      raise ValueError("gemm_mode must be 'simt' or 'tc3'")
    self.gemm_mode = gemm_mode
    _lib.check(L.seedrl_r2d2_net_set_mode(h, modes[gemm_mode]))
    self._n_tensors = L.seedrl_r2d2_net_num_param_tensors(h)
    self.arena_floats = int(L.seedrl_r2d2_net_arena_floats(h))
    self.num_params = int(L.seedrl_r2d2_net_num_params(h))
    self.param_info = []
    for i in range(self._n_tensors):
      name = ctypes.create_string_buffer(128)
      dims = (ctypes.c_int64 * 4)()
      rank = ctypes.c_int()
      off = ctypes.c_size_t()
      _lib.check(L.seedrl_r2d2_net_param_info(h, i, name, 128, dims, ctypes.byref(rank), ctypes.byref(off)))
      self.param_info.append((name.value.decode(), tuple(int(dims[k]) for k in range(rank.value)), int(off.value)))
    self.device = torch.device(device if device is not None else ('cuda:%d' % torch.cuda.current_device()))
    self.params = torch.zeros(self.arena_floats, dtype=torch.float32, device=self.device)
    self.grads = torch.zeros_like(self.params)
    self._init_parameters(seed)
    self._workspaces = {}
    self._lock = threading.Lock()
    self._saved = None

  def __del__(self):
    try:
      if getattr(self, '_h', None):
        _lib.lib().seedrl_r2d2_net_destroy(self._h)
        self._h = None
    except Exception:   # interpreter shutdown
      pass

  # ---- parameters ---------------------------------------------------------------
  def _view(self, arena, i):
    _, shape, off = self.param_info[i]
    n = int(np.prod(shape)) if shape else 1
    return arena[off:off + n].view(shape if shape else ())

  @property
  def trainable_variables(self):
    return [self._view(self.params, i) for i in range(self._n_tensors)]

  @property
  def variable_names(self):
    return [p[0] for p in self.param_info]

  def named_parameters(self):
    return collections.OrderedDict((self.param_info[i][0], self._view(self.params, i)) for i in range(self._n_tensors))

  def named_gradients(self):
    return collections.OrderedDict((self.param_info[i][0], self._view(self.grads, i)) for i in range(self._n_tensors))

  def load_named_parameters(self, named):
    mine = self.named_parameters()
    for k, v in named.items():
      t = torch.as_tensor(np.asarray(v, np.float32))
      if tuple(t.shape) != tuple(mine[k].shape):
        raise ValueError('shape mismatch for %s: %s vs %s' % (k, tuple(t.shape), tuple(mine[k].shape)))
      mine[k].copy_(t)

  def assign_from(self, other):
    """update_target_agent (agents/r2d2/learner.py:535-544): target_var.assign(source_var)."""
    if other.arena_floats != self.arena_floats:
      raise ValueError('Mismatch in number of net tensors')
    self.params.copy_(other.params)

  def _init_parameters(self, seed):
    """Keras defaults (TF 2.4.1): glorot_uniform kernels, zero biases, orthogonal recurrent
    kernel, unit_forget_bias."""
    rng = np.random.default_rng(seed)
    for i in range(self._n_tensors):
      name, shape, _ = self.param_info[i]
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
      self._view(self.params, i).copy_(torch.from_numpy(a))

  # ---- protocol ---------------------------------------------------------------
  def initial_state(self, batch_size):
    z = torch.zeros([batch_size, LSTM_UNITS], dtype=torch.float32, device=self.device)
    return AgentState(core_state=(z, z.clone()),
                      frame_stacking_state=initial_frame_stacking_state(
                          self._stack_size, batch_size, self._observation_shape, device=self.device))

  def _workspace(self, T, B):
    key = (T, B, threading.get_ident())
    with self._lock:
      ws = self._workspaces.get(key)
      if ws is None:
        nbytes = int(_lib.lib().seedrl_r2d2_net_workspace_bytes(self._h, T, B))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._workspaces[key] = ws
    return ws

  def check_errors(self):
    if self._saved is None:
      return
    T, B, _, ws, _ = self._saved
    _lib.check(_lib.lib().seedrl_r2d2_net_check_error(self._h, T, B, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))

  def __call__(self, input_, agent_state, unroll=False, is_training=False):
    prev_actions, env_outputs = input_
    reward, done, observation = env_outputs[0], env_outputs[1], env_outputs[2]
    prev_actions = _lib.require_cuda(prev_actions.to(torch.int64), torch.int64, 'prev_actions')
    reward = _lib.require_cuda(reward, torch.float32, 'reward')
    done = _lib.require_cuda(done, torch.bool, 'done')
    observation = _lib.require_cuda(observation, torch.uint8, 'observation')
    if not unroll:    # add the time dimension (networks.py:309-312)
      prev_actions, reward, done, observation = (t.unsqueeze(0) for t in (prev_actions, reward, done, observation))
    T, B = int(prev_actions.shape[0]), int(prev_actions.shape[1])
    if tuple(observation.shape[2:]) != self._observation_shape:
      raise ValueError('observation shape %s, expected %s' % (tuple(observation.shape[2:]), self._observation_shape))
    stacked, frame_state = stack_frames(observation, agent_state.frame_stacking_state, done, self._stack_size)
    h0 = _lib.require_cuda(agent_state.core_state[0], torch.float32, 'core_state.h')
    c0 = _lib.require_cuda(agent_state.core_state[1], torch.float32, 'core_state.c')
    A = self._num_actions
    q = torch.empty([T, B, A], dtype=torch.float32, device=self.device)
    action = torch.empty([T, B], dtype=torch.int32, device=self.device)
    h = torch.empty_like(h0)
    c = torch.empty_like(c0)
    ws = self._workspace(T, B)
    _lib.check(_lib.lib().seedrl_r2d2_net_forward(
        self._h, _lib.ptr(self.params), T, B, _lib.ptr(prev_actions), _lib.ptr(reward), _lib.ptr(done),
        _lib.ptr(stacked), _lib.ptr(h0), _lib.ptr(c0), _lib.ptr(q), _lib.ptr(action), _lib.ptr(h), _lib.ptr(c),
        _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    if is_training:
      self._saved = (T, B, done, ws, stacked)
    out = AgentOutput(action, q)
    if not unroll:
      out = AgentOutput(*(t.squeeze(0) for t in out))
    return out, AgentState((h, c), frame_state)

  def backward(self, dq):
    """d loss / d parameters of the last is_training unroll -> self.grads (overwritten)."""
    if self._saved is None:
      raise RuntimeError('backward() needs a preceding __call__(..., unroll=True, is_training=True)')
    T, B, done, ws, stacked = self._saved
    dq = _lib.require_cuda(dq, torch.float32, 'dq')
    if tuple(dq.shape) != (T, B, self._num_actions):
      raise ValueError('dq must be [T, B, num_actions] of the training unroll')
    _lib.check(_lib.lib().seedrl_r2d2_net_backward(
        self._h, _lib.ptr(self.params), T, B, _lib.ptr(stacked), _lib.ptr(done), _lib.ptr(dq), _lib.ptr(self.grads),
        _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return self.grads

  def state_dict(self):
    return {'params': self.params.detach().cpu(), 'param_info': self.param_info}

  def load_state_dict(self, d):
    self.params.copy_(d['params'].to(self.device))
