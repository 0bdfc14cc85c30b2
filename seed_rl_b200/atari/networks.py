"""reference atari/networks.py -- round-1 scope: the bit-packed frame stacking only
(`stack_frames`, `initial_frame_stacking_state`, :33-173).  DuelingLSTMDQNNet's CUDA schedule
is the next step of SURVEY 8(a) row a11; its oracle is oracle/r2d2_net_oracle.py.

STATUS: the kernel behind `stack_frames` was written and compiled in round 1 but has not been
executed on hardware yet (tests/test_gpu_r2d2.py is gated)."""
import torch

from seed_rl_b200 import _lib

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
