"""reference agents/r2d2/learner.py -- round-1 scope: the post-network arithmetic of the R2D2
learner behind the reference's function names:

  compute_loss_and_priorities_from_agent_outputs   :258-330 (+ value rescaling :180-192 and
                                                   n-step Bellman targets :195-255, fused)
  PrioritizedReplay.sample's probabilities / importance weights   common/utils.py:327-352
  tf.clip_by_global_norm of the minimize step      :608

STATUS: kernels written and compiled in round 1, not yet executed on hardware
(tests/test_gpu_r2d2.py is gated); the network unrolls, replay storage and the learner loop of
cfg 5 are not built."""
import collections

import torch

from seed_rl_b200 import _lib

AgentOutput = collections.namedtuple('AgentOutput', 'action q_values')

# flag defaults of the reference (learner.py:80-87)
N_STEPS = 5
VALUE_FUNCTION_RESCALING_EPSILON = 1e-3


def compute_loss_and_priorities_from_agent_outputs(training_agent_output, target_agent_output, env_outputs,
                                                   agent_outputs, gamma, eta=0.9, n_steps=N_STEPS,
                                                   importance_weights=None,
                                                   value_function_rescaling_epsilon=VALUE_FUNCTION_RESCALING_EPSILON):
  """reference :258-330.  Returns (loss [B], priorities [B]); the gradient of
  mean(loss * importance_weights) w.r.t. training_agent_output.q_values (reference :604) is
  returned as the third element (the reference gets it from the tape)."""
  f32 = torch.float32
  q = _lib.require_cuda(training_agent_output.q_values, f32, 'training q_values')
  qt = _lib.require_cuda(target_agent_output.q_values, f32, 'target q_values')
  act = _lib.require_cuda(agent_outputs.action.to(torch.int64), torch.int64, 'replay actions')
  rew = _lib.require_cuda(env_outputs.reward, f32, 'reward')
  dn = _lib.require_cuda(env_outputs.done, torch.bool, 'done')
  if q.dim() != 3 or tuple(qt.shape) != tuple(q.shape):
    raise ValueError('q_values must be [time, batch, num_actions] for both agents')
  T, B, A = (int(x) for x in q.shape)
  w = None if importance_weights is None else _lib.require_cuda(importance_weights, f32, 'importance_weights')
  L = _lib.lib()
  loss = torch.empty(B, dtype=f32, device=q.device); prio = torch.empty_like(loss)
  dq = torch.empty_like(q)
  scratch = torch.empty(int(L.seedrl_r2d2_loss_scratch_bytes(T, B, n_steps)), dtype=torch.uint8, device=q.device)
  _lib.check(L.seedrl_r2d2_loss_fwd_bwd(T, B, A, _lib.ptr(q), _lib.ptr(qt), _lib.ptr(act), _lib.ptr(rew), _lib.ptr(dn),
                                        _lib.ptr(w), float(gamma), int(n_steps), float(eta),
                                        float(value_function_rescaling_epsilon), _lib.ptr(loss), _lib.ptr(prio),
                                        _lib.ptr(dq), _lib.ptr(scratch), _lib.stream_ptr()))
  return loss, prio, dq


def replay_sample(priorities, num_inserted, num_samples, priority_exp, importance_sampling_exponent,
                  uniforms=None, generator=None):
  """PrioritizedReplay.sample's index / weight part (common/utils.py:327-352) for
  priority_exp != 0.  priorities: float32 [size] on the GPU; returns (indices int64
  [num_samples], weights float32 [num_samples], probabilities [limit])."""
  pr = _lib.require_cuda(priorities, torch.float32, 'priorities')
  limit = min(int(pr.numel()), int(num_inserted))
  if limit <= 0:
    raise ValueError('Cannot sample if replay buffer is empty')
  if uniforms is None:
    uniforms = torch.rand(num_samples, device=pr.device, generator=generator)
  u = _lib.require_cuda(uniforms, torch.float32, 'uniforms')
  idx = torch.empty(num_samples, dtype=torch.int64, device=pr.device)
  wts = torch.empty(num_samples, dtype=torch.float32, device=pr.device)
  probs = torch.empty(limit, dtype=torch.float32, device=pr.device)
  _lib.check(_lib.lib().seedrl_replay_sample(limit, _lib.ptr(pr), float(priority_exp),
                                             float(importance_sampling_exponent), int(num_samples), _lib.ptr(u),
                                             _lib.ptr(idx), _lib.ptr(wts), _lib.ptr(probs), _lib.stream_ptr()))
  return idx, wts, probs


def clip_by_global_norm(flat_grads, clip_norm):
  """tf.clip_by_global_norm over the flat gradient arena, in place.  Returns the global norm."""
  g = _lib.require_cuda(flat_grads, torch.float32, 'gradients')
  L = _lib.lib()
  norm = torch.empty((), dtype=torch.float32, device=g.device)
  scratch = torch.empty(int(L.seedrl_clip_scratch_bytes()), dtype=torch.uint8, device=g.device)
  _lib.check(L.seedrl_clip_by_global_norm(g.numel(), _lib.ptr(g), float(clip_norm), _lib.ptr(norm), _lib.ptr(scratch),
                                          _lib.stream_ptr()))
  return norm
