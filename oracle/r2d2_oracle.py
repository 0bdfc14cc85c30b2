"""CPU restatement (numpy fp32) of the R2D2 post-network arithmetic -- SURVEY 8(a) row a11.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and, once the CUDA path of row a11 exists, by
its parity tests); never by the product package.

Follows the reference line by line (paths relative to the reference repository):
  value_function_rescaling / inverse     agents/r2d2/learner.py:180-192
  n_step_bellman_target                  agents/r2d2/learner.py:195-255
  compute_loss_and_priorities_from_agent_outputs   agents/r2d2/learner.py:258-330
  get_envs_epsilon                       agents/r2d2/learner.py:129-152
  PrioritizedReplay probabilities / importance weights   common/utils.py:327-352
  stack_frames (bit-packed frame stacking)               atari/networks.py:33-173
Pinned (tests/test_oracle_r2d2.py) against the known-answer cases of
agents/r2d2/learner_test.py:60-70,114-198, atari/networks_test.py:176-247 and against tests/golden/r2d2_golden.npz, produced
by executing the UNMODIFIED reference functions over tests/golden/tf_numpy_shim.py
(tests/golden/make_golden_r2d2.py); the replay functions additionally against the unmodified
PrioritizedReplay class run over a tf.Variable stand-in (tests/golden/make_golden_replay.py).
What stays unpinned: tf.random.categorical's Philox
stream (sampling is statistical in the reference's own tests too).
"""
import numpy as np

F = np.float32


def value_function_rescaling(x, eps=1e-3):
  """h(x) = sign(x) (sqrt(|x| + 1) - 1) + eps x        (learner.py:180-183)."""
  x = np.asarray(x, F)
  return (np.sign(x) * (np.sqrt(np.abs(x) + F(1.)) - F(1.)) + F(eps) * x).astype(F)


def inverse_value_function_rescaling(x, eps=1e-3):
  """h^-1, Proposition A.2 of "Observe and Look Further"   (learner.py:186-192)."""
  x = np.asarray(x, F)
  e = F(eps)
  inner = (np.sqrt(F(1.) + F(4.) * e * (np.abs(x) + F(1.) + e)) - F(1.)) / (F(2.) * e)
  return (np.sign(x) * (np.square(inner) - F(1.))).astype(F)


def n_step_bellman_target(rewards, done, q_target, gamma, n_steps):
  """learner.py:237-255: q_target padded with the last value divided by gamma^k, rewards/done
  padded with n_steps zeros, then n_steps passes of
      target = r + gamma (1 - done) target[1:]
  each dropping the last row of rewards/done."""
  rewards = np.asarray(rewards, F); q_target = np.asarray(q_target, F)
  done = np.asarray(done, bool)
  g = F(gamma)
  target = np.concatenate([np.zeros_like(q_target[0:1]), q_target] +
                          [q_target[-1:] / F(gamma ** k) for k in range(1, n_steps)], axis=0)
  done = np.concatenate([done] + [np.zeros_like(done[0:1])] * n_steps, axis=0)
  rewards = np.concatenate([rewards] + [np.zeros_like(rewards[0:1])] * n_steps, axis=0)
  for _ in range(n_steps):
    rewards = rewards[:-1]
    done = done[:-1]
    target = (rewards + g * (F(1.) - done.astype(F)) * target[1:]).astype(F)
  return target


def loss_and_priorities(train_q, train_action, target_q, replay_action, reward, done, gamma,
                        n_steps=5, eta=0.9, eps=1e-3):
  """compute_loss_and_priorities_from_agent_outputs, learner.py:290-330.
  train_q, target_q: [T,B,A]; train_action (argmax of the online net), replay_action: [T,B];
  reward, done: [T,B].  Returns (loss [B], priorities [B], abs_td [T-1,B])."""
  train_q = np.asarray(train_q, F); target_q = np.asarray(target_q, F)
  T, B, A = train_q.shape
  tt, bb = np.meshgrid(np.arange(T), np.arange(B), indexing='ij')
  replay_q = train_q[tt, bb, np.asarray(replay_action)]                       # :295-296
  qtarget_max = inverse_value_function_rescaling(target_q[tt, bb, np.asarray(train_action)], eps)   # :303-305
  target = n_step_bellman_target(reward, done, qtarget_max, gamma, n_steps)   # :308-313
  target = target[1:]                                                         # :316
  replay_q = replay_q[:-1]                                                    # :318
  target = value_function_rescaling(target, eps)                              # :320
  abs_td = np.abs(target - replay_q).astype(F)                                # :322
  priorities = (F(eta) * abs_td.max(axis=0) + F(1 - eta) * abs_td.mean(axis=0, dtype=F)).astype(F)   # :325-326
  loss = (F(0.5) * np.square(abs_td).sum(axis=0, dtype=F)).astype(F)          # :329
  return loss, priorities, abs_td


def get_envs_epsilon(env_ids, num_training_envs, num_eval_envs, eval_epsilon):
  """learner.py:146-152: 0.4 ** linspace(1, 8, num_training_envs), then eval epsilons."""
  eps = np.concatenate([np.power(F(0.4), np.linspace(1., 8., num_training_envs, dtype=F)).astype(F),
                        np.full([num_eval_envs], eval_epsilon, F)])
  return eps[np.asarray(env_ids)]


def replay_probabilities(priorities, num_inserted, priority_exp):
  """common/utils.py:337-345: p_i = prio_i^alpha / sum over the filled part of the ring."""
  size = len(priorities)
  limit = min(size, int(num_inserted))
  prob = np.power(np.asarray(priorities, F)[:limit], F(priority_exp)).astype(F)
  return (prob / prob.sum(dtype=F)).astype(F)


def replay_importance_weights(prob, indices, importance_sampling_exponent):
  """common/utils.py:347-351: ((1/limit) / p_i)^beta, normalised by the max."""
  limit = F(len(prob))
  w = np.power((F(1.) / limit) / prob[np.asarray(indices)], F(importance_sampling_exponent)).astype(F)
  return (w / w.max()).astype(F)


def replay_insert_indices(num_inserted, append_size, size):
  """common/utils.py:296-303: FIFO ring insertion indices."""
  return np.arange(num_inserted, num_inserted + append_size) % size


def stack_frames(frames, frame_stacking_state, done, stack_size):
  """atari/networks.py:57-173.  frames f32 [T,B,...obs,1] in [0,255]; state int32
  [B, prod(obs)] bit-packed (LSB byte = oldest of the stack_size-1 kept frames); done bool
  [T,B].  Returns (stacked f32 [T,B,...obs,stack_size], newest first, frames across an
  episode boundary zeroed; new int32 state [B, prod(obs)]).  Byte/integer work: bit-exact."""
  frames = np.asarray(frames, F); done = np.asarray(done, bool)
  if frames.shape[0:2] != done.shape[0:2]:
    raise ValueError('Expected same first 2 dims for frames and dones. Got {} vs {}.'.format(
        frames.shape[0:2], done.shape[0:2]))
  if stack_size > 4:
    raise ValueError('Only up to stack size 4 is supported due to bit-packing.')
  if stack_size > 1 and frames.shape[-1] != 1:
    raise ValueError('Due to frame stacking, we require last observation dimension to be 1. Got {}'.format(
        frames.shape[-1]))
  if stack_size == 1:
    return frames, ()
  T, B = frames.shape[:2]
  obs_shape = frames.shape[2:-1]
  state = np.asarray(frame_stacking_state)
  if state.dtype != np.int32:
    raise ValueError('Expected dtype int32 got {}'.format(state.dtype))
  state = state.reshape((B,) + obs_shape)
  # unpacked previous frames, oldest first (:113-119)
  prev = [((state >> (8 * i)) & 0xFF).astype(F) for i in range(stack_size - 1)]
  ext = np.concatenate([p.reshape((1,) + p.shape + (1,)) for p in prev] + [frames], axis=0)   # :124-128
  stacked = np.concatenate([ext[stack_size - 1 - i: ext.shape[0] - i] for i in range(stack_size)], axis=-1)  # :135-138
  # masks of frames that cross an episode boundary (:143-168)
  row = done.reshape(done.shape + (1,) * (frames.ndim - 2))
  masks = [np.zeros_like(row), row]
  while len(masks) < stack_size:
    p = masks[-1]
    masks.append(p | np.concatenate([np.zeros_like(p[:1]), p[:-1]], axis=0))
  stacked = np.where(np.concatenate(masks, axis=-1), F(0), stacked).astype(F)
  # new bit-packed state from the zeroed stack: MSB byte = newest (:175-183)
  shifts = np.array([8 * i for i in range(stack_size - 2, -1, -1)], np.int32)
  new_state = (stacked[-1, ..., :-1].astype(np.int32) << shifts).sum(axis=-1, dtype=np.int32)
  return stacked, new_state.reshape(B, int(np.prod(obs_shape)))


def initial_frame_stacking_state(stack_size, batch_size, observation_shape):
  """atari/networks.py:33-54."""
  if stack_size == 1:
    return ()
  return np.zeros([batch_size, int(np.prod(observation_shape))], np.int32)
