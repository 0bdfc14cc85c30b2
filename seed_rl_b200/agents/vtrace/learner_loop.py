"""`learner_loop` -- mirror of reference agents/vtrace/learner.py:170-483: the central
inference closure bound on the RPC server (a6, :349-407), the unroll store / queue plumbing
(a8/a9, :314-336,394-399,418-445) and the training loop (:467-483), for ONE replica
(= one GPU / process; `torchrun` starts one per GPU, each with its own env shard, like the
reference's per-host server/store, :314-416).

Data path per inference batch: actor payloads land in a pinned slab (C++ batcher) ->
one H2D copy per field -> T=1 `seedrl_net_forward` + in-kernel sampling -> scatter into the
HBM-resident UnrollStore -> completed unrolls (already on the GPU) go to the capacity-1
queue -> the learner thread stacks B of them straight into time-major order.
Observations are never copied on the host after the slab (reference: >= 4 host copies).
"""
import collections
import math
import os
import threading
import time

from absl import flags
from absl import logging
import numpy as np
import torch

from seed_rl_b200.agents.vtrace import learner as learner_lib
from seed_rl_b200.common import utils
from seed_rl_b200.common.parametric_distribution import get_parametric_distribution_for_action_space
from seed_rl_b200.dmlab import networks
from seed_rl_b200.grpc import ops as grpc

FLAGS = flags.FLAGS
Unroll = learner_lib.Unroll


class InferenceHost(object):
  """Everything `create_host` builds in the reference (learner.py:314-413) for one GPU."""

  def __init__(self, agent, num_envs, unroll_length, inference_batch_size, obs_shape,
               num_action_repeats=1, device='cuda', info_queue=None):
    self.agent = agent
    self.device = torch.device(device)
    self.N = inference_batch_size
    self.num_action_repeats = num_action_repeats
    TS = utils.TensorSpec
    self.env_output_specs = utils.EnvOutput(
        TS([], 'float32', 'reward'), TS([], 'bool', 'done'),
        TS(list(obs_shape), 'uint8', 'observation'), TS([], 'bool', 'abandoned'),
        TS([], 'int32', 'episode_step'))
    action_specs = TS([], 'int64', 'action')
    A = agent._num_actions
    agent_output_specs = networks.AgentOutput(
        TS([], 'int64', 'action'), TS([A], 'float32', 'policy_logits'), TS([], 'float32', 'baseline'))
    agent_state_specs = (TS([networks.LSTM_UNITS], 'float32', 'h'), TS([networks.LSTM_UNITS], 'float32', 'c'))
    # time_major=True: completed unrolls come out as [T+1, n, ...] (no make_time_major pass)
    self.store = utils.UnrollStore(num_envs, unroll_length,
                                   (action_specs, self.env_output_specs, agent_output_specs),
                                   device=device, time_major=True)
    # run ids / episode stats feed host-side logging only -> host tables (learner.py:321-324)
    self.env_run_ids = np.zeros([num_envs], np.int64)
    self.env_infos = [np.zeros([num_envs], np.int64), np.zeros([num_envs], np.float32),
                      np.zeros([num_envs], np.float32)]
    self.first_agent_states = utils.Aggregator(num_envs, agent_state_specs, 'first_agent_states', device)
    self.agent_states = utils.Aggregator(num_envs, agent_state_specs, 'agent_states', device)
    self.actions = utils.Aggregator(num_envs, action_specs, 'actions', device)
    self.unroll_specs = Unroll(agent_state_specs, *self.store.unroll_specs)
    self.unroll_queue = utils.StructuredFIFOQueue(1, self.unroll_specs)      # capacity 1, :336
    self.info_queue = info_queue
    N = self.N
    self.inference_specs = (
        TS([N], 'int32', 'env_id'), TS([N], 'int64', 'run_id'),
        utils.map_structure(lambda s: TS([N] + list(s.shape), s.dtype, s.name), self.env_output_specs),
        TS([N], 'float32', 'raw_reward'))
    self.output_specs = TS([N], 'int64', 'action')
    self.stream = torch.cuda.Stream(device=self.device) if self.device.type == 'cuda' else None

    @grpc.function(self.inference_specs, self.output_specs)
    def inference(env_ids, run_ids, env_outputs, raw_rewards):
      return self._inference(env_ids, run_ids, env_outputs, raw_rewards)
    self.inference = inference

  def _inference(self, env_ids, run_ids, env_outputs, raw_rewards):
    """reference learner.py:351-405."""
    env_ids = np.asarray(env_ids); run_ids = np.asarray(run_ids)
    reward, done = np.asarray(env_outputs.reward), np.asarray(env_outputs.done)
    # Reset the environments that had their first run or crashed (:353-366).
    previous = self.env_run_ids[env_ids]
    self.env_run_ids[env_ids] = run_ids
    reset_ids = env_ids[previous != run_ids]
    with torch.cuda.stream(self.stream):
      if reset_ids.size:
        logging.info('Environment ids needing reset: %s', reset_ids)
        for t in self.env_infos:
          t[reset_ids] = 0
        self.store.reset(reset_ids)
        init = self.agent.initial_state(len(reset_ids))
        self.first_agent_states.replace(reset_ids, init)
        self.agent_states.replace(reset_ids, init)
        self.actions.reset(reset_ids)
      if np.asarray(env_outputs.abandoned).any():                       # :368-370
        raise ValueError('Abandoned done states are not supported in VTRACE.')
      # Update steps and return (:373-378).
      self.env_infos[1][env_ids] += reward
      self.env_infos[2][env_ids] += np.asarray(raw_rewards)
      done_ids = env_ids[done]
      if self.info_queue is not None and done_ids.size:
        self.info_queue.enqueue_many(tuple(torch.as_tensor(t[done_ids]) for t in self.env_infos))
      for t in self.env_infos:
        t[done_ids] = 0
      self.env_infos[0][env_ids] += self.num_action_repeats
      # Inference (:381-390): one H2D copy per field, T=1 forward on the GPU.
      ids_dev = torch.as_tensor(env_ids.astype(np.int64)).to(self.device, non_blocking=True)
      env_dev = utils.EnvOutput(*(torch.as_tensor(np.asarray(x)).to(self.device, non_blocking=True)
                                  for x in env_outputs))
      prev_actions = self.actions.read(ids_dev)
      prev_states = self.agent_states.read(ids_dev)
      agent_outputs, curr_states = self.agent(prev_actions, env_dev, prev_states, is_training=False)
      # Append to the unroll store, enqueue completed unrolls (:394-399).
      completed_ids, unrolls = self.store.append(env_ids, (prev_actions, env_dev, agent_outputs),
                                                 check_duplicates=True)
      n_done = int(completed_ids.numel())
      pending = []
      if n_done:
        first = self.first_agent_states.read(completed_ids)
        flat = utils.flatten(unrolls)
        for i in range(n_done):     # one queue element per unroll, as in the reference
          u = utils.pack_sequence_as(self.store._specs, [f[:, i] for f in flat])
          pending.append(Unroll((first[0][i], first[1][i]), *u))
        self.first_agent_states.replace(completed_ids, self.agent_states.read(completed_ids))
      # Update current state (:402-403) and return the actions (:405).
      self.agent_states.replace(ids_dev, curr_states)
      self.actions.replace(ids_dev, agent_outputs.action)
      out = agent_outputs.action.cpu()       # D2H + sync of this stream
    # The unrolls were produced on self.stream, which the blocking copy above has drained: only
    # now are they handed to the learner thread (which consumes them on another stream).
    for u in pending:
      self.unroll_queue.enqueue(u)
    return out.numpy()


def dequeue_batch(unroll_queue, batch_size):
  """reference learner.py:418-432: B unrolls -> one time-major batch.  Unrolls are already
  [T+1, ...] on the GPU; stacking along dim 1 IS the time-major layout."""
  items = [unroll_queue.dequeue() for _ in range(batch_size)]
  if torch.cuda.is_available():
    # the items were allocated on the inference stream: tell the caching allocator that this
    # stream reads them, so their blocks are not recycled under the (asynchronous) stack below
    cur = torch.cuda.current_stream()
    for it in items:
      for t in utils.flatten(it):
        if isinstance(t, torch.Tensor) and t.is_cuda:
          t.record_stream(cur)
  state = tuple(torch.stack([it.agent_state[k] for it in items]) for k in range(2))
  def stack(field):
    return utils.map_structure(lambda *xs: torch.stack(xs, dim=1),
                               *[getattr(it, field) for it in items])
  return Unroll(state, stack('prev_actions'), stack('env_outputs'), stack('agent_outputs'))


def learner_loop(create_env_fn, create_agent_fn, create_optimizer_fn):
  """reference learner.py:170-483 (single replica)."""
  logging.info('Starting learner loop')
  utils.validate_learner_config(FLAGS)
  env = create_env_fn(0, FLAGS)
  dist = get_parametric_distribution_for_action_space(env.action_space)
  agent = create_agent_fn(env.action_space, env.observation_space, dist)
  if not hasattr(agent, '_entropy_mul'):
    agent.init_entropy_cost(FLAGS.entropy_cost, FLAGS.entropy_cost_adjustment_speed)
  iter_frame_ratio = FLAGS.batch_size * FLAGS.unroll_length * FLAGS.num_action_repeats
  final_iteration = int(math.ceil(FLAGS.total_environment_frames / iter_frame_ratio))
  optimizer, learning_rate_fn = create_optimizer_fn(final_iteration)
  settings = learner_lib.loss_settings_from_flags()
  step = learner_lib.LearnerStep(agent, optimizer, dist, settings, grad_reduce=FLAGS.grad_reduce)

  ckpt_path = os.path.join(FLAGS.logdir, 'ckpt.pt')
  os.makedirs(FLAGS.logdir, exist_ok=True)
  init = FLAGS.init_checkpoint or (ckpt_path if os.path.exists(ckpt_path) else None)
  if init:                                                                 # :286-296
    logging.info('Restoring checkpoint: %s', init)
    d = torch.load(init, map_location='cpu')
    agent.load_state_dict(d['agent'])
    optimizer.load_state_dict(d['optimizer'])

  def save():
    torch.save({'agent': agent.state_dict(), 'optimizer': optimizer.state_dict()}, ckpt_path + '.tmp')
    os.replace(ckpt_path + '.tmp', ckpt_path)

  info_specs = (utils.TensorSpec([], 'int64', 'episode_num_frames'),
                utils.TensorSpec([], 'float32', 'episode_returns'),
                utils.TensorSpec([], 'float32', 'episode_raw_returns'))
  info_queue = utils.StructuredFIFOQueue(-1, info_specs)
  world = step.world
  host = InferenceHost(agent, FLAGS.num_envs, FLAGS.unroll_length, FLAGS.inference_batch_size,
                       env.observation_space.shape, FLAGS.num_action_repeats, info_queue=info_queue)
  server = grpc.Server([FLAGS.server_address])
  server.bind(host.inference)
  server.start()

  last_ckpt_time = 0
  last_log, last_frames = time.time(), optimizer.iterations * iter_frame_ratio
  per_replica = FLAGS.batch_size // world                                   # :422
  try:
    while optimizer.iterations < final_iteration:                           # :467-476
      now = time.time()
      if now - last_ckpt_time >= FLAGS.save_checkpoint_secs:
        save()
        last_ckpt_time = now
      batch = dequeue_batch(host.unroll_queue, per_replica)
      loss, logs = step.minimize(batch)
      if optimizer.iterations % FLAGS.log_batch_frequency == 0:
        frames = optimizer.iterations * iter_frame_ratio
        dt = time.time() - last_log
        logging.info('step %d  speed/steps_per_sec %.1f  %s', optimizer.iterations,
                     (frames - last_frames) / max(dt, 1e-9),
                     {k: round(float(v), 5) for k, v in logs})
        last_log, last_frames = time.time(), frames
        n = info_queue.size()
        n -= n % FLAGS.log_episode_frequency
        if n:
          fr, ret, raw = info_queue.dequeue_many(n)
          logging.info('episode_return %.3f raw %.3f frames %.1f', float(ret.float().mean()),
                       float(raw.float().mean()), float(fr.float().mean()))
  finally:
    save()
    server.shutdown()
    host.unroll_queue.close()
