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

from seed_rl_b200 import _lib
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
               num_action_repeats=1, device='cuda', info_queue=None, training_batch_size=None,
               cuda_graph=None):
    """training_batch_size: when given, completed unrolls are gathered straight into the columns
    of preallocated time-major training batches (`self.assembler`, utils.BatchAssembler: zero-copy
    minibatch assembly); otherwise they go through the reference's capacity-1 `unroll_queue` of
    single unrolls and `dequeue_batch` stacks them."""
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
    self.assembler = None
    if training_batch_size:
      self.assembler = utils.BatchAssembler((action_specs, self.env_output_specs, agent_output_specs),
                                            agent_state_specs, unroll_length + 1, training_batch_size,
                                            slots=2, device=device)
    self.info_queue = info_queue
    N = self.N
    self.inference_specs = (
        TS([N], 'int32', 'env_id'), TS([N], 'int64', 'run_id'),
        utils.map_structure(lambda s: TS([N] + list(s.shape), s.dtype, s.name), self.env_output_specs),
        TS([N], 'float32', 'raw_reward'))
    self.output_specs = TS([N], 'int64', 'action')
    self.stream = torch.cuda.Stream(device=self.device) if self.device.type == 'cuda' else None
    # CUDA-graph replay of the device side of a full inference batch (default with the assembler):
    # gather -> T=1 forward -> sample -> write-back -> store append are ~40 small dependent launches
    # whose CPU issue cost, not their GPU time, bounded the step.
    self.use_graph = bool(cuda_graph if cuda_graph is not None else (self.assembler is not None))
    self._graph = None

    @grpc.function(self.inference_specs, self.output_specs)
    def inference(env_ids, run_ids, env_outputs, raw_rewards):
      return self._inference(env_ids, run_ids, env_outputs, raw_rewards)
    self.inference = inference

  # ---- CUDA-graph path ---------------------------------------------------------------------------
  def _device_step(self, ids32, env_dev, rng_counter):
    """Everything of one inference batch that runs on the device (learner.py:381-403), on static
    buffers: captured once, replayed per batch."""
    n = int(ids32.numel())
    prev_actions = torch.empty([n], dtype=torch.int64, device=self.device)
    prev_states = tuple(torch.empty([n, networks.LSTM_UNITS], dtype=torch.float32, device=self.device)
                        for _ in range(2))
    _lib.rows_multi([(self.actions._state[0], prev_actions, _lib.ROW_GATHER),
                     (self.agent_states._state[0], prev_states[0], _lib.ROW_GATHER),
                     (self.agent_states._state[1], prev_states[1], _lib.ROW_GATHER)], ids32)
    agent_outputs, curr_states = self.agent(prev_actions, env_dev, prev_states, is_training=False,
                                            rng_counter=rng_counter)
    self.store.device_append(ids32, utils.flatten((prev_actions, env_dev, agent_outputs)))
    _lib.rows_multi([(self.agent_states._state[0], curr_states[0].contiguous(), _lib.ROW_SCATTER),
                     (self.agent_states._state[1], curr_states[1].contiguous(), _lib.ROW_SCATTER),
                     (self.actions._state[0], agent_outputs.action.contiguous(), _lib.ROW_SCATTER)], ids32)
    return prev_states, agent_outputs

  def _build_graph(self):
    N, dev = self.N, self.device
    dt = utils.as_torch_dtype
    self._g_ids = torch.zeros([N], dtype=torch.int32, device=dev)
    self._g_env = utils.EnvOutput(*(torch.zeros([N] + list(s.shape), dtype=dt(s.dtype), device=dev)
                                    for s in self.env_output_specs))
    self._g_pin = [torch.zeros_like(t, device='cpu').pin_memory() for t in (self._g_ids,) + tuple(self._g_env)]
    self._g_counter = torch.zeros([], dtype=torch.int64, device=dev)
    with torch.cuda.stream(self.stream):
      self._g_ids.copy_(torch.arange(N, dtype=torch.int32))       # distinct ids for the warm-up / capture
      # warm-up on scratch copies of the mutable tables is not needed: the capture run below does
      # not execute, and the one eager warm-up is undone by restoring the tables it touches
      saved = [t.clone() for t in self.actions._state + self.agent_states._state + self.store._state] + \
              [self.store._index.clone()]
      self._device_step(self._g_ids, self._g_env, self._g_counter)   # eager: lazy inits (func attributes, workspaces)
      for t, sv in zip(self.actions._state + self.agent_states._state + self.store._state + [self.store._index], saved):
        t.copy_(sv)
      self._g_counter.zero_()
      self.stream.synchronize()
      g = torch.cuda.CUDAGraph()
      # thread_local: the learner thread may allocate / launch on its own stream during the capture
      with torch.cuda.graph(g, stream=self.stream, capture_error_mode='thread_local'):
        self._g_prev_states, self._g_out = self._device_step(self._g_ids, self._g_env, self._g_counter)
      self._g_actions_pin = torch.zeros([N], dtype=torch.int64).pin_memory()
    self._graph = g

  def _inference_graph(self, env_ids, run_ids, env_outputs, raw_rewards):
    """reference learner.py:351-405 with the device side as one graph replay."""
    reward, done = np.asarray(env_outputs.reward), np.asarray(env_outputs.done)
    previous = self.env_run_ids[env_ids]
    self.env_run_ids[env_ids] = run_ids
    reset_ids = env_ids[previous != run_ids]
    if np.asarray(env_outputs.abandoned).any():                         # :368-370
      raise ValueError('Abandoned done states are not supported in VTRACE.')
    utils._check_no_duplicates(None, env_ids, 'inference batch')
    with torch.cuda.stream(self.stream):
      if self._graph is None:
        self._build_graph()
      if reset_ids.size:                                                # :353-366 (rare: eager)
        logging.info('Environment ids needing reset: %s', reset_ids)
        for t in self.env_infos:
          t[reset_ids] = 0
        self.store.reset(reset_ids)
        init = self.agent.initial_state(len(reset_ids))
        self.first_agent_states.replace(reset_ids, init)
        self.agent_states.replace(reset_ids, init)
        self.actions.reset(reset_ids)
      self.env_infos[1][env_ids] += reward                              # :373-378
      self.env_infos[2][env_ids] += np.asarray(raw_rewards)
      done_ids = env_ids[done]
      if self.info_queue is not None and done_ids.size:
        self.info_queue.enqueue_many(tuple(torch.as_tensor(t[done_ids]) for t in self.env_infos))
      for t in self.env_infos:
        t[done_ids] = 0
      self.env_infos[0][env_ids] += self.num_action_repeats
      # inputs: host arrays -> pinned staging -> the graph's static device buffers
      srcs = (env_ids.astype(np.int32),) + tuple(np.asarray(x) for x in env_outputs)
      for pin, dst, src in zip(self._g_pin, (self._g_ids,) + tuple(self._g_env), srcs):
        t = torch.from_numpy(np.ascontiguousarray(src))
        if t.numel() >= 65536 and t.is_pinned():
          dst.copy_(t, non_blocking=True)        # the batcher's slabs are pinned: DMA straight from them
        else:
          pin.numpy()[...] = src                 # small fields / pageable memory: own pinned staging
          dst.copy_(pin, non_blocking=True)
      self._graph.replay()
      # completed unrolls (known on the host) -> columns of the training batch; first agent states
      done_host, pos = self.store.host_advance(env_ids)
      if done_host.size:
        pos_dev = torch.as_tensor(pos.astype(np.int64)).to(self.device, non_blocking=True)
        def on_placed(slot, col0, ids):
          first = self.first_agent_states.read(ids.to(torch.int64))
          for dst, src in zip(self.assembler._states[slot], first):
            dst[col0:col0 + int(ids.numel())].copy_(src)
        completed_ids, _ = self.store.complete_into(int(done_host.size), self.assembler, on_placed)
        # the state the next unroll starts from = the state this step started from (:400-401)
        self.first_agent_states.replace(completed_ids, tuple(s.index_select(0, pos_dev) for s in self._g_prev_states),
                                        check_unique=False)
      self._g_actions_pin.copy_(self._g_out.action, non_blocking=True)
      self.stream.synchronize()
    return self._g_actions_pin.numpy().copy()

  def _inference(self, env_ids, run_ids, env_outputs, raw_rewards):
    """reference learner.py:351-405."""
    env_ids = np.asarray(env_ids); run_ids = np.asarray(run_ids)
    if self.use_graph and self.assembler is not None and len(env_ids) == self.N:
      return self._inference_graph(env_ids, run_ids, env_outputs, raw_rewards)
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
      ids32 = ids_dev.to(torch.int32)
      env_dev = utils.EnvOutput(*(torch.as_tensor(np.asarray(x)).to(self.device, non_blocking=True)
                                  for x in env_outputs))
      # previous action + recurrent state of these environments: ONE gather launch (:381-383)
      n = int(ids32.numel())
      prev_actions = torch.empty([n], dtype=torch.int64, device=self.device)
      prev_states = tuple(torch.empty([n, networks.LSTM_UNITS], dtype=torch.float32, device=self.device)
                          for _ in range(2))
      _lib.rows_multi([(self.actions._state[0], prev_actions, _lib.ROW_GATHER),
                       (self.agent_states._state[0], prev_states[0], _lib.ROW_GATHER),
                       (self.agent_states._state[1], prev_states[1], _lib.ROW_GATHER)], ids32)
      agent_outputs, curr_states = self.agent(prev_actions, env_dev, prev_states, is_training=False)
      # Append to the unroll store, enqueue completed unrolls (:394-399).
      pending = []
      if self.assembler is not None:
        def on_placed(slot, col0, ids):
          first = self.first_agent_states.read(ids.to(torch.int64))
          for dst, src in zip(self.assembler._states[slot], first):
            dst[col0:col0 + int(ids.numel())].copy_(src)
        completed_ids, placed = self.store.append(env_ids, (prev_actions, env_dev, agent_outputs),
                                                  check_duplicates=True, into=self.assembler,
                                                  on_placed=on_placed)
        n_done = 0
        if placed:
          self.first_agent_states.replace(completed_ids, self.agent_states.read(completed_ids), check_unique=False)
      else:
        completed_ids, unrolls = self.store.append(env_ids, (prev_actions, env_dev, agent_outputs),
                                                   check_duplicates=True)
        n_done = int(completed_ids.numel())
      if n_done:
        first = self.first_agent_states.read(completed_ids)
        flat = utils.flatten(unrolls)
        for i in range(n_done):     # one queue element per unroll, as in the reference
          u = utils.pack_sequence_as(self.store._specs, [f[:, i] for f in flat])
          pending.append(Unroll((first[0][i], first[1][i]), *u))
        self.first_agent_states.replace(completed_ids, self.agent_states.read(completed_ids), check_unique=False)
      # Update current state (:402-403) and return the actions (:405).
      # (ids are unique: UnrollStore.append checked them)  ONE scatter launch
      _lib.rows_multi([(self.agent_states._state[0], curr_states[0].contiguous(), _lib.ROW_SCATTER),
                       (self.agent_states._state[1], curr_states[1].contiguous(), _lib.ROW_SCATTER),
                       (self.actions._state[0], agent_outputs.action.contiguous(), _lib.ROW_SCATTER)], ids32)
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


def save_checkpoint(path, agent, optimizer, extra=None):
  """tf.train.CheckpointManager.save analogue (reference learner.py:283-296,470-476): one
  `torch.save` blob {agent: flat fp32 arena + tensor table, optimizer: Adam m / v / iterations}
  written atomically.  NOT interchangeable with the reference's tf.train.Checkpoint files (no
  TensorFlow here); `named_parameters()` gives the Keras-layout tensors by name for conversion."""
  blob = {'agent': agent.state_dict(), 'optimizer': optimizer.state_dict(), 'format': 'seed_rl_b200/1'}
  if extra:
    blob.update(extra)
  torch.save(blob, path + '.tmp')
  os.replace(path + '.tmp', path)


def restore_checkpoint(path, agent, optimizer):
  """ckpt.restore(...).assert_consumed() analogue: raises on a tensor-table mismatch."""
  d = torch.load(path, map_location='cpu', weights_only=False)
  info = d['agent'].get('param_info')
  if info is not None and [tuple(x) for x in info] != [tuple(x) for x in agent.param_info]:
    raise ValueError('checkpoint %s was written by a different network (tensor table mismatch)' % path)
  agent.load_state_dict(d['agent'])
  optimizer.load_state_dict(d['optimizer'], device=str(agent.device))
  return d


def rank_server_address(address, rank):
  """One server per replica (reference: one per host, learner.py:339-347): replica r binds the
  flag's address with its port (or unix-socket path) offset by r."""
  if rank == 0:
    return address
  if address.startswith('unix:'):
    return '%s.%d' % (address, rank)
  host, _, port = address.rpartition(':')
  return '%s:%d' % (host, int(port) + rank)


def init_replicas():
  """torchrun starts one learner process per GPU; returns (rank, world)."""
  import torch.distributed as td
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  if world > 1 and not td.is_initialized():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if torch.cuda.is_available():
      torch.cuda.set_device(local)
      td.init_process_group('nccl', device_id=torch.device('cuda', local))
    else:
      td.init_process_group('gloo')
  return rank, world


def assembled_batch(assembler):
  """The zero-copy counterpart of `dequeue_batch`: the next full time-major batch of the
  assembler as an `Unroll` of views (no copy).  Returns (slot, Unroll); call
  assembler.release(slot) after the step that consumes it has been enqueued."""
  slot, state, (prev_actions, env_outputs, agent_outputs) = assembler.get()
  return slot, Unroll(tuple(state), prev_actions, utils.EnvOutput(*env_outputs),
                      networks.AgentOutput(*agent_outputs))


def learner_loop(create_env_fn, create_agent_fn, create_optimizer_fn):
  """reference learner.py:170-483 (one replica per process / GPU)."""
  logging.info('Starting learner loop')
  utils.validate_learner_config(FLAGS)
  rank, world = init_replicas()
  env = create_env_fn(0, FLAGS)
  dist = get_parametric_distribution_for_action_space(env.action_space)
  agent = create_agent_fn(env.action_space, env.observation_space, dist)
  if not hasattr(agent, '_entropy_mul'):
    agent.init_entropy_cost(FLAGS.entropy_cost, FLAGS.entropy_cost_adjustment_speed)
  iter_frame_ratio = FLAGS.batch_size * FLAGS.unroll_length * FLAGS.num_action_repeats
  final_iteration = int(math.ceil(FLAGS.total_environment_frames / iter_frame_ratio))
  optimizer, learning_rate_fn = create_optimizer_fn(final_iteration)
  settings = learner_lib.loss_settings_from_flags()
  # summaries + periodic progress export (learner.py:236-237,280,286,447-465)
  os.makedirs(FLAGS.logdir, exist_ok=True)
  summary_writer = utils.SummaryWriter(FLAGS.logdir) if rank == 0 else None
  logger = utils.ProgressLogger(summary_writer=summary_writer,
                                starting_step=optimizer.iterations * iter_frame_ratio)
  step = learner_lib.LearnerStep(agent, optimizer, dist, settings, logger=logger, grad_reduce=FLAGS.grad_reduce)

  ckpt_path = os.path.join(FLAGS.logdir, 'ckpt.pt')
  init = FLAGS.init_checkpoint or (ckpt_path if os.path.exists(ckpt_path) else None)
  if init:                                                                 # :286-296
    logging.info('Restoring checkpoint: %s', init)
    restore_checkpoint(init, agent, optimizer)
    logger.reset(summary_writer, optimizer.iterations * iter_frame_ratio)

  def save():
    if rank == 0:                      # replicas are bit-identical; one writer, no os.replace race
      save_checkpoint(ckpt_path, agent, optimizer)

  info_specs = (utils.TensorSpec([], 'int64', 'episode_num_frames'),
                utils.TensorSpec([], 'float32', 'episode_returns'),
                utils.TensorSpec([], 'float32', 'episode_raw_returns'))
  info_queue = utils.StructuredFIFOQueue(-1, info_specs)
  assert step.world == world
  per_replica = FLAGS.batch_size // world                                   # :422
  host = InferenceHost(agent, FLAGS.num_envs, FLAGS.unroll_length, FLAGS.inference_batch_size,
                       env.observation_space.shape, FLAGS.num_action_repeats, info_queue=info_queue,
                       training_batch_size=per_replica)
  server = grpc.Server([rank_server_address(FLAGS.server_address, rank)])
  server.bind(host.inference)
  server.start()

  def additional_logs():                                                    # :447-463
    if summary_writer:
      summary_writer.scalar('learning_rate', learning_rate_fn(optimizer.iterations))
    n = info_queue.size()
    n -= n % FLAGS.log_episode_frequency
    if n:
      stats = info_queue.dequeue_many(n)
      for key, values in zip(('episode_num_frames', 'episode_return', 'episode_raw_return'), stats):
        for chunk in torch.split(values.float(), FLAGS.log_episode_frequency):
          if summary_writer:
            summary_writer.scalar(key, float(chunk.mean()))
      for fr, ret, raw in zip(*stats):
        logging.info('Return: %f Raw return: %f Frames: %i', float(ret), float(raw), int(fr))

  logger.start(additional_logs)
  last_ckpt_time = 0
  try:
    while optimizer.iterations < final_iteration:                           # :467-476
      now = time.time()
      if now - last_ckpt_time >= FLAGS.save_checkpoint_secs:
        save()
        last_ckpt_time = now
      slot, batch = assembled_batch(host.assembler)
      _, logs = step.minimize(batch)
      host.assembler.release(slot)
      logger.step_end(logs, None, iter_frame_ratio)                         # :280
  finally:
    logger.shutdown()
    save()
    server.shutdown()
    host.unroll_queue.close()
    host.assembler.close()
    if summary_writer:
      summary_writer.close()
