"""V-trace learner -- mirror of the reference's `agents/vtrace/learner.py`:

  flags                         learner.py:38-68   (same names and defaults)
  compute_loss                  learner.py:73-159
  Unroll                        learner.py:162-163
  minimize (LearnerStep)        learner.py:255-280
  learner_loop                  learner.py:170-483 (see learner_loop.py for the RPC wiring)

Device work per step = seedrl_net_forward -> seedrl_vtrace_loss_fwd_bwd ->
seedrl_net_backward -> [one NCCL all-reduce(SUM) of the flat gradient arena]
-> seedrl_adam_apply.  There is no autograd tape: the loss kernel emits the analytic
gradient w.r.t. the network outputs and the network backward is an explicit schedule.
"""
import collections
import math

from absl import flags
import torch

from seed_rl_b200 import _lib
from seed_rl_b200.common import common_flags  # pylint: disable=unused-import
from seed_rl_b200.common import utils

# Training.
common_flags.define_once(flags.DEFINE_integer, 'save_checkpoint_secs', 1800, 'Checkpoint save period in seconds.')
common_flags.define_once(flags.DEFINE_integer, 'total_environment_frames', int(1e9),
                     'Total environment frames to train for.')
common_flags.define_once(flags.DEFINE_integer, 'batch_size', 32, 'Batch size for training.')
common_flags.define_once(flags.DEFINE_integer, 'inference_batch_size', -1, 'Batch size for inference, -1 for auto-tune.')
common_flags.define_once(flags.DEFINE_integer, 'unroll_length', 100, 'Unroll length in agent steps.')
common_flags.define_once(flags.DEFINE_integer, 'num_training_tpus', 1, 'Unused on B200 (kept for flag compatibility).')
common_flags.define_once(flags.DEFINE_string, 'init_checkpoint', None,
                    'Path to the checkpoint used to initialize the agent.')
# Loss settings.
common_flags.define_once(flags.DEFINE_float, 'entropy_cost', 0.00025, 'Entropy cost/multiplier.')
common_flags.define_once(flags.DEFINE_float, 'target_entropy', None, 'If not None, the entropy cost is '
                   'automatically adjusted to reach the desired entropy level.')
common_flags.define_once(flags.DEFINE_float, 'entropy_cost_adjustment_speed', 10., 'Controls how fast '
                   'the entropy cost coefficient is adjusted.')
common_flags.define_once(flags.DEFINE_float, 'baseline_cost', .5, 'Baseline cost/multiplier.')
common_flags.define_once(flags.DEFINE_float, 'kl_cost', 0., 'KL(old_policy|new_policy) loss multiplier.')
common_flags.define_once(flags.DEFINE_float, 'discounting', .99, 'Discounting factor.')
common_flags.define_once(flags.DEFINE_float, 'lambda_', 1., 'Lambda.')
common_flags.define_once(flags.DEFINE_float, 'max_abs_reward', 0., 'Maximum absolute reward when calculating loss.'
                   'Use 0. to disable clipping.')
# Logging
common_flags.define_once(flags.DEFINE_integer, 'log_batch_frequency', 100, 'We average that many batches '
                     'before logging batch statistics like entropy.')
common_flags.define_once(flags.DEFINE_integer, 'log_episode_frequency', 1, 'We average that many episodes'
                     ' before logging average episode return and length.')
# B200 additions
common_flags.define_once(flags.DEFINE_enum, 'grad_reduce', 'sum', ['sum', 'mean'],
                  'Cross-replica gradient reduction. The reference SUMs '
                  '(tests/utils_test.py:609-650).')

FLAGS = flags.FLAGS

LossSettings = collections.namedtuple(
    'LossSettings',
    'discounting lambda_ baseline_cost entropy_cost kl_cost max_abs_reward '
    'target_entropy entropy_cost_adjustment_speed')


def loss_settings_from_flags():
  return LossSettings(FLAGS.discounting, FLAGS.lambda_, FLAGS.baseline_cost,
                      FLAGS.entropy_cost, FLAGS.kl_cost, FLAGS.max_abs_reward,
                      FLAGS.target_entropy, FLAGS.entropy_cost_adjustment_speed)


def default_loss_settings(**kw):
  d = dict(discounting=.99, lambda_=1., baseline_cost=.5, entropy_cost=0.00025, kl_cost=0.,
           max_abs_reward=0., target_entropy=None, entropy_cost_adjustment_speed=10.)
  d.update(kw)
  return LossSettings(**d)


class _NullLogger(object):
  def log_session(self):
    return []

  def log(self, session, name, value):
    session.append((name, value))


_LOG_NAMES = [  # learner.py:138-157
    ('V/value function', 'v_mean'), ('V/L2 error', 'v_l2_error'),
    ('losses/policy', 'policy'), ('losses/V', 'V'), ('losses/entropy', 'entropy'),
    ('losses/kl', 'kl'), ('losses/total', 'total'),
    ('policy/max_action_abs(before_tanh)', 'max_action_abs'),
    ('policy/entropy', 'mean_entropy'), ('policy/entropy_cost', 'entropy_cost'),
    ('policy/kl(old|new)', 'mean_kl')]

_scratch_cache = {}


def _loss_scratch(T1, B, A, device):
  key = (T1, B, A, str(device))
  if key not in _scratch_cache:
    n = int(_lib.lib().seedrl_vtrace_loss_scratch_bytes(T1, B, A))
    _scratch_cache[key] = torch.zeros(n, dtype=torch.uint8, device=device)   # zeroed ONCE
  return _scratch_cache[key]


def vtrace_loss_fwd_bwd(settings, learner_logits, learner_baseline, behaviour_logits,
                        actions, rewards, done, entropy_cost_param, want_vtrace=False):
  """The fused kernel of compute_loss (learner.py:82-157) + its gradient.  All inputs
  have T+1 rows.  Returns dict(loss_terms[16], dlogits, dbaseline, d_entropy_cost_param,
  vs, pg_advantages)."""
  f32 = torch.float32
  ll = _lib.require_cuda(learner_logits, f32, 'learner_logits')
  lb = _lib.require_cuda(learner_baseline, f32, 'learner_baseline')
  bl = _lib.require_cuda(behaviour_logits, f32, 'behaviour_logits')
  act = _lib.require_cuda(actions, torch.int64, 'actions')
  rew = _lib.require_cuda(rewards, f32, 'rewards')
  dn = _lib.require_cuda(done, torch.bool, 'done')
  ecp = _lib.require_cuda(entropy_cost_param, f32, 'entropy_cost_param')
  if ll.dim() != 3 or lb.dim() != 2:
    raise ValueError('learner outputs must be [T+1,B,A] and [T+1,B]')
  T1, B, A = (int(x) for x in ll.shape)
  for t, shp, nm in ((lb, (T1, B), 'learner_baseline'), (bl, (T1, B, A), 'behaviour_logits'),
                     (act, (T1, B), 'actions'), (rew, (T1, B), 'rewards'), (dn, (T1, B), 'done')):
    if tuple(t.shape) != shp:
      raise ValueError('%s has shape %s, expected %s' % (nm, tuple(t.shape), shp))
  dev = ll.device
  cfg = _lib.LossConfig(
      settings.discounting, settings.lambda_, settings.baseline_cost, settings.kl_cost,
      settings.max_abs_reward or 0.0, 1.0, 1.0,     # compute_loss uses vtrace's default clips
      settings.target_entropy or 0.0, 1 if settings.target_entropy else 0,
      settings.entropy_cost_adjustment_speed)
  out = dict(
      loss_terms=torch.empty(_lib.LOSS_TERMS, dtype=f32, device=dev),
      dlogits=torch.empty_like(ll), dbaseline=torch.empty_like(lb),
      d_entropy_cost_param=torch.empty((), dtype=f32, device=dev),
      vs=torch.empty([T1 - 1, B], dtype=f32, device=dev) if want_vtrace else None,
      pg_advantages=torch.empty([T1 - 1, B], dtype=f32, device=dev) if want_vtrace else None)
  import ctypes
  _lib.check(_lib.lib().seedrl_vtrace_loss_fwd_bwd(
      T1, B, A, _lib.ptr(ll), _lib.ptr(lb), _lib.ptr(bl), _lib.ptr(act), _lib.ptr(rew),
      _lib.ptr(dn), ctypes.byref(cfg), _lib.ptr(ecp), _lib.ptr(out['loss_terms']),
      _lib.ptr(out['dlogits']), _lib.ptr(out['dbaseline']),
      _lib.ptr(out['d_entropy_cost_param']), _lib.ptr(out['vs']),
      _lib.ptr(out['pg_advantages']), _lib.ptr(_loss_scratch(T1, B, A, dev)),
      _lib.stream_ptr()))
  return out


def compute_loss(logger, parametric_action_distribution, agent, agent_state,
                 prev_actions, env_outputs, agent_outputs, settings=None):
  """reference learner.py:73-159.  Returns (total_loss, log session).  The gradient of
  total_loss w.r.t. the network outputs is left on the agent for `minimize`."""
  settings = settings or loss_settings_from_flags()
  learner_outputs, _ = agent(prev_actions, env_outputs, agent_state,
                             unroll=True, is_training=True)                 # :75-79
  r = vtrace_loss_fwd_bwd(settings, learner_outputs.policy_logits, learner_outputs.baseline,
                          agent_outputs.policy_logits, agent_outputs.action,
                          env_outputs[0], env_outputs[1], agent.entropy_cost_param)
  agent._loss_grads = r
  logger = logger or _NullLogger()
  session = logger.log_session()
  lt = r['loss_terms']
  for name, key in _LOG_NAMES:
    logger.log(session, name, lt[_lib.LT[key]])
  return lt[_lib.LT['total']], session


Unroll = collections.namedtuple(
    'Unroll', 'agent_state prev_actions env_outputs agent_outputs')


def reduce_gradients(flat_grads, world, process_group=None, grad_reduce='sum'):
  """The ONE exchange step of a data-parallel iteration: all-reduce(SUM) of the flat
  gradient arena in place (NCCL over NVLink on GPUs, gloo in the CPU tests).  Returns the
  scale Adam must apply to the reduced gradient: 1 for the reference's semantics --
  every replica's *mean*-loss gradient is SUMMED across replicas
  (reference tests/utils_test.py:609-650: `expected_a = 1 - N*0.2`) -- or 1/world for
  `grad_reduce='mean'`."""
  if grad_reduce not in ('sum', 'mean'):
    raise ValueError('grad_reduce must be "sum" or "mean"')
  if world <= 1:
    return 1.0
  import torch.distributed as td
  td.all_reduce(flat_grads, op=td.ReduceOp.SUM, group=process_group)
  return 1.0 / world if grad_reduce == 'mean' else 1.0


def env_shard(rank, world, num_envs):
  """Environments owned by replica `rank`: {i : i mod world == rank}."""
  return list(range(rank, num_envs, world))


class LearnerStep(object):
  """`minimize` of reference learner.py:255-280 for one replica (= one GPU/process)."""

  def __init__(self, agent, optimizer, parametric_action_distribution=None, settings=None,
               logger=None, process_group=None, grad_reduce='sum', check_errors_every=64,
               overlap_reduce=True):
    self.agent = agent
    self.overlap_reduce = overlap_reduce
    self._side = self._head_ev = self._head_work = None
    # the kernels' bounded-wait error flag is polled (one 4-byte D2H + stream sync) every
    # `check_errors_every` steps; 0 = never (the caller polls agent.check_errors() itself)
    self.check_errors_every = int(check_errors_every)
    self._steps = 0
    self.optimizer = optimizer
    self.dist = parametric_action_distribution
    self.settings = settings or default_loss_settings()
    self.logger = logger
    self.pg = process_group
    self.grad_reduce = grad_reduce
    import torch.distributed as td
    self.world = td.get_world_size(process_group) if (td.is_available() and td.is_initialized()) else 1
    if not hasattr(agent, '_entropy_mul'):
      agent.init_entropy_cost(self.settings.entropy_cost,
                              self.settings.entropy_cost_adjustment_speed)       # :225-234
    optimizer._create_slots(agent.params)                                        # :244-245
    self.last_loss_terms = None

  def compute_gradients(self, unroll):
    loss, logs = compute_loss(self.logger, self.dist, self.agent, unroll.agent_state,
                              unroll.prev_actions, unroll.env_outputs, unroll.agent_outputs,
                              self.settings)
    r = self.agent._loss_grads
    self._head_work = None
    if self.world > 1 and self.overlap_reduce and torch.cuda.is_available():
      # bucket 1 (heads, Dense, LSTM = 94 % of the arena) is all-reduced on a side stream while
      # the convolution torso's backward still runs; bucket 2 follows in apply_gradients
      import torch.distributed as td
      if self._side is None:
        self._side, self._head_ev = torch.cuda.Stream(), torch.cuda.Event()
      grads = self.agent.backward(r['dlogits'], r['dbaseline'], head_ready_event=self._head_ev)
      with torch.cuda.stream(self._side):
        self._side.wait_event(self._head_ev)
        self._head_work = td.all_reduce(grads[:self.agent.grad_split], op=td.ReduceOp.SUM, group=self.pg,
                                        async_op=True)
    else:
      grads = self.agent.backward(r['dlogits'], r['dbaseline'])                   # :264
    grads[self.agent.entropy_cost_param_index] = r['d_entropy_cost_param']
    self.last_loss_terms = r['loss_terms']
    return loss, logs

  def apply_gradients(self):
    grads = self.agent.grads
    if getattr(self, '_head_work', None) is not None:
      import torch.distributed as td
      td.all_reduce(grads[self.agent.grad_split:], op=td.ReduceOp.SUM, group=self.pg)
      self._head_work.wait()              # the compute stream waits for the side-stream bucket
      self._head_work = None
      scale = 1.0 / self.world if self.grad_reduce == 'mean' else 1.0
    else:
      scale = reduce_gradients(grads, self.world, self.pg, self.grad_reduce)
    mul = self.settings.entropy_cost_adjustment_speed
    self.optimizer.apply_gradients(
        self.agent.params, grads, grad_scale=scale,
        clamp_index=self.agent.entropy_cost_param_index,
        clamp_lo=-20.0 / mul, clamp_hi=20.0 / mul)                                # :229-231

  def minimize(self, unroll):
    loss, logs = self.compute_gradients(unroll)
    self.apply_gradients()
    self._steps += 1
    if self.check_errors_every and (self._steps == 1 or self._steps % self.check_errors_every == 0):
      self.agent.check_errors()
    return loss, logs


class DeviceFeeder(object):
  """Double-buffered host -> device feed of training batches (the role tf.data prefetching
  onto the accelerator plays for the reference's `minimize(it)`, learner.py:457-466).

  `put(pinned)` enqueues the H2D copies of one batch (a dict of pinned host tensors) on a
  private copy stream into the slot the learner is not using; `get()` returns the dict of
  device tensors of the oldest pending batch and makes the current (compute) stream wait for
  its copies.  While step i trains from one slot, batch i+1 streams into the other, so the
  upload is hidden behind the step instead of serialised in front of it."""

  def __init__(self, example, device='cuda', slots=2):
    self._slots = [{k: torch.empty_like(v, device=device) for k, v in example.items()} for _ in range(slots)]
    self._copied = [torch.cuda.Event() for _ in range(slots)]
    self._consumed = [None] * slots
    self._stream = torch.cuda.Stream(device=device)
    self._put = 0
    self._get = 0

  @property
  def slots(self):
    return self._slots

  def put(self, pinned):
    if self._put - self._get >= len(self._slots):
      raise RuntimeError('DeviceFeeder: every slot holds an unconsumed batch')
    s = self._put % len(self._slots)
    with torch.cuda.stream(self._stream):
      if self._consumed[s] is not None:
        self._stream.wait_event(self._consumed[s])     # the step that read this slot has finished
      for k, v in pinned.items():
        self._slots[s][k].copy_(v, non_blocking=True)
      self._copied[s].record(self._stream)
    self._put += 1

  def get(self):
    if self._get >= self._put:
      raise RuntimeError('DeviceFeeder: no batch pending')
    s = self._get % len(self._slots)
    torch.cuda.current_stream().wait_event(self._copied[s])
    self._get += 1
    return s, self._slots[s]

  def done_with(self, slot):
    """Call after the step that consumed `slot` has been enqueued on the compute stream."""
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    self._consumed[slot] = ev
