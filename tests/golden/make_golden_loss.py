"""Generates tests/golden/loss_golden.npz.  Run ONLY in the build container (where
/root/reference exists):   python tests/golden/make_golden_loss.py

Executes the UNMODIFIED reference `compute_loss` (agents/vtrace/learner.py:73-159, pulled out
of the file by AST) over tf_numpy_shim, with the UNMODIFIED common/vtrace.py behind it, a stub
agent that returns given learner outputs, and a categorical distribution stub with TFP's
published log_prob / entropy semantics (log-softmax gather; -sum p log p).  This pins the
COMPOSITION of the loss -- which rows are dropped, reward clipping, discounts, the five terms
and their weights, the logged scalars and their names -- to the reference source rather than
to our reading of it.  (Gradients are not produced: the shim has no tape.)"""
import collections
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402
from make_golden import _extract_function, _load  # noqa: E402

AgentOutput = collections.namedtuple('AgentOutput', 'action policy_logits baseline')
EnvOutput = collections.namedtuple('EnvOutput', 'reward done observation abandoned episode_step')


def main():
  tf = tf_numpy_shim.install()
  T, raw = tf_numpy_shim.Tensor, tf_numpy_shim._raw
  f32 = np.float32

  def map_structure(fn, *structs):
    s0 = structs[0]
    if isinstance(s0, tuple) and hasattr(s0, '_fields'):
      return type(s0)(*[map_structure(fn, *[getattr(s, f) for s in structs]) for f in s0._fields])
    if isinstance(s0, (tuple, list)):
      return type(s0)(map_structure(fn, *xs) for xs in zip(*structs))
    return fn(*structs)
  tf.nest = types.ModuleType('nest'); tf.nest.map_structure = map_structure
  tf.clip_by_value = lambda x, lo, hi: T(np.clip(raw(x), f32(lo), f32(hi)))
  tf.reduce_mean = lambda x, axis=None: T(np.mean(raw(x), axis=axis, dtype=f32))
  tf.reduce_max = lambda x, axis=None: T(np.max(raw(x), axis=axis))
  tf.square = lambda x: T(np.square(raw(x)))
  tf.sqrt = lambda x: T(np.sqrt(raw(x)))
  tf.abs = lambda x: T(np.abs(raw(x)))
  ref_vtrace = _load(os.path.join(REF, 'common/vtrace.py'), 'ref_vtrace')

  class Dist(object):            # TFP Categorical semantics (pinned for log_prob by tests/vtrace_test.py:88-115)
    @staticmethod
    def _lsm(logits):
      l = np.asarray(raw(logits), f32)
      m = l.max(-1, keepdims=True)
      return (l - m - np.log(np.exp(l - m).sum(-1, keepdims=True, dtype=f32))).astype(f32)

    def log_prob(self, logits, actions):
      lsm = self._lsm(logits)
      a = np.asarray(raw(actions))
      return T(np.take_along_axis(lsm, a[..., None], -1)[..., 0])

    def entropy(self, logits):
      lsm = self._lsm(logits)
      return T(-(np.exp(lsm) * lsm).sum(-1, dtype=f32))

    def create_dist(self, logits):
      return object()

  class Logger(object):
    def log_session(self): return []
    def log(self, session, key, value): session.append((key, np.asarray(raw(value))))

  out = {}
  rng = np.random.default_rng(11)
  cases = {'a': dict(), 'b': dict(kl_cost=0.3, entropy_cost=0.01, max_abs_reward=1.0, target_entropy=1.5, lambda_=0.9),
           'c': dict(discounting=0.9, baseline_cost=1.0)}
  for name, kw in cases.items():
    cfg = dict(discounting=0.99, lambda_=1.0, baseline_cost=0.5, entropy_cost=0.00025, kl_cost=0.0,
               max_abs_reward=0.0, target_entropy=None)
    cfg.update(kw)
    T1, B, A = 9, 4, 6
    ll = rng.normal(size=(T1, B, A)).astype(f32); lb = rng.normal(size=(T1, B)).astype(f32)
    bl = rng.normal(size=(T1, B, A)).astype(f32); act = rng.integers(0, A, (T1, B))
    rew = (rng.normal(size=(T1, B)) * 2).astype(f32); done = rng.random((T1, B)) < 0.2

    class Agent(object):
      def __call__(self, prev_actions, env_outputs, agent_state, unroll=True, is_training=True):
        return AgentOutput(T(act), T(ll), T(lb)), None

      def entropy_cost(self):
        return T(f32(cfg['entropy_cost']))
    flags = types.SimpleNamespace(**cfg)
    ns = {'tf': tf, 'FLAGS': flags, 'vtrace': ref_vtrace}
    compute_loss = _extract_function(os.path.join(REF, 'agents/vtrace/learner.py'), 'compute_loss', ns)
    env = EnvOutput(T(rew), T(done), T(np.zeros((T1, B), f32)), T(np.zeros((T1, B), bool)), T(np.zeros((T1, B), np.int32)))
    total, session = compute_loss(Logger(), Dist(), Agent(), None, T(act), env, AgentOutput(T(act), T(bl), T(np.zeros((T1, B), f32))))
    out.update({'%s_ll' % name: ll, '%s_lb' % name: lb, '%s_bl' % name: bl, '%s_act' % name: act, '%s_rew' % name: rew,
                '%s_done' % name: done, '%s_total' % name: np.asarray(raw(total)),
                '%s_cfg' % name: np.asarray([cfg['discounting'], cfg['lambda_'], cfg['baseline_cost'], cfg['entropy_cost'],
                                             cfg['kl_cost'], cfg['max_abs_reward'], cfg['target_entropy'] or 0.0])})
    for k, v in session:
      out['%s_log_%s' % (name, k.replace('/', '__'))] = v
  np.savez_compressed(os.path.join(HERE, 'loss_golden.npz'), **out)
  print('wrote loss_golden.npz:', sorted(k for k in out if k.startswith('a_')))


if __name__ == '__main__':
  main()
