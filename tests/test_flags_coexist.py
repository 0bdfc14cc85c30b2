"""The reference's V-trace and R2D2 learners are separate binaries that re-use absl flag names
(batch_size, unroll_length, discounting, save_checkpoint_secs ...).  Both mirrors must be importable
into one process in either order (the first definition of a name stands) and the R2D2 defaults must
not depend on that order."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = '''
import sys
sys.path.insert(0, %r)
from seed_rl_b200.agents.%s import learner as first
from seed_rl_b200.agents.%s import learner as second
from seed_rl_b200.agents.r2d2 import learner as r2d2
from seed_rl_b200.agents.vtrace import learner_loop
s = r2d2.default_settings()
assert (s.batch_size, s.burn_in, s.unroll_length, s.n_steps) == (64, 40, 100, 5), s
assert abs(s.discounting - 0.997) < 1e-9 and s.replay_buffer_size == 100
assert r2d2.get_replay_insertion_batch_size(s) == 42          # int(64 / 1.5), learner.py:115-119
print("ok")
'''


def _run(a, b):
  r = subprocess.run([sys.executable, '-c', CODE % (ROOT, a, b)], capture_output=True, text=True, timeout=300)
  assert r.returncode == 0 and r.stdout.strip().endswith('ok'), r.stderr[-2000:]


def test_vtrace_then_r2d2():
  _run('vtrace', 'r2d2')


def test_r2d2_then_vtrace():
  _run('r2d2', 'vtrace')
