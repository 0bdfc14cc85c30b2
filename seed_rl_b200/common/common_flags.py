"""Flags common across learners and actors -- same names/defaults as the reference's
`common/common_flags.py:21-43`."""
from absl import flags

flags.DEFINE_string('logdir', '/tmp/agent', 'Log directory.')
flags.DEFINE_string('server_address', 'localhost:8686', 'Server address.')
flags.DEFINE_enum('run_mode', None, ['learner', 'actor'],
                  'Whether we run the learner or the actor.')
flags.DEFINE_integer('num_eval_envs', 0, 'Number of environments that will be used for eval.')
flags.DEFINE_integer('env_batch_size', 1,
                     'How many environments to operate on together in a batch.')
flags.DEFINE_integer('num_envs', 4, 'Total number of environments in all actors.')
flags.DEFINE_integer('num_action_repeats', 1, 'Number of action repeats.')


def define_once(define_fn, name, *args, **kwargs):
  """The reference's learners are separate binaries that re-use flag names (batch_size,
  unroll_length, discounting, save_checkpoint_secs ...).  Here both mirrors can live in one
  process (the test-suite imports both): the first definition of a name stands."""
  if name not in flags.FLAGS:
    define_fn(name, *args, **kwargs)
