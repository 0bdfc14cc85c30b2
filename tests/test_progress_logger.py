"""ProgressLogger / SummaryWriter (reference common/utils.py:546-677) and the replica helpers of
learner_loop -- host logic, CPU only."""
import json
import os
import time

import torch

from seed_rl_b200.common import utils


def test_progress_logger_exports_scalars_and_speed(tmp_path):
  w = utils.SummaryWriter(str(tmp_path))
  logger = utils.ProgressLogger(summary_writer=w, initial_period=0.01, period_factor=1.0, starting_step=100)
  calls = []
  logger.start(lambda: calls.append(1))
  try:
    for it in range(3):
      session = logger.log_session()
      # the reference's scalar names (agents/vtrace/learner.py:138-157)
      logger.log(session, 'losses/total', torch.tensor(1.5 + it))
      logger.log(session, 'V/value function', torch.tensor(0.25))
      logger.step_end(session, None, step_increment=1280)
      time.sleep(0.06)
  finally:
    logger.shutdown()
  w.close()
  rows = [json.loads(l) for l in open(os.path.join(str(tmp_path), 'summaries.jsonl'))]
  tags = {r['tag'] for r in rows}
  assert {'losses/total', 'V/value function', 'speed/steps_per_sec'} <= tags
  assert logger.log_keys == ['losses/total', 'V/value function']
  assert calls, 'logging callback never ran'
  steps = sorted({r['step'] for r in rows})
  assert steps[0] > 100 and (steps[-1] - 100) % 1280 == 0 and steps[-1] == 100 + 3 * 1280
  last = [r for r in rows if r['tag'] == 'losses/total'][-1]
  assert last['value'] == 3.5
  speed = [r['value'] for r in rows if r['tag'] == 'speed/steps_per_sec']
  assert all(v > 0 for v in speed)


def test_progress_logger_key_value_mismatch_is_reported(tmp_path, caplog):
  logger = utils.ProgressLogger(summary_writer=None, initial_period=0.01, period_factor=1.0)
  s = logger.log_session_from_dict({'a': 1.0, 'b': 2.0})
  assert s == [1.0, 2.0] and logger.log_keys == ['a', 'b']
  logger.step_end([1.0], None, 1)            # wrong number of values: the logger thread must not die silently
  try:
    logger._log()
    raised = False
  except AssertionError as e:
    raised = 'Mismatch between number of keys and values' in str(e)
  assert raised


def test_rank_server_address_and_backoff():
  from seed_rl_b200.agents.vtrace import learner_loop
  assert learner_loop.rank_server_address('localhost:8686', 0) == 'localhost:8686'
  assert learner_loop.rank_server_address('localhost:8686', 3) == 'localhost:8689'
  assert learner_loop.rank_server_address('unix:/tmp/foo', 2) == 'unix:/tmp/foo.2'
  logger = utils.ProgressLogger(initial_period=0.5, period_factor=2.0, max_period=3.0)
  logger.start()
  time.sleep(0.05)
  logger.shutdown()
  assert 0.5 < logger.period <= 3.0          # exponential back-off, capped (utils.py:672-676)
