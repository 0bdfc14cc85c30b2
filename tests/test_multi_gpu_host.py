"""CPU, world_size 2, gloo: the N>1 host logic of the learner -- the single gradient
exchange and its SUM semantics (reference tests/utils_test.py:609-650 `MinimizeTest`, which
needs a real TPU there), env sharding, and bench.py's reference arm under torchrun-style
env vars (rank 0 prints, other ranks exit quietly)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world))
  sys.path.insert(0, ROOT)
  from seed_rl_b200.agents.vtrace import learner
  dist.init_process_group('gloo', rank=rank, world_size=world)
  # every replica computes the gradient of ITS mean loss: g_r = 0.1 * (rank + 1)
  g = torch.full((1000,), 0.1 * (rank + 1))
  scale = learner.reduce_gradients(g, world, None, 'sum')
  res = {'sum': (float(g[0]), float(g[-1]), scale)}
  g = torch.full((1000,), 0.1 * (rank + 1))
  scale = learner.reduce_gradients(g, world, None, 'mean')
  res['mean'] = (float(g[0]) * scale, scale)
  # SGD step a -= lr * g with the reduced gradient, like MinimizeTest: a starts at 1, lr = 1
  a = 1.0 - res['sum'][0] * res['sum'][2]
  res['a'] = a
  res['shard'] = learner.env_shard(rank, world, 7)
  dist.barrier()
  dist.destroy_process_group()
  json.dump(res, open(os.path.join(out, 'r%d.json' % rank), 'w'))


def test_gradient_exchange_is_sum_across_replicas(tmp_path):
  world, port = 2, 29573
  mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
  r = [json.load(open(tmp_path / ('r%d.json' % i))) for i in range(world)]
  for x in r:
    assert abs(x['sum'][0] - 0.3) < 1e-6 and abs(x['sum'][1] - 0.3) < 1e-6 and x['sum'][2] == 1.0
    assert abs(x['mean'][0] - 0.15) < 1e-6 and x['mean'][1] == 0.5
    assert abs(x['a'] - 0.7) < 1e-6           # 1 - (g_0 + g_1): SUM, not mean
  assert r[0]['shard'] == [0, 2, 4, 6] and r[1]['shard'] == [1, 3, 5]
  assert r[0]['a'] == r[1]['a']               # replicas stay identical


def test_reference_arm_only_rank0_prints():
  env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
  p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                      '--gpus', '2', '--steps', '1', '--warmup', '0'], env=env,
                     capture_output=True, text=True, timeout=120)
  assert p.returncode == 0 and p.stdout.strip() == ''
