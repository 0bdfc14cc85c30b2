#!/usr/bin/env python
"""Central-inference throughput against the inference batch size and the number of hosts per GPU
(bench.py's inference_path_bench / inference_lanes_bench at other sizes).  One JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]]
import torch
import bench
from seed_rl_b200.dmlab import networks

agent = networks.ImpalaDeep(bench.A, bench.OBS, seed=0, conv_mode='tc3p')
out = {}
keep = ('inference_batch_size', 'inferences_per_sec', 'us_per_batch_p50')
for N, envs, iters in ((1024, 2048, 40),):
  r = bench.inference_path_bench(agent, N=N, num_envs=envs, iters=iters, warmup=10)
  out['one_host_batch_%d' % N] = {k: r[k] for k in keep}
r = bench.inference_lanes_bench(agent, lanes=2, N=256, num_envs=1024, iters=80, warmup=15)
out['two_hosts_batch_256'] = r
r = bench.inference_lanes_bench(agent, lanes=3, N=64, num_envs=256, iters=150, warmup=30)
out['three_hosts_batch_64'] = r
print(json.dumps(out))
