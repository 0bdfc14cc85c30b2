"""A few eager (no CUDA graph) central-inference batches, for an ncu launch list:
ncu --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/inference_launches.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from seed_rl_b200.dmlab import networks
agent = networks.ImpalaDeep(18, (84, 84, 4), seed=0, conv_mode='tc3p')
r = bench.inference_path_bench(agent, cuda_graph=False, iters=4, warmup=4)
print(r['us_per_batch_p50'])
