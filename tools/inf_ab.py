import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from seed_rl_b200.dmlab import networks
agent = networks.ImpalaDeep(18, (84, 84, 4), seed=0, conv_mode='tc3p')
for g in (True, False, True, False):
  r = bench.inference_path_bench(agent, cuda_graph=g, iters=300)
  print(json.dumps({k: r[k] for k in ('cuda_graph', 'inferences_per_sec', 'us_per_batch_mean', 'us_per_batch_p50', 'us_per_batch_p99', 'library_launches_per_batch')}))
