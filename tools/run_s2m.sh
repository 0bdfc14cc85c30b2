set +e
O=gpurun_out/s2m; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > $O/bench_2gpu.json 2> $O/bench_2gpu.err
tail -3 $O/bench_2gpu.err; python -c "
import json; d=json.load(open('gpurun_out/s2m/bench_2gpu.json')); print({k:d.get(k) for k in ('n_gpus','ms_per_step','value','replicas_in_sync','grad_exchange')}, d['e2e']['value'])"
