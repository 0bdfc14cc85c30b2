set +e
O=gpurun_out/s2n; mkdir -p $O
for ov in 1 0 1 0; do
SEEDRL_OVERLAP_REDUCE=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 10 --no-extras > $O/bench_2gpu_ov$ov.json 2> $O/bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/s2n/bench_2gpu_ov$ov.json')); print('overlap=$ov', d['ms_per_step'], d['ms_per_step_median'], d['replicas_in_sync'])"
done
timeout 200 python bench.py --steps 40 --warmup 10 --no-extras > $O/bench_1gpu.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/s2n/bench_1gpu.json')); print('1gpu', d['ms_per_step'], d['ms_per_step_median'])"
timeout 600 python -m pytest tests/test_gpu_inference.py tests/test_gpu_r2d2.py tests/test_gpu_parity.py -q -x > $O/pytest.log 2>&1; grep -n "^E  \|^FAILED\|passed\|failed" $O/pytest.log | cut -c1-250 | tail -8
