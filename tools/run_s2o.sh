set +e
O=gpurun_out/s2o; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_inference.py -q -x > $O/pytest_inf.log 2>&1; grep -n "^E  \|^FAILED\|passed\|failed" $O/pytest_inf.log | cut -c1-300 | tail -10
timeout 500 python bench.py --steps 20 --warmup 5 > $O/bench_tc3p.json 2> $O/bench.err; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s2o/bench_tc3p.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernel_time_ms_per_step'])
print(json.dumps(d.get('inference_path'))[:900])
PY
