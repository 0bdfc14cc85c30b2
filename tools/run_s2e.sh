set +e
O=gpurun_out/s2e; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_r2d2.py tests/test_gpu_inference.py tests/test_gpu_checkpoint.py tests/test_gpu_planes.py tests/test_gpu_parity.py -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -n "^E  \|^FAILED\|passed\|failed" $O/pytest.log | cut -c1-300 | tail -30
timeout 500 python bench.py --steps 20 --warmup 5 > $O/bench_tc3p.json 2> $O/bench.err; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s2e/bench_tc3p.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernel_time_ms_per_step'])
print(json.dumps(d.get('inference_path'))[:1500])
PY
