set +e
O=gpurun_out/s2i; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_planes.py -q -k "conv0pool or first_layer" > $O/pytest_c0.log 2>&1; echo "rc=$?" >> $O/pytest_c0.log
grep -n "^E  \|^FAILED\|passed\|failed" $O/pytest_c0.log | cut -c1-300 | tail -20
timeout 900 python -m pytest tests/test_gpu_planes.py tests/test_gpu_fullsize.py tests/test_gpu_dmlab_shape.py tests/test_gpu_checkpoint.py -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -n "^E  \|^FAILED\|passed\|failed" $O/pytest.log | cut -c1-300 | tail -20
timeout 500 python bench.py --steps 20 --warmup 5 > $O/bench_tc3p.json 2> $O/bench.err; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s2i/bench_tc3p.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernel_time_ms_per_step'], d['kernel_launches_per_step'])
PY
