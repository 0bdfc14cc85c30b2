set +e
O=gpurun_out/s2j; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_planes.py tests/test_gpu_fullsize.py tests/test_gpu_dmlab_shape.py -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -n "^E  \|^FAILED\|passed\|failed" $O/pytest.log | cut -c1-300 | tail -12
timeout 500 python bench.py --steps 20 --warmup 5 > $O/bench_tc3p.json 2> $O/bench.err; tail -3 $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s2j/bench_tc3p.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['kernel_time_ms_per_step'], d['kernel_launches_per_step'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_tc3p.csv python bench.py --steps 2 --warmup 3 --no-extras > $O/ncu_launch_bench.log 2>&1
python tools/ncu_summary.py launches $O/launches_tc3p.csv > $O/launches_summary.txt 2>&1; head -16 $O/launches_summary.txt
