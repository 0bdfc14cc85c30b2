set +e
O=gpurun_out/s2r; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_planes.py tests/test_gpu_dmlab_shape.py tests/test_gpu_fullsize.py -q > $O/pytest.log 2>&1; grep -n "^E  \|^FAILED\|passed\|failed" $O/pytest.log | cut -c1-300 | tail -10
timeout 300 python bench.py --steps 30 --warmup 8 --no-extras > $O/bench_tc3p.json 2> $O/bench.err; tail -2 $O/bench.err
python -c "
import json; d=json.load(open('gpurun_out/s2r/bench_tc3p.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
