set +e
O=gpurun_out/s2c; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_r2d2.py -q -x > $O/pytest_r2d2.log 2>&1; echo "rc=$?" >> $O/pytest_r2d2.log
tail -30 $O/pytest_r2d2.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_r2d2.py > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
tail -6 $O/pytest_all.log
timeout 300 python tools/r2d2_time.py > $O/r2d2_time.json 2> $O/r2d2_time.err; tail -3 $O/r2d2_time.err; cat $O/r2d2_time.json
MODE=simt timeout 300 python tools/r2d2_time.py > $O/r2d2_time_simt.json 2>> $O/r2d2_time.err; cat $O/r2d2_time_simt.json
timeout 300 python bench.py --net shallow --conv tc3 --steps 20 --warmup 5 > $O/bench_cfg2_shallow.json 2> $O/bench_cfg2.err; head -c 700 $O/bench_cfg2_shallow.json
