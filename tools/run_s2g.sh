set +e
O=gpurun_out/s2g; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -n "^E  \|^FAILED\|passed\|failed\|R2D2_NET" $O/pytest.log | cut -c1-300 | tail -30
timeout 300 python bench.py --net shallow --steps 20 --warmup 5 --no-extras > $O/bench_cfg2_shallow.json 2> $O/bench_cfg2.err; tail -2 $O/bench_cfg2.err; head -c 400 $O/bench_cfg2_shallow.json; echo
timeout 600 python bench.py --agent r2d2 --steps 5 --warmup 8 > $O/bench_cfg5_r2d2.json 2> $O/bench_r2d2.err; tail -3 $O/bench_r2d2.err; head -c 500 $O/bench_cfg5_r2d2.json; echo
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > $O/bench_tc3p.json 2> $O/bench.err; head -c 400 $O/bench_tc3p.json
