set +e
O=gpurun_out/s2f; mkdir -p $O
timeout 300 python tools/r2d2_debug.py > $O/r2d2_debug.log 2>&1; cat $O/r2d2_debug.log | tail -70
timeout 600 python bench.py --agent r2d2 --steps 5 --warmup 3 > $O/bench_cfg5_r2d2.json 2> $O/bench_r2d2.err; tail -5 $O/bench_r2d2.err; head -c 2500 $O/bench_cfg5_r2d2.json
