set +e
O=gpurun_out/s2b; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -8 $O/pytest.log
timeout 300 python bench.py --batch 256 --steps 10 --warmup 3 --no-extras > $O/bench_cfg3_b256.json 2> $O/bench_cfg3.err
timeout 300 python bench.py --net shallow --steps 20 --warmup 5 > $O/bench_cfg2_shallow.json 2> $O/bench_cfg2.err
timeout 200 python tools/planes_bench.py > $O/planes_bench.json 2>&1
# launch list of the default bench command (two steps)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_tc3p.csv python bench.py --steps 2 --warmup 3 --no-extras > $O/ncu_launch_bench.log 2>&1
# full capture of the dominant conv instance at the cfg-3 frame count (5376 frames)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:convp_kernel -s 2 -c 1 -o $O/convp_16_16_b256 python tools/planes_one.py conv 16 16 42 5376 > $O/ncu_convp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgradp_kernel -s 2 -c 1 -o $O/wgradp_16_16_b256 python tools/planes_one.py wgrad 16 16 42 5376 > $O/ncu_wgradp.log 2>&1
ls -la $O
head -c 1500 $O/bench_cfg3_b256.json; echo; head -c 600 $O/bench_cfg2_shallow.json
