set +e
O=gpurun_out/s2h; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_r2d2.py tests/test_gpu_dmlab_shape.py tests/test_gpu_zz_tc.py -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -n "^E  \|^FAILED\|passed\|failed\|R2D2_NET\|DMLAB_SHAPE" $O/pytest.log | cut -c1-300 | tail -30
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_tc3p.csv python bench.py --steps 2 --warmup 3 --no-extras > $O/ncu_launch_bench.log 2>&1
python tools/ncu_summary.py launches $O/launches_tc3p.csv | head -24
