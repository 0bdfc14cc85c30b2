set +e
O=gpurun_out/s2q; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:convp_kernel -s 2 -c 1 -o $O/convp_16_16 python tools/planes_one.py conv 16 16 42 1344 > $O/ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv0pool_kernel -s 3 -c 1 -o $O/conv0pool python bench.py --steps 1 --warmup 3 --no-extras > $O/ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:first_wgrad_pooled_kernel -s 3 -c 1 -o $O/first_wgrad python bench.py --steps 1 --warmup 3 --no-extras > $O/ncu3.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lstm2_bwd_kernel -s 3 -c 1 -o $O/lstm2_bwd python bench.py --steps 1 --warmup 3 --no-extras > $O/ncu4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_tc3p.csv python bench.py --steps 2 --warmup 3 --no-extras > $O/ncu_launch_bench.log 2>&1
python tools/ncu_summary.py launches $O/launches_tc3p.csv > $O/launches_summary.txt 2>&1; head -22 $O/launches_summary.txt
ls -la $O
