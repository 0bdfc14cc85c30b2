# Round-end measurement batch (GPU box): full GPU test suite, every bench config, the reference arm, smoke,
# launch lists of the shallow / R2D2 steps and one ncu --set full capture of the gathered-operand GEMM.
# Usage: bash tools/run_final.sh [quick]    (quick: skips the reference arm and the ncu passes)
set +e
O=gpurun_out/final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 10 > $O/bench_cfg4_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
timeout 300 python bench.py --batch 256 --steps 20 --warmup 5 --no-extras > $O/bench_cfg3_b256.json 2> $O/bench_cfg3.err
timeout 300 python bench.py --net shallow --steps 50 --warmup 10 > $O/bench_cfg2_shallow.json 2> $O/bench_cfg2.err
timeout 600 python bench.py --agent r2d2 --steps 10 --warmup 16 > $O/bench_cfg5_r2d2.json 2> $O/bench_cfg5.err; tail -2 $O/bench_cfg5.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
if [ "$1" != "quick" ]; then
  timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference_arm.json 2> $O/bench_ref.err
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_shallow.csv \
    python bench.py --net shallow --steps 2 --warmup 3 --no-extras > $O/shallow_ncu.log 2>&1
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_r2d2.csv \
    python tools/r2d2_time.py > $O/r2d2_ncu.log 2>&1
  # the four gathered-operand GEMMs of one shallow-net step: conv0 / conv1 forward, conv1 / conv0 weight gradient
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:gemm_tc_kernel<\(bool\)[01], \(bool\)0, \(bool\)1, \(int\)32, \(bool\)1>' -c 4 -o $O/gemm_gather -f \
    python bench.py --net shallow --steps 2 --warmup 3 --no-extras > $O/gemm_gather_ncu.log 2>&1
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/final/bench_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], round(d['ms_per_step'], 3), round(d['value']), round(d['e2e']['value']))
    except Exception as e:
        print(f, 'ERR', e)
PY
