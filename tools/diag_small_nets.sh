# Per-kernel launch lists (ncu, cold-cache: shares only) of the shallow-net and R2D2 learner steps, and the
# effect of the split-K wave count on both.  Usage (GPU box): bash tools/diag_small_nets.sh
set +e
O=gpurun_out/diag; mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_shallow.csv \
  python bench.py --net shallow --steps 2 --warmup 3 --no-extras > $O/shallow_ncu.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_r2d2.csv \
  python tools/r2d2_time.py > $O/r2d2_ncu.log 2>&1
for w in 2 4; do
  SEEDRL_GEMM_WAVES=$w timeout 200 python bench.py --net shallow --steps 30 --warmup 8 --no-extras > $O/shallow_w$w.json 2> $O/shallow_w$w.err
done
SEEDRL_GEMM_WAVES=4 timeout 300 python bench.py --agent r2d2 --steps 6 --warmup 8 --no-extras > $O/r2d2_w4.json 2> $O/r2d2_w4.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/diag/*_w*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], round(d['ms_per_step'], 3))
    except Exception as e:
        print(f, 'ERR', e)
PY
