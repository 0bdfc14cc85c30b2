#!/usr/bin/env python
"""DRAM traffic / duration per kernel from `ncu --set full` reports, written as the small JSON that
bench.py reads for `roofline.traffic`:  tools/ncu_traffic.py out.json frames=<N> a.ncu-rep [b.ncu-rep ...]"""
import csv, io, json, subprocess, sys
out, frames, reps = sys.argv[1], int(sys.argv[2].split('=')[1]), sys.argv[3:]
res = {}
for path in reps:
  raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  h, units = rows[0], rows[1]
  col = lambda name: [i for i, x in enumerate(h) if x == name][0]
  def to_bytes(v, u):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
  for r in rows[2:]:
    name = r[col('Kernel Name')].split('(')[0]
    ir, iw, it = col('dram__bytes_read.sum'), col('dram__bytes_write.sum'), col('gpu__time_duration.sum')
    tc = [i for i, x in enumerate(h) if x == 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active']
    res[name] = {'dram_read_bytes': to_bytes(r[ir], units[ir]), 'dram_write_bytes': to_bytes(r[iw], units[iw]),
                 'duration_us_under_ncu': float(r[it].replace(',', '')) * ({'us': 1, 'ms': 1e3, 'ns': 1e-3}.get(units[it], 1)),
                 'tensor_pipe_active_pct': float(r[tc[0]]) if tc else None, 'frames': frames, 'report': path.split('/')[-1]}
json.dump(res, open(out, 'w'), indent=1)
print(json.dumps(res, indent=1))
