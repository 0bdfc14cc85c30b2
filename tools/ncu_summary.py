"""Small text summaries of ncu artefacts (so profiles/ holds readable evidence, not 10 MB reports).

  python tools/ncu_summary.py launches <launches.csv>     per-kernel totals of a --metrics gpu__time_duration.sum list
  python tools/ncu_summary.py report <file.ncu-rep>       key metrics + stall hot spots of a --set full capture
"""
import collections, csv, io, re, subprocess, sys


def launches(path):
  rows = list(csv.reader(open(path)))
  hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
  h = rows[hi]; kn = h.index('Kernel Name'); mv = h.index('Metric Value')
  d = collections.defaultdict(list)
  for r in rows[hi + 2:]:
    if len(r) > mv:
      d[re.sub(r'\(.*', '', r[kn])].append(float(r[mv].replace(',', '')))
  tot = sum(sum(v) for v in d.values())
  print('# %s: %d launches, %.1f us total (cold-cache, serialised: shares are meaningful, absolutes are not)'
        % (path, sum(len(v) for v in d.values()), tot / 1e3))
  for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print('%-72s n=%4d total %9.1f us  avg %8.1f us  %5.1f%%' % (k[:72], len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))


KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_tc', 'sm__pipe_tc', 'sm__inst_executed_pipe_xu.sum.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max']


def report(path):
  raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  h, units = rows[0], rows[1]
  for kr in rows[2:]:
    name = kr[h.index('Kernel Name')]
    print('## %s' % name[:110])
    for k in KEYS:
      for i, x in enumerate(h):
        if x.startswith(k) and 'per_second' not in x:
          print('  %-75s %s %s' % (x, kr[i], units[i]))
  src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(src)))
  body = [r for r in rows[2:] if len(r) > 5 and r[0].startswith('0x')]
  if not body:
    return
  ts = sum(float(r[2] or 0) for r in body) or 1; ti = sum(float(r[5] or 0) for r in body) or 1
  print('## warp-sample hot spots (SASS, >= 1.5%% of samples; %d samples, %d warp-instructions)' % (ts, ti))
  for idx, r in enumerate(body):
    s = 100 * float(r[2] or 0) / ts
    if s >= 1.5:
      print('  %5d  %5.1f%% smp %5.2f%% inst  %s' % (idx, s, 100 * float(r[5] or 0) / ti, r[1].strip()[:90]))


if __name__ == '__main__':
  {'launches': launches, 'report': report}[sys.argv[1]](sys.argv[2])
