#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS instructions in libseedrl_b200.so:
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (cp.async.bulk), SYNCS (mbarrier).  Usage: tools/sass_counts.py > profiles/rNN_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'seed_rl_b200', 'libseedrl_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)),
                       capture_output=True, text=True).stdout.splitlines()
KEYS = ['UTCHMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'UTCBAR']
counts, cur, i = collections.OrderedDict(), None, 0
for line in sass.splitlines():
  m = re.search(r'Function : (\S+)', line)
  if m:
    cur = names[i] if i < len(names) else m.group(1)
    i += 1
    counts[cur] = collections.Counter()
    continue
  if cur is None:
    continue
  for k in KEYS:
    if re.search(r'\b' + k + r'\b', line.split('/*')[1] if line.strip().startswith('/*') and line.count('/*') > 1 else line):
      counts[cur][k] += 1
print('# %s' % subprocess.run(['cuobjdump', '--version'], capture_output=True, text=True).stdout.strip().splitlines()[-1])
print('# kernels with at least one of %s (of %d kernels in the library)' % (KEYS, len(counts)))
print('%-8s %-6s %-8s %-8s %-7s %-6s %-6s  kernel' % tuple(KEYS))
tot = collections.Counter()
for name, c in counts.items():
  if sum(c[k] for k in KEYS[:5]) == 0:
    continue
  tot.update(c)
  j = name.rfind('>('); short = (name[:j + 1] if j > 0 else name.split('(')[0])[:150]
  print('%-8d %-6d %-8d %-8d %-7d %-6d %-6d  %s' % tuple([c[k] for k in KEYS] + [short]))
print('%-8d %-6d %-8d %-8d %-7d %-6d %-6d  TOTAL' % tuple(tot[k] for k in KEYS))
