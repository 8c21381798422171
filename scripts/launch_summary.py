"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).
Usage: python scripts/launch_summary.py gpurun_out/bench_launches.csv "<command that was profiled>" """
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki]).replace('(anonymous namespace)', '<unnamed>')
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(r[ui], 1e-6)
    tot[name] += float(r[vi].replace(',', '')) * scale
    cnt[name] += 1
total = sum(tot.values())
print('# ncu --metrics gpu__time_duration.sum --clock-control none --csv  %s' % (sys.argv[2] if len(sys.argv) > 2 else ''))
print('# cold-cache, serialised: compare SHARES.  total %.3f ms over %d launches' % (total, sum(cnt.values())))
for name in sorted(tot, key=lambda n: -tot[n]):
    print('%-72s launches=%5d  sum_ms=%9.3f  share=%5.1f%%' % (name[:72], cnt[name], tot[name], 100 * tot[name] / total))
