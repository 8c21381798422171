"""profiles/traffic.json from an ncu capture of scripts/profile_step.py 2 (scripts/gpu_round.sh:
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum): DRAM bytes of the
rank-2048 trailing-update launches (gemm_nt_kernel<128,64>, grid >= 272 lower-triangular tiles) of
ONE trial-step, second run of the script (the first is warm-up).
Usage: python scripts/traffic_summary.py gpurun_out/traffic.csv profiles/r02_traffic.csv"""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
h = rows[0]
ki, gi, mi, vi, ii = (h.index(k) for k in ('Kernel Name', 'Grid Size', 'Metric Name', 'Metric Value', 'ID'))
launches = {}
for r in rows[1:]:
    launches.setdefault(int(r[ii]), {'kernel': r[ki], 'grid': r[gi]})[r[mi]] = float(r[vi].replace(',', ''))
# outer updates of an m = 8192 factorisation: lower-triangular 2:1 tile sets of 48, 32, 16 row tiles
# (grids 2352 / 1056 / 272), or with look-ahead (one problem per launch) their (b) + (a) splits:
# 40*41 + 48*16, 24*25 + 32*16, 8*9 + 16*16
grids = {2352, 1056, 272, 1640, 768, 600, 512, 72, 256}
grid_of = lambda l: int(l['grid'].strip('()').split(',')[0])
outer = [l for _, l in sorted(launches.items()) if 'gemm_nt_kernel<128' in l['kernel'] and grid_of(l) in grids
         and l['gpu__time_duration.sum'] > 30e3]   # (the short-k in-panel launches of the same grids take < 30 us)
outer = outer[len(outer) // 2:]  # second run of the script
dram = sum(l['dram__bytes_read.sum'] + l['dram__bytes_write.sum'] for l in outer)
ns = sum(l['gpu__time_duration.sum'] for l in outer)
alg = 0.0
for rows_ in (6144, 4096, 2048):
    alg += (rows_ * (rows_ + 128) / 2.0) * 8 * 2 + rows_ * 2048 * 8   # C lower tiles read + written, panel read once
out = {'kernel': 'gemm_nt_kernel<128,64,4,2,TMA> rank-2048 trailing updates of chol(Sigma)',
       'launches': len(outer), 'dram_bytes': dram, 'algorithmic_bytes': alg, 'ncu_ms': ns / 1e6,
       'scope': 'the %d outer-update launches of ONE trial-step (m=8192, batch 1, OCBO_POTRF_OUTER=2048, banded tile '
                'order; with one problem per launch the look-ahead splits each update in two); algorithmic = C lower tiles read + written once, L panel read once'
                % len(outer),
       'source': '%s (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum; python '
                 'scripts/profile_step.py 2, second run)' % (sys.argv[2] if len(sys.argv) > 2 else src)}
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    shutil.copy(src, os.path.join(ROOT, sys.argv[2]))
    with open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w') as fh:
        json.dump(out, fh, indent=1)
