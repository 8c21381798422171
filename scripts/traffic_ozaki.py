"""profiles/traffic.json for the dominant kernel (ozaki_update_kernel) from an ncu capture of
`python scripts/profile_step.py 2` with --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
-k regex:ozaki_update: DRAM bytes of ALL its launches of ONE trial-step (second run of the script: V^T V of the
posterior covariance, the rank-1024 trailing updates and the rank-512 middle updates of chol(Sigma); batch 1, so the
launch-level look-ahead splits each trailing update in two), against the algorithmic bytes: every C tile read and
written once, the int8 slice images of the operand rows read once, the row scales.
Usage: python scripts/traffic_ozaki.py gpurun_out/r02_traffic_ozaki.csv profiles/r02_traffic_ozaki.csv"""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
h = rows[0]
ki, mi, vi, ii = (h.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID'))
launches = {}
for r in rows[1:]:
    launches.setdefault(int(r[ii]), {'kernel': r[ki]})[r[mi]] = float(r[vi].replace(',', ''))
upd = [l for _, l in sorted(launches.items()) if 'ozaki_update' in l['kernel']]
upd = upd[len(upd) // 2:]   # second run of the script
dram = sum(l['dram__bytes_read.sum'] + l['dram__bytes_write.sum'] for l in upd)
ns = sum(l['gpu__time_duration.sum'] for l in upd)

# algorithmic bytes of the same launches (m = 8192, n = 1024, outer 2048, middle 1024, in-panel updates on int8 too;
# 7 slices of 1 byte per element): every C tile read and written once (the sampling tail writes nothing), the int8
# slice images of the operand rows read once per launch, the row scales
M, N, S, OUT, MID, T = 8192, 1024, 256, 2048, 1024, 128
tile = 128 * 64 * 8.0


def tri_tiles(r):        # lower 128 x 64 tiles of a square region of r row tiles
    return r * (r + 1)


alg = tri_tiles(M // T) * tile + M * N * 7.0 + M * 8.0                # posterior covariance: C written, V slices
nt = M // T
for J0 in range(0, nt, OUT // T):
    J1 = min(J0 + OUT // T, nt)
    for M0 in range(J0, J1, MID // T):
        M1 = min(M0 + MID // T, J1)
        for j in range(M0 + 1, M1):                                    # left-looking in-panel updates of block column j
            if nt - j >= 8:
                rws = nt - j
                alg += rws * 2 * tile * 2 + rws * T * (j - M0) * T * 7.0 + rws * T * 8.0
        if M1 < J1 and nt - M1 >= 8:                                    # middle update inside the outer panel
            rws, cols = nt - M1, J1 - M1
            tiles = sum(min(2 * cols, 2 * i + 2) for i in range(rws))  # lower_skip rectangle
            alg += tiles * tile * 2 + rws * T * (M1 - M0) * T * 7.0 + rws * T * 8.0
    R = nt - J1
    if R >= 8:                                                         # trailing update
        alg += tri_tiles(R) * tile * 2 + R * T * (J1 - J0) * T * 7.0 + R * T * 8.0
alg += (M * (M + T) / 2.0) * 7.0 + S * M * 7.0 + (M + S) * 8.0         # sampling tail: L (lower blocks) and Z^T slices
out = {'kernel': 'ozaki_update_kernel (tcgen05.mma kind::i8), all three modes: posterior covariance, in-panel / middle / '
                 'trailing updates of chol(Sigma), fused sampling tail',
       'launches': len(upd), 'dram_bytes': dram, 'algorithmic_bytes': alg, 'ncu_ms': ns / 1e6,
       'dram_bytes_per_launch': dram / max(len(upd), 1), 'algorithmic_bytes_per_launch': alg / max(len(upd), 1),
       'scope': 'all %d launches of the kernel in ONE trial-step (m=8192, n=1024, batch 1: the look-ahead splits each '
                'trailing update in two); algorithmic = C tiles read + written once, int8 slice images of the operand '
                'rows read once, row scales' % len(upd),
       'source': '%s (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k '
                 'regex:ozaki_update; python scripts/profile_step.py 2, second run)' % (sys.argv[2] if len(sys.argv) > 2 else src)}
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    shutil.copy(src, os.path.join(ROOT, sys.argv[2]))
    with open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w') as fh:
        json.dump(out, fh, indent=1)
