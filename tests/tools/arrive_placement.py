"""Controlled A/B of WHERE the "stage drained" mbarrier arrival of the free-running TMA main loop sits
(gemm_core.cuh, tma_mainloop): the shipped placement (first instruction of the next k-tile) against the
straight-line placement behind the k-tile's DMMAs (-DOCBO_ARRIVE_AT_END), everything else equal.
Builds the variant library next to the shipped one (nvcc, ~1 min), runs tests/tools/tma_modes.py on both
and prints where ptxas put the arrival relative to the last DMMA of the loop body in each.
    python tests/tools/arrive_placement.py > gpurun_out/r02_arrive_placement.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def arrival_positions(lib):
    """(#DMMA issued before the drained arrival inside the free-run loop body) for gemm_nt<128,64,TMA>."""
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    body, on = [], False
    for line in sass.splitlines():
        if 'Function :' in line:
            on = 'gemm_nt_kernelILi128ELi64ELi4ELi2ELb1' in line
        if on and re.search(r'/\*[0-9a-f]{4}\*/', line):
            body.append(line)
    out, dmma = [], 0
    for line in body:
        if 'DMMA' in line:
            dmma += 1
        if 'SYNCS.ARRIVE' in line and 'A1T0' in line:
            out.append(dmma)
    return out, dmma


def main():
    from ocbo_b200 import build
    var_dir = os.path.join(ROOT, 'gpurun_out', 'lib_arrive_at_end')
    var = build.build(defines=['OCBO_ARRIVE_AT_END=1'], out_dir=var_dir)
    for name, lib in (('shipped (arrival = first instruction of the next k-tile)', build.LIB_PATH),
                      ('variant (arrival straight-line behind the DMMAs)', var)):
        pos, total = arrival_positions(lib)
        print('== %s' % name)
        print('   gemm_nt_kernel<128,64,TMA>: %d DMMA in the kernel text (two 64-DMMA loop bodies: free-run, barrier); '
              '"drained" arrivals sit after DMMA #%s' % (total, pos))
        sys.stdout.flush()
        subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'tools', 'tma_modes.py'), '1'],
                       env=dict(os.environ, OCBO_B200_LIB=lib), check=False)
        sys.stdout.flush()


if __name__ == '__main__':
    main()
