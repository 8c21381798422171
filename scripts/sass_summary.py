"""Per-kernel SASS evidence of libocbo_b200.so: DMMA / UTMALDG (TMA) / SYNCS (mbarrier) / LDGSTS
(cp.async) instruction counts, registers and stack from cuobjdump.  Runs on the build container
(no GPU):  python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'ocbo_b200', 'lib', 'libocbo_b200.so')


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout
    return out.splitlines()


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    res = subprocess.run(['cuobjdump', '--dump-resource-usage', LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r'Function (\S+):', line)
        if m:
            cur = m.group(1)
            continue
        if cur and 'REG:' in line:
            reg = re.search(r'REG:(\d+)', line)
            stack = re.search(r'STACK:(\d+)', line)
            shared = re.search(r'SHARED:(\d+)', line)
            usage[cur] = (int(reg.group(1)), int(stack.group(1)), int(shared.group(1)))
            cur = None
    counts, name = {}, None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = m.group(1)
            counts[name] = {'DMMA': 0, 'UTCIMMA': 0, 'LDTM': 0, 'UBLKCP': 0, 'UTMALDG': 0, 'SYNCS': 0, 'LDGSTS': 0, 'LDS': 0, 'BAR': 0}
            continue
        if name:
            for key in counts[name]:
                if re.search(r'\b' + key + r'[\.\s]', line):
                    counts[name][key] += 1
    names = sorted(counts)
    pretty = demangle(names)
    print('# cuobjdump -sass / --dump-resource-usage of ocbo_b200/lib/libocbo_b200.so (sm_100a)')
    print('# UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld (TMEM -> registers), UBLKCP = cp.async.bulk (contiguous),')
    print('# UTMALDG = cp.async.bulk.tensor (TMA), SYNCS = mbarrier ops, LDGSTS = cp.async, DMMA = fp64 tensor MMA')
    print('%-80s %5s %5s %6s | %5s %7s %4s %6s %7s %5s %6s %5s %4s' % ('kernel', 'regs', 'stack', 'smem', 'DMMA', 'UTCIMMA',
                                                                     'LDTM', 'UBLKCP', 'UTMALDG', 'SYNCS', 'LDGSTS', 'LDS', 'BAR'))
    tot = {}
    for n, p in zip(names, pretty):
        c = counts[n]
        r = usage.get(n, (0, 0, 0))
        short = re.sub(r'\(.*', '', p.replace('(anonymous namespace)::', '')).replace('ocbo::', '')[:86]
        print('%-80s %5d %5d %6d | %5d %7d %4d %6d %7d %5d %6d %5d %4d' % (short[:80], r[0], r[1], r[2], c['DMMA'], c['UTCIMMA'],
                                                                        c['LDTM'], c['UBLKCP'], c['UTMALDG'], c['SYNCS'],
                                                                        c['LDGSTS'], c['LDS'], c['BAR']))
        for k, v in c.items():
            tot[k] = tot.get(k, 0) + v
    print('# totals: ' + ', '.join('%s %d' % kv for kv in sorted(tot.items())))


if __name__ == '__main__':
    main()
