"""Device kernel totals of one batched LML pass (64 candidates): python tests/tools/profile_lml.py [n] [d]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    import torch
    from torch.profiler import profile, ProfilerActivity
    from ocbo_b200 import synth
    from ocbo_b200.gp import hp_search
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 755
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    rs = np.random.RandomState(1)
    X = rs.uniform(size=(n, d))
    y = np.asarray([synth.hartmann6(x) for x in X]) if d == 6 else np.sin(3 * X.sum(1))
    thetas = hp_search.hp_candidates(X, y, 64, uniforms=rs.uniform(size=(63, 3 + d)))
    f = lambda: hp_search.batched_lml(X, y, 0, thetas)[0]
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    print('n=%d d=%d: %.3f ms per pass' % (n, d, (time.perf_counter() - t0) / reps * 1e3))
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(reps):
            f()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / reps, e.count / reps) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    print('--- device time per pass: %.1f us' % sum(r[1] for r in rows))
    for k, us, c in rows[:14]:
        print('  %8.1f us  x%5.1f  %s' % (us, c, k[:90]))


if __name__ == '__main__':
    main()
