"""Aggregate MTS decision steps/s with R independent BO trials in flight: the R x T
independent-GP problems of R trials go through ONE batched device pass
(batched.mts_step is agnostic of which trial a task belongs to).  BASELINE.json's sweep
shards "independent BO trials"; this is the same axis at the option-file shapes, where one
trial alone leaves the GPU idle (set2d: 5 problems of <= 205 rows)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402



SHAPES = {'set2d': (5, 2, 50), 'rand6d': (30, 6, 55)}


def make_step(name, R, rs):
    from tests.tools.bench_small import _hp
    from ocbo_b200 import batched
    from ocbo_b200.gp.gp_interface import make_gp
    T, d, nf = SHAPES[name]
    Xl = [rs.uniform(size=(nf, d)) for _ in range(R * T)]
    yl = [np.sin(3 * X.sum(1)) + 0.01 * rs.normal(size=nf) for X in Xl]
    gb = []
    for X, y in zip(Xl, yl):
        h = _hp(X, y)
        gb.append(make_gp(X, y, 'se', h['scale'], h['bw'], h['noise'], h['mu0'], build=False))
    pts = [np.vstack([X, np.random.uniform(size=(100, d))]) for X in Xl]
    Z = [np.random.normal(size=(len(p), 1)) for p in pts]
    return lambda: batched.mts_step(gb, pts, [nf] * (R * T), Z)


def measure(Rs=(1, 2, 4, 8, 16, 32, 64), reps=10):
    import gc
    import torch
    from tests.tools.bench_small import _timeit
    rs = np.random.RandomState(0)
    np.random.seed(0)
    out = {}
    for name, (T, d, nf) in SHAPES.items():
        res = {}
        for R in Rs:
            if R * T > 1024:
                continue
            step = make_step(name, R, rs)
            gc.collect()
            gc.freeze()
            t = _timeit(step, reps, torch.cuda.synchronize)
            res[R] = {'ms_per_pass': 1e3 * t, 'steps_per_s': R / t}
        out[name] = res
    return out


def profile(name, R, reps=10):
    """Host (cProfile) and device (CUPTI) time of one pass with R trials in flight."""
    import cProfile
    import io
    import pstats
    import torch
    from torch.profiler import profile as tprofile, ProfilerActivity
    step = make_step(name, R, np.random.RandomState(0))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(24)
    print('=== %s R=%d' % (name, R))
    print('\n'.join(l for l in s.getvalue().splitlines()[4:] if l.strip())[:7000])
    with tprofile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(reps):
            step()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / reps, e.count / reps) for e in prof.key_averages()
            if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    print('--- device time per pass: %.1f us' % sum(r[1] for r in rows))
    for k, us, c in rows[:14]:
        print('  %8.1f us  x%5.1f  %s' % (us, c, k[:90]))


if __name__ == '__main__':
    if len(sys.argv) > 2:
        profile(sys.argv[1], int(sys.argv[2]))
    else:
        print(json.dumps(measure()))
