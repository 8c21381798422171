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

from tests.tools.bench_small import _hp, _timeit  # noqa: E402


def measure(Rs=(1, 2, 4, 8, 16, 32, 64), reps=10):
    import gc
    import torch
    from ocbo_b200 import batched
    from ocbo_b200.gp.gp_interface import make_gp
    rs = np.random.RandomState(0)
    np.random.seed(0)
    out = {}
    for name, T, d, nf in (('set2d', 5, 2, 50), ('rand6d', 30, 6, 55)):
        res = {}
        for R in Rs:
            if R * T > 1024:
                continue
            Xl = [rs.uniform(size=(nf, d)) for _ in range(R * T)]
            yl = [np.sin(3 * X.sum(1)) + 0.01 * rs.normal(size=nf) for X in Xl]
            gb = []
            for X, y in zip(Xl, yl):
                h = _hp(X, y)
                gb.append(make_gp(X, y, 'se', h['scale'], h['bw'], h['noise'], h['mu0'], build=False))
            pts = [np.vstack([X, np.random.uniform(size=(100, d))]) for X in Xl]
            Z = [np.random.normal(size=(len(p), 1)) for p in pts]
            gc.collect()
            gc.freeze()
            t = _timeit(lambda: batched.mts_step(gb, pts, [nf] * (R * T), Z), reps, torch.cuda.synchronize)
            res[R] = {'ms_per_pass': 1e3 * t, 'steps_per_s': R / t}
        out[name] = res
    return out


if __name__ == '__main__':
    print(json.dumps(measure(), indent=1))
