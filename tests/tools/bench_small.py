"""MTS BO decision steps/s at the four option-file shapes (fixed hyper-parameters):
ocbo_b200's batched device pass vs the numpy oracle loop (the reference's structure:
per-task / per-context Python loop of LAPACK calls).  Used by bench.py's CPU-baseline
leg ("mts_bo_steps") and runnable on its own.  Shapes: BASELINE.md section 1."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def _timeit(f, reps, sync):
    f()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    sync()
    return (time.perf_counter() - t0) / reps


def _hp(X, y):
    v = max(float(np.var(y)), 1e-6)
    return dict(scale=v, bw=0.3 * (X.max(0) - X.min(0) + 1e-3), noise=0.01 * v, mu0=float(np.median(y)))


def measure(gpu_reps=20, cpu_reps=3):
    import torch
    from ocbo_b200 import batched, synth
    from ocbo_b200.gp.gp_interface import make_gp
    from oracle import gp_numpy as o, picks
    sync = torch.cuda.synchronize
    nosync = lambda: None
    out = {}
    rs = np.random.RandomState(0)
    np.random.seed(0)
    # set2d (T=5, d=2, n_f=50 mid-run) and rand6d (T=30, d=6, n_f=55): independent GPs (mts.py)
    for name, T, d, nf in (('set2d', 5, 2, 50), ('rand6d', 30, 6, 55)):
        Xl = [rs.uniform(size=(nf, d)) for _ in range(T)]
        yl = [np.sin(3 * X.sum(1)) + 0.01 * rs.normal(size=nf) for X in Xl]
        hl = [_hp(X, y) for X, y in zip(Xl, yl)]
        gb = [make_gp(X, y, 'se', h['scale'], h['bw'], h['noise'], h['mu0'], build=False)
              for X, y, h in zip(Xl, yl, hl)]
        go = [o.make_gp(X, y, 'se', h['scale'], h['bw'], h['noise'], h['mu0'])
              for X, y, h in zip(Xl, yl, hl)]

        def gpu_step():
            pts = [np.vstack([X, np.random.uniform(size=(100, d))]) for X in Xl]
            Z = [np.random.normal(size=(len(p), 1)) for p in pts]
            return batched.mts_step(gb, pts, [nf] * T, Z)

        def cpu_step():
            res = []
            for g, X in zip(go, Xl):
                g.gp_core.build_posterior()  # get_next_gp always rebuilds (gp_interface.py:132)
                res.append(g.draw_sample(samp_pts=np.vstack([X, np.random.uniform(size=(100, d))])))
            return picks.mts_pick(res, [nf] * T)
        out[name] = {'b200_steps_per_s': 1 / _timeit(gpu_step, gpu_reps, sync),
                     'cpu_steps_per_s': 1 / _timeit(cpu_step, cpu_reps, nosync),
                     'shape': 'T=%d, d=%d, n_f=%d, A=100' % (T, d, nf)}
    # jointbran (T=10, n=160, m=1160): joint GP (joint_mts.py)
    T, A, n = 10, 100, 160
    X = np.hstack([np.repeat(np.linspace(0, 1, T), n // T)[:, None], rs.uniform(size=(n, 1))])
    y = np.asarray([synth.branin(x) for x in X])
    h = _hp(X, y)
    gbj = make_gp(X, y, 'se', h['scale'], h['bw'], h['noise'], h['mu0'], build=False)
    goj = o.make_gp(X, y, 'se', h['scale'], h['bw'], h['noise'], h['mu0'])
    seg = np.concatenate([np.arange(0, n + 1, n // T), n + A * np.arange(1, T + 1)]).astype(np.int32)
    mkgrid = lambda: np.hstack([np.repeat(np.linspace(0, 1, T), A)[:, None],
                                np.random.uniform(size=(T * A, 1))])

    def gpu_step():
        gbj.gp_core._post = None
        return gbj.gp_core.sample_and_pick(np.vstack([X, mkgrid()]), seg, T)

    def cpu_step():
        goj.gp_core.build_posterior()
        s = goj.draw_sample(samp_pts=np.vstack([X, mkgrid()]))
        return picks.joint_mts_pick(s, [n // T] * T, T)
    out['jointbran'] = {'b200_steps_per_s': 1 / _timeit(gpu_step, gpu_reps, sync),
                        'cpu_steps_per_s': 1 / _timeit(cpu_step, cpu_reps, nosync),
                        'shape': 'T=10, n=160, m=1160'}
    # conth42 (n=755, 50 contexts x 100 actions): profile_cts.py
    n, P = 755, 50
    X = rs.uniform(size=(n, 6))
    y = np.asarray([synth.hartmann6(x) for x in X])
    h = _hp(X, y)
    gbc = make_gp(X, y, 'se', h['scale'], h['bw'], h['noise'], h['mu0'], build=False)
    goc = o.make_gp(X, y, 'se', h['scale'], h['bw'], h['noise'], h['mu0'])

    def sets():
        return [np.hstack([np.tile(c, (100, 1)), np.random.uniform(size=(100, 2))])
                for c in np.random.uniform(size=(P, 4))]

    def gpu_step():
        acts = sets()
        gbc.gp_core._post = None
        return batched.profile_step(gbc, acts, [np.random.normal(size=(100, 1)) for _ in acts])

    def cpu_step():
        goc.gp_core.build_posterior()
        res = []
        for a in sets():
            mu, cov = goc.eval(a, include_covar=True)
            res.append(picks.cmts_gain(goc.draw_sample(means=mu, covar=cov), mu)[1])
        return picks.profile_pick(res)
    out['conth42'] = {'b200_steps_per_s': 1 / _timeit(gpu_step, gpu_reps, sync),
                      'cpu_steps_per_s': 1 / _timeit(cpu_step, cpu_reps, nosync),
                      'shape': 'n=755, 50 contexts x 100 actions, d=6'}
    return out


if __name__ == '__main__':
    print(json.dumps(measure()))
