"""Where does one MTS decision step spend its time at the option-file shapes?
Host profile (cProfile, cumulative) and device kernel totals (CUPTI through
torch.profiler, which also sees the kernels launched through the C ABI).
Usage: python tests/tools/profile_small.py [set2d|rand6d|jointbran|conth42] [reps]"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def steps():
    """The four gpu_step closures of bench_small.measure (same shapes)."""
    from ocbo_b200 import batched, synth
    from ocbo_b200.gp.gp_interface import make_gp
    from tests.tools.bench_small import _hp
    rs = np.random.RandomState(0)
    np.random.seed(0)
    out = {}
    for name, T, d, nf in (('set2d', 5, 2, 50), ('rand6d', 30, 6, 55)):
        Xl = [rs.uniform(size=(nf, d)) for _ in range(T)]
        yl = [np.sin(3 * X.sum(1)) + 0.01 * rs.normal(size=nf) for X in Xl]
        hl = [_hp(X, y) for X, y in zip(Xl, yl)]
        gb = [make_gp(X, y, 'se', h['scale'], h['bw'], h['noise'], h['mu0'], build=False)
              for X, y, h in zip(Xl, yl, hl)]

        def gpu_step(Xl=Xl, gb=gb, T=T, d=d, nf=nf):
            pts = [np.vstack([X, np.random.uniform(size=(100, d))]) for X in Xl]
            Z = [np.random.normal(size=(len(p), 1)) for p in pts]
            return batched.mts_step(gb, pts, [nf] * T, Z)
        out[name] = gpu_step
    T, A, n = 10, 100, 160
    X = np.hstack([np.repeat(np.linspace(0, 1, T), n // T)[:, None], rs.uniform(size=(n, 1))])
    y = np.asarray([synth.branin(x) for x in X])
    h = _hp(X, y)
    gbj = make_gp(X, y, 'se', h['scale'], h['bw'], h['noise'], h['mu0'], build=False)
    seg = np.concatenate([np.arange(0, n + 1, n // T), n + A * np.arange(1, T + 1)]).astype(np.int32)
    mkgrid = lambda: np.hstack([np.repeat(np.linspace(0, 1, T), A)[:, None],
                                np.random.uniform(size=(T * A, 1))])

    def joint_step():
        gbj.gp_core._post = None
        return gbj.gp_core.sample_and_pick(np.vstack([X, mkgrid()]), seg, T)
    out['jointbran'] = joint_step
    n, P = 755, 50
    Xc = rs.uniform(size=(n, 6))
    yc = np.asarray([synth.hartmann6(x) for x in Xc])
    h = _hp(Xc, yc)
    gbc = make_gp(Xc, yc, 'se', h['scale'], h['bw'], h['noise'], h['mu0'], build=False)

    def cts_step():
        acts = [np.hstack([np.tile(c, (100, 1)), np.random.uniform(size=(100, 2))])
                for c in np.random.uniform(size=(P, 4))]
        gbc.gp_core._post = None
        return batched.profile_step(gbc, acts, [np.random.normal(size=(100, 1)) for _ in acts])
    out['conth42'] = cts_step
    return out


def main():
    import torch
    from ocbo_b200 import engine
    names = [a for a in sys.argv[1:] if not a.isdigit()] or ['set2d', 'rand6d', 'jointbran', 'conth42']
    reps = next((int(a) for a in sys.argv[1:] if a.isdigit()), 20)
    fs = steps()
    for name in names:
        f = fs[name]
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        l0 = engine.LAUNCHES[0]
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print('=== %s: %.3f ms/step (%.0f steps/s), %d counted launches/step' % (
            name, dt * 1e3, 1 / dt, (engine.LAUNCHES[0] - l0) // reps))
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(reps):
            f()
        torch.cuda.synchronize()
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28)
        print('\n'.join(l for l in s.getvalue().splitlines()[4:] if l.strip())[:6000])
        try:
            from torch.profiler import profile, ProfilerActivity
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                for _ in range(reps):
                    f()
                torch.cuda.synchronize()
            rows = [(e.key, e.device_time_total / reps, e.count / reps) for e in prof.key_averages()
                    if e.device_time_total > 0]
            rows.sort(key=lambda r: -r[1])
            tot = sum(r[1] for r in rows)
            print('--- device time per step: %.1f us' % tot)
            for k, us, c in rows[:16]:
                print('  %8.1f us  x%5.1f  %s' % (us, c, k[:90]))
        except Exception as e:  # profiling is best effort
            print('torch.profiler unavailable:', e)


if __name__ == '__main__':
    main()
