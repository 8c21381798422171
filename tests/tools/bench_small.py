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


def _cpu_best(f, reps):
    """Oracle steps/s at its best BLAS thread count: the option-file shapes are small, and
    LAPACK on 16+ threads is several times SLOWER than on one for them (set2d: 25 vs 330
    steps/s); the baseline gets whichever is faster.  Returns (steps/s, threads)."""
    best = (0.0, None)
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:
        return 1.0 / _timeit(f, reps, lambda: None), 'default'
    for limit in (1, None):
        with threadpool_limits(limits=limit):
            rate = 1.0 / _timeit(f, reps, lambda: None)
        if rate > best[0]:
            best = (rate, 1 if limit == 1 else (os.cpu_count() or 1))
    return best


def _hp(X, y):
    v = max(float(np.var(y)), 1e-6)
    return dict(scale=v, bw=0.3 * (X.max(0) - X.min(0) + 1e-3), noise=0.01 * v, mu0=float(np.median(y)))


FP64_CORE_TFLOPS = 33.5   # DFMA issue rate measured on this pool's B200 (bench.py ocbo_probe_dfma)
HBM_GBS = 6551.0          # MEASURED_PEAKS.json


def _step_work(name):
    """Algorithmic flops and minimum HBM bytes of ONE decision at the bench shapes (symmetric /
    triangular aware, as BASELINE.md section 4): chol(K) n^3/3, V = L^-1 K n^2 m, Sigma n m^2,
    chol(Sigma) m^3/3, sample m^2 -- per problem."""
    per = lambda n, m: n ** 3 / 3.0 + float(n) ** 2 * m + float(n) * m ** 2 + m ** 3 / 3.0 + float(m) ** 2
    byts = lambda n, m: 8.0 * (n * n / 2.0 + n * m + m * m / 2.0) * 2  # factor + V + Sigma written and read once
    if name == 'set2d':
        return 5 * per(50, 150), 5 * byts(50, 150)
    if name == 'rand6d':
        return 30 * per(55, 155), 30 * byts(55, 155)
    if name == 'jointbran':
        return per(160, 1160), byts(160, 1160)
    n = 755  # conth42: one posterior, 50 contexts x 100 actions
    return n ** 3 / 3.0 + 50 * (float(n) ** 2 * 100 + n * 100.0 ** 2 + 100 ** 3 / 3.0 + 100.0 ** 2), \
        8.0 * (n * n + 50 * (n * 100 + 100 * 100)) * 2


def _limiter(name, wall_ms):
    """What bounds a decision at this shape: device time of the step's CUDA graph replayed back to
    back (no host work between), the host share of the wall time, kernels per step, and the device
    time against the FP64 CUDA-core and HBM rooflines.  These steps are chains of 25-60 dependent
    kernels on matrices of 128-1280 rows: the bound that binds is launch / dependency LATENCY
    (device_ms / kernels = a few us per kernel), not a throughput roofline."""
    from ocbo_b200 import engine
    plan = engine.StepPlan.last_run
    if plan is None or plan.graph is None:
        return {}
    dev_ms = plan.device_ms_per_replay()
    flops, byts = _step_work(name)
    t_flop, t_hbm = flops / (FP64_CORE_TFLOPS * 1e12) * 1e3, byts / (HBM_GBS * 1e9) * 1e3
    return {'wall_ms': wall_ms, 'device_ms': dev_ms, 'host_ms': max(wall_ms - dev_ms, 0.0),
            'kernels_per_step': plan.kernels, 'us_per_kernel': 1e3 * dev_ms / max(plan.kernels, 1),
            'algorithmic_gflop': flops / 1e9,
            'fp64_core_bound_ms': t_flop, 'hbm_bound_ms': t_hbm,
            'device_frac_of_fp64_core_roofline': t_flop / dev_ms, 'device_frac_of_hbm_roofline': t_hbm / dev_ms,
            'limiter': 'latency: %d dependent kernels at %.1f us each; host staging + launch is %.0f %% of the wall time'
                       % (plan.kernels, 1e3 * dev_ms / max(plan.kernels, 1), 100.0 * max(wall_ms - dev_ms, 0.0) / wall_ms)}


def measure(gpu_reps=20, cpu_reps=3, cpu=True):
    import torch
    from ocbo_b200 import batched, synth
    from ocbo_b200.gp.gp_interface import make_gp
    from oracle import gp_numpy as o, picks
    sync = torch.cuda.synchronize
    nosync = lambda: None
    out, jobs = {}, []
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

        def gpu_step(Xl=Xl, gb=gb, T=T, d=d, nf=nf):
            pts = [np.vstack([X, np.random.uniform(size=(100, d))]) for X in Xl]
            Z = [np.random.normal(size=(len(p), 1)) for p in pts]
            return batched.mts_step(gb, pts, [nf] * T, Z)

        def cpu_step(Xl=Xl, go=go, T=T, d=d, nf=nf):
            res = []
            for g, X in zip(go, Xl):
                g.gp_core.build_posterior()  # get_next_gp always rebuilds (gp_interface.py:132)
                res.append(g.draw_sample(samp_pts=np.vstack([X, np.random.uniform(size=(100, d))])))
            return picks.mts_pick(res, [nf] * T)
        jobs.append((name, gpu_step, cpu_step, 'T=%d, d=%d, n_f=%d, A=100' % (T, d, nf)))
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

    def gpu_step(X=X, T=T):
        gbj.gp_core._post = None
        return gbj.gp_core.sample_and_pick(np.vstack([X, mkgrid()]), seg, T)

    def cpu_step(X=X, T=T, n=n):
        goj.gp_core.build_posterior()
        s = goj.draw_sample(samp_pts=np.vstack([X, mkgrid()]))
        return picks.joint_mts_pick(s, [n // T] * T, T)
    jobs.append(('jointbran', gpu_step, cpu_step, 'T=10, n=160, m=1160'))
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
    jobs.append(('conth42', gpu_step, cpu_step, 'n=755, 50 contexts x 100 actions, d=6'))
    # A step here is ~1 ms of which half is Python: the interpreter's cyclic collector matters.
    # Inside bench.py the heap holds ~180 k long-lived objects (sweep groups, clock samples ...)
    # and every full collection triggered by a step's temporaries walks all of them (rand6d:
    # 2.07 ms/step, 1.05 after freezing).  Freeze what exists now, as a long-running service would.
    import gc
    gc.collect()
    gc.freeze()
    # device passes first, then the host loops: BLAS worker threads keep spinning after a LAPACK
    # call and would steal the core that launches the next device step
    for name, gpu_f, _, shape in jobs:
        dt = _timeit(gpu_f, gpu_reps, sync)
        out[name] = {'b200_steps_per_s': 1 / dt, 'shape': shape}
        try:
            out[name]['limiter'] = _limiter(name, 1e3 * dt)
        except Exception as exc:  # a diagnostic, never the reason to lose the figure
            out[name]['limiter'] = {'error': repr(exc)}
    # the same decisions with R independent trials in flight: the R x T problems share ONE device
    # pass (ocbo_b200/inflight.py; aggregate decisions/s over the R trials)
    from tests.tools import bench_inflight
    for name, R in (('set2d', 64), ('rand6d', 16)):
        step = bench_inflight.make_step(name, R, np.random.RandomState(5))
        out[name]['b200_steps_per_s_in_flight'] = R / _timeit(step, max(gpu_reps // 2, 3), sync)
        out[name]['trials_in_flight'] = R
    if not cpu:
        return out
    out['bo_loop'] = measure_bo_loop()
    out['revi'] = measure_revi(cpu_reps)
    lml = measure_lml(gpu_reps, cpu_reps)
    for name, _, cpu_f, _ in jobs:
        out[name]['cpu_steps_per_s'], out[name]['cpu_threads'] = _cpu_best(cpu_f, cpu_reps)
    for v in lml.values():
        v['cpu_evals_per_s'], v['cpu_threads'] = _cpu_best(v.pop('_cpu'), cpu_reps)
        v['cpu_evals_per_s'] *= 64
    out['lml_evals'] = lml
    return out


def measure_revi(cpu_reps=1):
    """REVI decisions/s: all candidates' knowledge gradients over all
    judgement groups in one device pass (batched.revi_improvements / ocbo_kg) against the numpy
    oracle's restatement of the reference's loops (corr_opt.py:92-119, postmax_cts.py:139-160)
    timed on a SAMPLE of the candidates and scaled (the full CPU decision takes ~10 s)."""
    import torch
    from ocbo_b200 import batched, synth
    from ocbo_b200.gp.gp_interface import make_gp
    from oracle import gp_numpy as o, kg
    import gc
    gc.collect()
    gc.freeze()  # (see measure(): the driver loops before this leave a large heap behind)
    rs = np.random.RandomState(3)
    res = {}
    for name, n, d, C, G, E in (('jointbran', 160, 2, 1000, 10, 100), ('conth42', 755, 6, 100, 100, 100)):
        X = rs.uniform(size=(n, d))
        y = np.asarray([synth.hartmann6(x) for x in X]) if d == 6 else np.asarray([synth.branin(x) for x in X])
        h = _hp(X, y)
        gb = make_gp(X, y, 'se', h['scale'], h['bw'], h['noise'], h['mu0'])
        go = o.make_gp(X, y, 'se', h['scale'], h['bw'], h['noise'], h['mu0'])
        cand, judge = rs.uniform(size=(C, d)), rs.uniform(size=(G * E, d))
        tg = _timeit(lambda: batched.revi_improvements(gb, cand, judge, G, E), 10, torch.cuda.synchronize)
        sample = max(C // 50, 2)

        def cpu():
            means = go.eval(judge)[0].reshape(G, E)
            inter = go.get_pt_relations(cand[:sample], judge)
            cv = go.eval(cand[:sample], include_covar=True)[1].diagonal()
            return kg.revi_improvements(means, inter, cv, go.get_estimated_noise())
        tc = _timeit(cpu, cpu_reps, lambda: None) * (C / float(sample))
        res[name] = {'b200_decisions_per_s': 1 / tg, 'cpu_decisions_per_s': 1 / tc,
                     'shape': 'n=%d, %d candidates x %d groups x %d judgement points' % (n, C, G, E),
                     'cpu_sample': '%d of %d candidates, scaled' % (sample, C)}
    return res


def measure_bo_loop(capital=40, in_flight=(8, 32), procs=8):
    """Whole BO iterations/s through the drivers that read the reference's option files
    (ocbo_b200.ocbo / cts_ocbo): hyper-parameter re-fit (64-candidate LML pass) + posterior draw
    + decision + objective evaluation + bookkeeping every iteration, one method, one trial -- and,
    for the independent-GP files, with R trials in flight (``--trials R --trials_in_flight R``:
    the trials' decisions and hyper-parameter fits share their device passes, ocbo_b200/inflight.py;
    aggregate iterations/s over the R trials).  Device-side only: the reference's own loop needs
    Dragonfly."""
    import tempfile
    import torch
    from tests.tools.profile_driver import OPTS
    res = {}
    for name, (kind, text) in OPTS.items():
        if kind == 'd':
            from ocbo_b200 import ocbo as drv
        else:
            from ocbo_b200 import cts_ocbo as drv
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, 'o.txt')
            with open(path, 'w') as fh:
                fh.write(text % capital)

            def timed(extra):
                argv = ['--options', path] + extra
                drv.run(argv=argv)  # warm-up: step plans, kernels
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                drv.run(argv=argv)
                torch.cuda.synchronize()
                return time.perf_counter() - t0
            res[name] = {'b200_iterations_per_s': capital / timed([]),
                         'iterations': capital, 'what': text.replace('\n', ' ').strip() % capital}
            if name in ('set2d', 'rand6d', 'jointbran'):
                for R in in_flight:
                    dt = timed(['--trials', str(R), '--trials_in_flight', str(R)])
                    res[name]['b200_iterations_per_s_%d_in_flight' % R] = R * capital / dt
    # ... and with several driver processes sharing this GPU (what ``torchrun --nproc-per-node P``
    # on one GPU gives, ocbo_b200/dist.py:init_from_env): the per-trial Python runs in parallel
    if procs:
        from tests.tools import bench_procs
        for name in ('set2d', 'rand6d', 'jointbran', 'conth42'):
            try:
                res[name]['processes_sharing_the_gpu'] = bench_procs.measure(name, procs, 8, capital)
            except Exception as exc:  # a measurement, not a dependency of the bench line
                res[name]['processes_sharing_the_gpu'] = {'error': repr(exc)}
    return res


def measure_lml(gpu_reps=20, cpu_reps=3):
    """LML evaluations/s over a batch of hyper-parameter candidates (SURVEY 8a row a7 / 8d:
    "LML throughput reported separately"): 64 candidates x one data set per device pass
    (Gram -> Cholesky -> solve -> LML for all candidates at once) vs the numpy oracle's loop."""
    import torch
    from ocbo_b200 import synth
    from ocbo_b200.gp import hp_search
    from oracle import gp_numpy as o
    rs = np.random.RandomState(1)
    res = {}
    for name, n, d in (('rand6d_task', 55, 6), ('jointbran', 160, 2), ('conth42', 755, 6)):
        X = rs.uniform(size=(n, d))
        y = np.asarray([synth.hartmann6(x) for x in X]) if d == 6 else np.sin(3 * X.sum(1))
        thetas = hp_search.hp_candidates(X, y, 64, uniforms=rs.uniform(size=(63, 3 + d)))
        kind = 0  # SE

        def gpu(X=X, y=y, thetas=thetas):
            return hp_search.batched_lml(X, y, kind, thetas)[0]

        def cpu(X=X, y=y, thetas=thetas):
            return [o.log_marginal_likelihood(X, y, 'se', th) for th in thetas]
        tg = _timeit(gpu, gpu_reps, torch.cuda.synchronize)
        res[name] = {'b200_evals_per_s': 64 / tg, '_cpu': cpu,
                     'shape': 'n=%d, d=%d, 64 candidates per pass' % (n, d)}
    return res


if __name__ == '__main__':
    print(json.dumps(measure()))
