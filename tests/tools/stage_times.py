"""In-flight cost of each stage of the sweep step: the 8 groups (streams) x batch 2 of bench.py run ONE stage over
and over (inputs of a finished step, so every stage sees valid data); ms per trial = wall / (reps * 16).
    python tests/tools/stage_times.py > gpurun_out/r02_stage_times.json"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ocbo_b200 import sweep, _lib, engine
    from ocbo_b200._lib import call
    from ocbo_b200.engine import _p
    G, B = 8, 2
    streams = [torch.cuda.Stream() for _ in range(G)]
    groups = [sweep.SweepStep(1024, 64, 128, 256, 6, batch=B, stream=s) for s in streams]
    for g, grp in enumerate(groups):
        ids = [g * B + b for b in range(B)]
        grp.set_inputs_device(ids, [sweep.make_trial_inputs(t) for t in ids])
        grp.run()
    torch.cuda.synchronize()

    def st_of(grp):
        return ctypes.c_void_p(grp.stream.cuda_stream)

    def postcov(grp):
        n, m, D = grp.n, grp.m, grp.D
        call('ocbo_post_cov_ws', grp.kind, D, _p(grp.Xs), m, grp.Xs.stride(0), None, _p(grp.V), n, m, grp.V.stride(0),
             _p(grp.theta), grp.theta.stride(0), _p(grp.Sigma), m, grp.Sigma.stride(0), 1, B, _p(grp.ws), grp.ws.numel(),
             st_of(grp))

    # keep a copy of Sigma (the covariance, re-assembled: the step's factorisation overwrote it) of every group
    for grp in groups:
        postcov(grp)
    torch.cuda.synchronize()
    Sig0 = [grp.Sigma.clone() for grp in groups]
    torch.cuda.synchronize()

    def potrf(grp):
        grp._potrf_sigma(st_of(grp))

    def restore(grp, i):
        with torch.cuda.stream(grp.stream):
            grp.Sigma.copy_(Sig0[i], non_blocking=True)
            grp.info.zero_()

    def vsolve(grp):
        n, m, D = grp.n, grp.m, grp.D
        X, Xs, th = grp.X, grp.Xs, grp.theta
        call('ocbo_gram', grp.kind, D, _p(X), n, X.stride(0), None, _p(Xs), m, Xs.stride(0), None, _p(th), th.stride(0),
             _p(grp.V), m, grp.V.stride(0), B, 0, st_of(grp))
        call('ocbo_trsm_lower', _p(grp.LK), n, n, grp.LK.stride(0), _p(grp.invdK), _p(grp.V), m, m, grp.V.stride(0),
             _p(grp.info[0]), B, st_of(grp))

    def sample(grp):
        grp._sample_and_reduce(slice(0, B), st_of(grp), l_in_ws=True)

    def whole(grp):
        grp.run()

    def timed(fn, reps, pre=None):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            for i, grp in enumerate(groups):
                if pre:
                    pre(grp, i)
                fn(grp)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / (reps * G * B)

    out = {}
    for grp in groups:
        grp.run()
    out['whole_step'] = timed(whole, 5)
    out['copy_sigma_only'] = timed(lambda grp: None, 5, pre=restore)
    out['potrf_sigma_plus_copy'] = timed(potrf, 5, pre=restore)
    out['potrf_sigma'] = out['potrf_sigma_plus_copy'] - out['copy_sigma_only']
    out['post_cov'] = timed(postcov, 5)
    for grp in groups:
        grp.run()
    out['gram_and_solve_for_V'] = timed(vsolve, 5)
    for grp in groups:
        grp.run()
    out['sampling_tail'] = timed(sample, 5)
    out['info_after'] = [int(grp.info.abs().sum().item()) for grp in groups]
    out['unit'] = 'ms per trial, 16 trials in flight (8 streams x batch 2), one stage at a time'
    print(json.dumps(out))
    sys.stderr.write(json.dumps(out, indent=1) + '\n')


if __name__ == '__main__':
    main()
