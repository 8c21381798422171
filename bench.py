#!/usr/bin/env python
"""bench.py -- GP joint Thompson samples/s of the MTS posterior-sampling step.

Workload (BASELINE.json configs[4], SURVEY.md 8d "sweep"): per trial n=1024
observations, m=8192 candidates (T=64 tasks x A=128 actions, task-major),
d=6, S=256 Thompson draws, fp64, SE-ARD kernel at fixed hyper-parameters.
One "step" = the whole hot path (Gram -> chol -> alpha -> mean -> cross-cov ->
trsm -> Sigma -> chol(Sigma) -> Philox Z -> mu + L Z -> per-task argmax -> pick)
for ``streams x batch`` independent trials on every GPU.

  python bench.py --gpus N --steps K --warmup W           (our arm)
  python bench.py --impl reference --gpus N --steps K --warmup W
      (the reference's CPU path: numpy/LAPACK restatement in oracle/, one trial
       per step on rank 0's host cores; the reference itself has no GPU path)

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ocbo_b200.workload import (DIM, N_ACT, N_DRAW, N_OBS, N_TASK, SEED,  # noqa: E402  (numpy only:
                                make_trial_inputs, step_flops, workload_config)  # no torch, no .so)

METRIC = 'GP joint Thompson samples/sec (m=8192 grid, fp64)'
UNIT = 'samples/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--streams', type=int, default=4, help='concurrent trial groups per GPU')
    ap.add_argument('--batch', type=int, default=4, help='trials per kernel launch (grid.z)')
    ap.add_argument('--cpu-trials', type=int, default=2, help='trials timed for cpu_baseline')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-peaks', action='store_true')
    ap.add_argument('--no-one-thread', action='store_true')
    ap.add_argument('--sweep-trials', type=int, default=512,
                    help='trials of the fixed-work (strong-scaling) leg; 0 skips it')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-ladder-variant', action='store_true')
    return ap.parse_args()


# --------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def mark(self):
        """Start of the timed region: nvidia-smi is started before the warm-up (it can take longer to come up than
        a short timed region lasts), only the rows from here on are used."""
        self.t0 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        t0, t1 = getattr(self, 't0', 0.0), time.perf_counter()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.05]
        window = 'timed region'
        if not rows and self.rows:   # region shorter than the sampling period: the warm-up ran the same load
            rows, window = [r for _, r in self.rows[-5:]], 'warm-up (timed region shorter than one sample)'
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(mx)) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons), 'window': window}


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get('num_threads') for p in threadpool_info() if p.get('user_api') == 'blas']
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


def load_traffic(batch):
    """DRAM bytes (read + write) of the DMMA update launches of one trial-step, from the
    committed ncu capture (profiles/traffic.json); scaled by the batch.  None if absent."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        with open(path) as fh:
            t = json.load(fh)
        out = {'dram_bytes': t['dram_bytes'] * batch, 'launches': t['launches'],
               'scope': t['scope'], 'source': t['source']}
        if 'algorithmic_bytes' in t:
            out['algorithmic_bytes'] = t['algorithmic_bytes'] * batch
            out['dram_over_algorithmic'] = t['dram_bytes'] / t['algorithmic_bytes']
        # per launch (the capture is at batch 1; a launch of `batch` problems moves `batch` times as much)
        out['dram_bytes_per_launch'] = t['dram_bytes'] * batch / max(1, t['launches'])
        if 'algorithmic_bytes' in t:
            out['algorithmic_bytes_per_launch'] = t['algorithmic_bytes'] * batch / max(1, t['launches'])
        return out
    except Exception:
        return None


def load_measured_peaks():
    """MEASURED_PEAKS.json (driver-written: HBM GB/s, dense bf16 TF/s of this pool's B200s) or None."""
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            d = json.load(fh)
        return d if d.get('bf16_tflops') else None
    except Exception:
        return None


def timed_default(torch, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


_BLAS_LIMIT = []


def use_all_blas_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core."""
    if _BLAS_LIMIT:
        return
    try:
        from threadpoolctl import threadpool_limits
        _BLAS_LIMIT.append(threadpool_limits(limits=os.cpu_count() or 1, user_api='blas'))
    except Exception:
        _BLAS_LIMIT.append(None)


def cpu_step(trial, threads=None, with_philox=False):
    """One full-size trial-step of the CPU restatement (numpy + LAPACK).  Returns
    (seconds, (task, action, gain, samples, mean)): the timing is the cpu_baseline figure,
    the results are what the device's trial is checked against (``with_philox``: the
    oracle is fed the device RNG's definition of the normals, generated outside the clock)."""
    from oracle import philox, sweep_step
    X, y, Xs = make_trial_inputs(trial, N_OBS, N_TASK, N_ACT)
    Z = philox.normals(SEED, trial, N_TASK * N_ACT, N_DRAW) if with_philox else None
    np.random.seed(trial)
    use_all_blas_threads()
    import contextlib
    ctx = contextlib.nullcontext()
    if threads is not None:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=threads, user_api='blas')
    with ctx:
        t0 = time.perf_counter()
        out = sweep_step.step(X, y, Xs, N_TASK, N_DRAW, Z=Z, return_samples=True)
        dt = time.perf_counter() - t0
    return dt, out


def cpu_step_seconds(trial):
    return cpu_step(trial)[0]


def run_reference(args):
    """Reference arm: the reference's own algorithm for this path on the host
    cores (oracle port -- Dragonfly, which holds it, is not installable here).
    Imports nothing of the product but its numpy-only workload generator."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    for w in range(args.warmup):
        cpu_step_seconds(w)
    t = [cpu_step_seconds(args.warmup + k) for k in range(args.steps)]
    total = float(np.sum(t))
    value = N_DRAW * args.steps / total
    cores = cpu_threads()
    sample = '%d steps x 1 full-size trial (n=%d, m=%d, S=%d) per step' % (
        args.steps, N_OBS, N_TASK * N_ACT, N_DRAW)
    one = None
    if not args.no_one_thread:
        try:
            one = N_DRAW / cpu_step(args.warmup + args.steps, threads=1)[0]
        except Exception:  # threadpoolctl missing: the all-cores figure stands alone
            one = None
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': workload_config(),
        'run': {'trials_per_step': 1,
                'note': 'numpy/LAPACK restatement of the Dragonfly path (oracle/), host cores only'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': sample, 'one_thread_value': one,
                         'one_thread_sample': '1 full-size trial, BLAS limited to 1 thread'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------
def measure_fp64_peaks(torch, lib_call, dev):
    """FP64 denominators (MEASURED_PEAKS.json has no FP64 row):
    cuBLAS DGEMM 8192^3 via torch.matmul, the raw DMMA.8x8x4 / DFMA issue rate and
    this library's own DMMA tile core on a plain 8192^3 NT GEMM."""
    out = {}
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out['cublas_dgemm_tflops'] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    blocks, iters = 148 * 8, 4000
    for name, per_thread_flops in (('ocbo_probe_dmma', None), ('ocbo_probe_dfma', 16 * 2)):
        lib_call(name, ctypes.c_void_p(sink.data_ptr()), blocks, 10, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib_call(name, ctypes.c_void_p(sink.data_ptr()), blocks, iters, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if per_thread_flops is None:
            flops = blocks * 8 * iters * 16 * 512.0  # 8 warps x 16 DMMA(8x8x4 = 512 flop)
        else:
            flops = blocks * 256 * iters * float(per_thread_flops)
        out[name.replace('ocbo_probe_', '') + '_issue_tflops'] = flops / (ms * 1e-3) / 1e12
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib_call('ocbo_probe_dgemm_nt', ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                 ctypes.c_void_p(c.data_ptr()), n, st)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out['own_dgemm_nt_tflops'] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
    del a, b, c
    return out


def check_trial(torch, grp, trial, ref, acc):
    """Headline-shape parity (outside every timed region): trial ``trial`` through the product's
    host-buffer API against the oracle's results for the same inputs and the same normals.
    Picks bit-exact for every draw; mean and samples to max(1e-8, 10 cond eps) relative
    (cond(Sigma) ~ 4e5 at this shape); gains 1e-8."""
    t_ref, a_ref, g_ref, F_ref, mu_ref, _ = ref
    ids = [trial] * grp.B
    inp = make_trial_inputs(trial, N_OBS, N_TASK, N_ACT)
    task, act, gain = grp.step_host(ids, [inp] * grp.B)
    torch.cuda.synchronize()
    mean = grp.mean[0].cpu().numpy()
    e_mean = float(np.max(np.abs(mean - mu_ref)) / np.max(np.abs(mu_ref)))
    e_gain = float(np.max(np.abs(gain[0] - g_ref)) / np.max(np.abs(g_ref)))
    e_samp = None
    F = grp.samples_host(0) if hasattr(grp, 'samples_host') else None
    if F is not None:
        e_samp = float(np.max(np.abs(F - F_ref)) / np.max(np.abs(F_ref)))
    tol = max(1e-8, 10 * 4.3e5 * 2.2e-16)
    n_task = int(np.sum(task[0] != t_ref))
    n_act = int(np.sum(act[0] != a_ref))
    ok = (n_task == 0 and n_act == 0 and e_mean <= tol and e_gain <= tol
          and (e_samp is None or e_samp <= tol))
    acc = acc or {'ok': True, 'trials': [], 'draws_compared': 0, 'task_mismatches': 0,
                  'action_mismatches': 0, 'max_rel_err_mean': 0.0, 'max_rel_err_gain': 0.0,
                  'max_rel_err_samples': None, 'tolerance': tol,
                  'against': 'oracle.sweep_step.step fed oracle.philox.normals, same trial inputs, '
                             'n=%d m=%d S=%d' % (N_OBS, N_TASK * N_ACT, N_DRAW)}
    acc['ok'] = bool(acc['ok'] and ok)
    acc['trials'].append(trial)
    acc['draws_compared'] += int(len(t_ref))
    acc['task_mismatches'] += n_task
    acc['action_mismatches'] += n_act
    acc['max_rel_err_mean'] = max(acc['max_rel_err_mean'], e_mean)
    acc['max_rel_err_gain'] = max(acc['max_rel_err_gain'], e_gain)
    if e_samp is not None:
        acc['max_rel_err_samples'] = max(acc['max_rel_err_samples'] or 0.0, e_samp)
    return acc


def run_ladder_variant(torch, sweep, streams, B, args, steps=3, rungs=3):
    """SURVEY 8d's LITERAL configuration -- isotropic bandwidth 0.25 -- beside the headline: Sigma is
    numerically singular (cond ~ 1e16), so every trial walks stable_cholesky's jitter ladder.  The
    rungs are enqueued on the device inside the step (SweepStep(ladder_rungs=...)): same host-buffer
    API, H2D / D2H in the clock, no host round trip per rung."""
    from ocbo_b200.workload import ISOTROPIC_BANDWIDTHS
    try:
        groups = [sweep.SweepStep(N_OBS, N_TASK, N_ACT, N_DRAW, DIM, batch=B, stream=s,
                                  bandwidths=ISOTROPIC_BANDWIDTHS, ladder_rungs=rungs) for s in streams]
        G = len(groups)
        ids = lambda step, g: [10000 + (step * G + g) * B + b for b in range(B)]
        pool = {t: make_trial_inputs(t, N_OBS, N_TASK, N_ACT)
                for step in range(steps + 1) for g in range(G) for t in ids(step, g)}
        hist, redone = {}, 0

        def one(step):
            nonlocal redone
            for g, grp in enumerate(groups):
                grp.submit_host(ids(step, g), [pool[t] for t in ids(step, g)])
            for grp in groups:
                grp.result_host()
                for _, r in grp.jitter:
                    hist[str(r)] = hist.get(str(r), 0) + 1
                redone += int(np.count_nonzero(grp.hinfo.numpy().any(axis=0)))
        one(0)
        hist.clear()
        redone = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            one(1 + k)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        del groups
        torch.cuda.empty_cache()
        return {'workload': 'as config.workload but SE-ARD bandwidth 0.25 on all 6 dims (SURVEY 8d literal)',
                'value': G * B * N_DRAW * steps / wall, 'unit': UNIT, 'ms_per_step': 1e3 * wall / steps,
                'trials_per_step': G * B, 'steps': steps, 'device_ladder_rungs_enqueued': rungs,
                'jitter_rung_histogram': hist, 'trials_redone_from_the_host': redone,
                'timing': 'wall clock around submit_host/result_host (host buffers, copies inside)'}
    except Exception as exc:  # secondary figure: never lose the headline line
        return {'error': repr(exc)}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from ocbo_b200 import engine, sweep
    from ocbo_b200._lib import call as lib_call

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl b200 needs a CUDA device; there is no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    G, B = args.streams, args.batch
    trials_per_step = G * B
    m = N_TASK * N_ACT
    streams = [torch.cuda.Stream() for _ in range(G)]
    groups = [sweep.SweepStep(N_OBS, N_TASK, N_ACT, N_DRAW, DIM, batch=B, stream=s) for s in streams]

    def trial_id(step, g, b):  # trial r lives on rank r mod W
        return ((step * G + g) * B + b) * world + rank

    total_steps = args.warmup + args.steps
    pool = {}
    for step in range(total_steps):
        for g in range(G):
            for b in range(B):
                t = trial_id(step, g, b)
                pool[t] = make_trial_inputs(t, N_OBS, N_TASK, N_ACT)

    def ids(step, g):
        return [trial_id(step, g, b) for b in range(B)]

    # ---- (1) device-resident throughput: inputs already in HBM -----------------
    for g, grp in enumerate(groups):
        grp.set_inputs_device(ids(0, g), [pool[t] for t in ids(0, g)])
    main = torch.cuda.current_stream()

    def run_resident(nsteps):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record(main)
        for s in streams:
            s.wait_event(start)
        for _ in range(nsteps):
            for grp in groups:
                grp.run()
        for s in streams:
            ev = torch.cuda.Event()
            ev.record(s)
            main.wait_event(ev)
        end.record(main)
        barrier()
        return start.elapsed_time(end)

    sampler = ClockSampler(local)
    sampler.start()
    run_resident(args.warmup)
    sampler.mark()
    from ocbo_b200._lib import LIB
    launches0 = int(LIB.ocbo_launch_count())
    ms_resident = run_resident(args.steps)
    launches = int(LIB.ocbo_launch_count()) - launches0
    clocks = sampler.stop()

    # host time to ISSUE one group's launches from a drained device (outside the clocks): x G = the
    # host's share of a step; well below ms_per_step means the device, not the issuing thread, bounds it
    issue = []
    for grp in groups[:4]:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        grp.run()
        issue.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    host_issue_ms = float(np.median(issue)) * G

    # ---- (2) end to end: host buffers in, picks out, every step ----------------
    def run_e2e(first_step, nsteps):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        start.record(main)
        for s in streams:
            s.wait_event(start)
        picks = []
        for k in range(nsteps):
            step = first_step + k
            for g, grp in enumerate(groups):
                if k > 0:
                    picks.append(grp.result_host())  # D2H result of the previous step
                grp.submit_host(ids(step, g), [pool[t] for t in ids(step, g)])
        for grp in groups:
            picks.append(grp.result_host())
        for s in streams:
            ev = torch.cuda.Event()
            ev.record(s)
            main.wait_event(ev)
        end.record(main)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        return max(start.elapsed_time(end), wall), picks

    run_e2e(0, args.warmup)
    ms_e2e, picks = run_e2e(args.warmup, args.steps)
    checksum = int(sum(int(p[0].sum()) + int(p[1].sum()) for p in picks))

    # ---- max over ranks ---------------------------------------------------------
    times = torch.tensor([ms_resident, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_resident, ms_e2e = [float(x) for x in times.tolist()]
    samples_per_step = world * trials_per_step * N_DRAW
    value = samples_per_step * args.steps / (ms_resident * 1e-3)
    e2e_value = samples_per_step * args.steps / (ms_e2e * 1e-3)

    # ---- fixed work: the SAME --sweep-trials trials at every N (SURVEY 8e, configs[4]) -------
    # trial r on rank r mod W through the host-buffer API, the one NCCL gather of the picks
    # INSIDE the clock; the digest over all trials x draws must not depend on N.
    sweep_fixed = None
    if args.sweep_trials > 0:
        from ocbo_b200 import dist as odist
        ntr = args.sweep_trials
        mine = {t: make_trial_inputs(t, N_OBS, N_TASK, N_ACT) for t in odist.shard_trials(ntr, rank, world)}
        # warm-up of the gather itself (same collective, same buffer shape, no trials): the first NCCL gather of a
        # process sets up its send / receive channels -- 1.2-1.4 s at 8 ranks, more than the 512 trials take
        odist.gather_picks([], np.zeros((0, N_DRAW), np.int32), np.zeros((0, N_DRAW), np.int32),
                           np.zeros((0, N_DRAW)), ntr)
        barrier()
        t0 = time.perf_counter()
        res = odist.run_sweep(ntr, None, mine.__getitem__, groups=groups)
        barrier()
        wall = time.perf_counter() - t0
        wt = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(wt, op=dist.ReduceOp.MAX)
        if rank == 0:
            import hashlib
            task_all, act_all, gain_all = res
            h = hashlib.sha256()
            h.update(np.ascontiguousarray(task_all).tobytes())
            h.update(np.ascontiguousarray(act_all).tobytes())
            sweep_fixed = {'trials': ntr, 'draws_per_trial': N_DRAW, 'wall_s': float(wt[0]),
                           'samples_per_s': ntr * N_DRAW / float(wt[0]),
                           'picks_sha256_16': h.hexdigest()[:16],
                           'picks_sum': int(task_all.astype(np.int64).sum() + act_all.astype(np.int64).sum()),
                           'scaling': 'strong (fixed %d trials, trial r on rank r mod W, gather inside '
                                      'the clock); the digest is identical at every N' % ntr,
                           'survey_target_s': {1: 5.9, 8: 0.85}.get(world)}
        del mine

    # ---- the option-file shapes at N > 1: every rank runs its own independent decisions ------
    # (BO trials shard with no collective, SURVEY 8e: the aggregate is the sum of the ranks' rates)
    small_multi = None
    if world > 1 and not args.no_cpu_baseline:
        try:
            from tests.tools import bench_small
            names = ['set2d', 'rand6d', 'jointbran', 'conth42']
            barrier()
            mine = bench_small.measure(cpu=False)
            rates = torch.tensor([mine[k]['b200_steps_per_s'] for k in names], dtype=torch.float64,
                                 device=dev)
            lo = rates.clone()
            dist.all_reduce(rates, op=dist.ReduceOp.SUM)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            small_multi = {k: {'b200_steps_per_s': float(rates[i]), 'slowest_rank_steps_per_s': float(lo[i]),
                               'n_gpus': world, 'shape': mine[k]['shape']} for i, k in enumerate(names)}
        except Exception as exc:
            small_multi = {'error': repr(exc)}

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernels, CUDA events on the group's own stream ------
        grp = groups[0]
        grp.run()
        torch.cuda.synchronize()
        st = ctypes.c_void_p(grp.stream.cuda_stream)
        vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        oz = grp.ws is not None
        ws_bytes = grp.ws.numel() if oz else 0

        def postcov():
            lib_call('ocbo_post_cov_ws', grp.kind, DIM, vp(grp.Xs), m, grp.Xs.stride(0), None, vp(grp.V), N_OBS, m,
                     grp.V.stride(0), vp(grp.theta), grp.theta.stride(0), vp(grp.Sigma), m, grp.Sigma.stride(0), 1, B,
                     vp(grp.ws), ws_bytes, st)
        # re-assemble Sigma (the factorisation above overwrote it) and time its kernels per launch class
        postcov()
        grp.info.zero_()
        torch.cuda.synchronize()
        ms7 = (ctypes.c_float * 7)()
        lib_call('ocbo_potrf_ws_profile', vp(grp.Sigma), m, m, grp.Sigma.stride(0), vp(grp.invdS), vp(grp.info[1]), B,
                 vp(grp.ws), ws_bytes, st, ms7)
        nt = m // 128
        # algorithmic flops of the update launches of one chol(Sigma): m^3/3 minus the
        # diagonal-block factorisations (128^3/3 each) and the panel solves (rows * 128^2)
        alg = (m ** 3 / 3.0 - nt * 128.0 ** 3 / 3.0
               - sum(r * 128.0 * 128.0 ** 2 for r in range(1, nt))) * B
        ms_upd = ms7[2] + ms7[3] + ms7[4] + ms7[5] + ms7[6]
        all_updates_tflops = alg / (ms_upd * 1e-3) / 1e12 if ms_upd > 0 else None
        # the rank-`outer` trailing updates (one per outer panel): algorithmic flops = lower triangle incl. the
        # diagonal of the trailing block, 2 k each
        outer = int(os.environ.get('OCBO_POTRF_OUTER', 2048))
        ot = max(outer // 128, 1)
        alg_outer = 0.0
        for j1 in range(ot, nt, ot):
            rows = (nt - j1) * 128.0
            alg_outer += rows * (rows + 1.0) / 2.0 * 2.0 * (ot * 128.0)
        alg_outer *= B
        ms_outer = ms7[5] if oz else ms7[3]
        trailing_tflops = alg_outer / (ms_outer * 1e-3) / 1e12 if ms_outer > 0 else all_updates_tflops
        peaks = {} if args.no_peaks else measure_fp64_peaks(torch, lib_call, dev)
        peak = peaks.get('cublas_dgemm_tflops')
        step_tflops = step_flops(N_OBS, m, N_DRAW) * trials_per_step * world * args.steps \
            / (ms_resident * 1e-3) / 1e12

        def timed(fn, reps=3):
            best = 1e9
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(grp.stream)
                fn()
                e1.record(grp.stream)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return best
        ms_cov = timed(postcov)
        grp.run()                       # leaves a valid factor in Sigma for the sampling product
        torch.cuda.synchronize()
        ms_lz = timed(lambda: grp._sample_and_reduce(slice(0, B), st, l_in_ws=True))  # as in run(): L's slices are in ws
        tf = lambda flops, ms: flops * B / (ms * 1e-3) / 1e12
        frac = lambda v: (v / peak) if (peak and v) else None
        cov_tf, lz_tf = tf(float(N_OBS) * m * m, ms_cov), tf(float(m) * m * N_DRAW, ms_lz)
        step_per_gpu = step_tflops / world
        by_class = {
            'chol_rank%d_trailing_updates' % (ot * 128): {
                'tflops': trailing_tflops, 'frac': frac(trailing_tflops),
                'share_of_update_flops': alg_outer / max(alg, 1.0), 'ms': ms_outer,
                'path': 'int8 tensor cores (ozaki_update_kernel)' if oz else 'fp64 DMMA (gemm_nt_kernel)'},
            'chol_all_update_launches': {'tflops': all_updates_tflops, 'frac': frac(all_updates_tflops),
                                         'ms': ms_upd},
            'posterior_covariance_k%d' % N_OBS: {
                'tflops': cov_tf, 'frac': frac(cov_tf), 'ms': ms_cov,
                'path': 'Gram + int8 split + ozaki_update_kernel' if oz else 'postcov_kernel (fp64 DMMA)'},
            'sample_LZ_fused_argmax': {'tflops': lz_tf, 'frac': frac(lz_tf), 'ms': ms_lz,
                                       'fused_tail': bool(grp.fused_tail)},
        }
        potrf_ms = {'diag': ms7[0], 'panel': ms7[1], 'inner_update_fp64': ms7[2], 'outer_update_fp64': ms7[3],
                    'int8_split': ms7[4], 'outer_update_int8': ms7[5], 'middle_update_int8': ms7[6], 'batch': B,
                    'outer_block': outer, 'inner': os.environ.get('OCBO_POTRF_INNER', 'left')}
        if oz:
            # Dominant kernel: ozaki_update_kernel (tcgen05.mma kind::i8).  Its roofline is the tensor pipe:
            # int8 operations ISSUED (28 slice-pair products per fp64 product, exact int32 accumulation) over the
            # CUDA-event time of the launches, against the dense int8 rate = 2 x the measured dense bf16 rate of
            # MEASURED_PEAKS.json (same tcgen05 datapath at half the operand width; cuBLASLt's own int8 GEMM,
            # measured in this run, is printed beside it).
            mp = load_measured_peaks()
            int8_peak = 2.0 * mp['bf16_tflops'] if mp else None
            int8_lib = None
            if not args.no_peaks:
                try:
                    a8 = torch.randint(-64, 64, (8192, 8192), dtype=torch.int8, device=dev)
                    torch._int_mm(a8, a8.t())
                    ms_i8 = timed_default(torch, lambda: torch._int_mm(a8, a8.t()))
                    int8_lib = 2.0 * 8192.0 ** 3 / (ms_i8 * 1e-3) / 1e12
                    del a8
                except Exception:
                    int8_lib = None
            if int8_peak is None:
                int8_peak = int8_lib
            pairs = int(LIB.ocbo_ozaki_slice_pairs())
            tops = pairs * trailing_tflops if trailing_tflops else None
            roof = {
                'bound': 'tensor',
                'scope': 'ozaki_update_kernel, the %d rank-%d trailing updates of one chol(Sigma) launch group '
                         '(batch %d), CUDA events on the launching stream' % (max(nt // ot - 1, 0), ot * 128, B),
                'achieved': tops, 'peak': int8_peak, 'unit': 'TOP/s (int8, dense)',
                'frac': (tops / int8_peak) if (tops and int8_peak) else None,
                'ops_counted': 'int8 multiply-adds x 2 of the %d slice-pair products per fp64 product (7 balanced '
                               'radix-256 slices per operand, pairs t + u <= 6): algorithmic fp64 flops (lower triangle '
                               'incl. the diagonal blocks, 2 k each) x %d' % (pairs, pairs),
                'peak_source': ('2 x bf16_tflops of MEASURED_PEAKS.json (%.1f TF/s burst; int8 runs on the same tcgen05 '
                                'pipe at twice the bf16 rate)' % mp['bf16_tflops']) if mp else
                               'cuBLASLt int8 GEMM 8192^3 measured in this run (MEASURED_PEAKS.json absent)',
                'cublaslt_int8_tops_this_run': int8_lib,
                # nominal dense 8-bit rate of a B200 (B200_PROFILING.md: 4.5 PFLOP/s fp8, int8 on the same pipe) for
                # context: the measured bf16 burst figure is itself taken at the power cap, so a kernel that moves
                # fewer operand bytes per product than the bf16 GEMM of the probe can pass 2 x that figure
                'frac_of_nominal_4500': (tops / 4500.0) if tops else None,
                'frac_note': 'frac > 1 is possible: MEASURED_PEAKS.json has no int8 row, the peak is 2 x its bf16 burst '
                             'figure, which was itself taken at the power cap; this kernel issues one MMA per P slice '
                             'against up to four stacked Q slices (fewer shared-memory operand reads per product than a '
                             'plain GEMM), so it draws less power per operation and passes it, as it passes the cuBLASLt '
                             'int8 GEMM measured in this run',
                'fp64_equivalent_tflops': trailing_tflops,
                'fp64_equivalent_over_cublas_dgemm': frac(trailing_tflops),
                'traffic': load_traffic(B),
                'dominant_kernel': 'ozaki_update_kernel (tcgen05.mma kind::i8, TMEM accumulators, bulk-copy operand '
                                   'pipeline): Cholesky trailing updates + posterior covariance, ~55 % of the device '
                                   'time of a step in flight',
                'step': {'scope': 'whole step, all kernels (algorithmic 278.12 GFLOP fp64 per trial-step)',
                         'fp64_equivalent_tflops_per_gpu': step_per_gpu, 'cublas_dgemm_tflops': peak,
                         'over_cublas_dgemm': frac(step_per_gpu)},
            }
        else:
            roof = {
                'bound': 'tensor',
                'scope': 'whole step, all kernels (algorithmic 278.12 GFLOP per trial-step)',
                'achieved': step_per_gpu, 'peak': peak, 'unit': 'TFLOP/s', 'frac': frac(step_per_gpu),
                'traffic': load_traffic(B),
                'peak_source': 'cuBLAS DGEMM 8192^3 (torch.matmul fp64) measured in this run; '
                               'MEASURED_PEAKS.json has no FP64 row',
                'dominant_kernel': 'gemm_nt_kernel<128,64> (DMMA.8x8x4 fp64): Cholesky rank-k updates, '
                                   '~55 % of the device time of a step',
            }
        roof.update({
            'step_tflops_per_gpu': step_per_gpu, 'step_frac_of_peak': frac(step_per_gpu),
            # launches of ONE group (batch trials) serialised on an otherwise idle GPU: short-k / small-grid
            # classes cannot fill 148 SMs alone -- in the timed region the other streams do;
            # tflops = fp64-equivalent (algorithmic fp64 flops / time), frac = over cuBLAS DGEMM of this run
            'by_launch_class': by_class,
            'potrf_ms': potrf_ms,
            'operand_loads': 'cp.async' if os.environ.get('OCBO_TMA', '1') == '0' else 'TMA (cp.async.bulk.tensor)',
            'all_update_launches_tflops': all_updates_tflops,
            'fp64_probes_tflops': peaks,
        })
        ladder_variant = None
        if world == 1 and not args.no_ladder_variant:
            ladder_variant = run_ladder_variant(torch, sweep, streams, B, args)
        cpu, parity = None, None
        if not args.no_cpu_baseline and world == 1:
            cpu_step_seconds(0)  # warm BLAS threads
            ts = []
            for k in range(args.cpu_trials):
                # the timed CPU trials double as the checker of the device's results on the
                # SAME full-size trial (the oracle is fed the device RNG's normals)
                dt, ref = cpu_step(1 + k, with_philox=not args.no_parity)
                ts.append(dt)
                if not args.no_parity:
                    parity = check_trial(torch, groups[0], 1 + k, ref, parity)
            cpu = {'value': N_DRAW * len(ts) / float(np.sum(ts)), 'unit': UNIT,
                   'cores': cpu_threads(), 'kind': 'port',
                   'sample': '%d full-size trials (n=%d, m=%d, S=%d), numpy+LAPACK oracle, '
                             '1 warm-up trial' % (len(ts), N_OBS, m, N_DRAW)}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_resident / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': workload_config(),
            'run': {'trials_per_step_per_gpu': trials_per_step, 'streams': G, 'batch': B,
                    'parallelism': 'trials sharded r mod W, no collective inside the step, one gather '
                                   'of picks', 'rng': 'device Philox4x32-10 keyed (seed, trial)',
                    'host_issue_ms_per_step': host_issue_ms,
                    'launches_per_step': launches / max(1, args.steps)},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT,
                    'h2d_bytes_per_step': sum(g.h2d_bytes for g in groups),
                    'd2h_bytes_per_step': sum(g.d2h_bytes for g in groups),
                    'ms_per_step': ms_e2e / args.steps, 'picks_checksum': checksum},
            'gpu_launches': launches,
            'roofline': roof,
            'cpu_baseline': cpu,
            # (the CPU checker runs on rank 0 at N = 1 only; at N > 1 the fixed-trials digest below, equal
            # to N = 1's, ties the picks to the checked ones)
            'parity_checked': bool(parity and parity['ok']) if world == 1 and parity is not None else None,
            'parity': parity,
            'sweep_fixed_trials': sweep_fixed,
            'ladder_variant': ladder_variant,
        }
        if not args.no_cpu_baseline and world == 1:
            # second half of BASELINE.json's metric: MTS BO decision steps/s at the option-file
            # shapes (fixed HPs), batched device pass vs the numpy oracle loop on the host cores
            try:
                from tests.tools import bench_small
                use_all_blas_threads()
                line['mts_bo_steps'] = bench_small.measure()
            except Exception as exc:  # secondary figure: never lose the headline line
                line['mts_bo_steps'] = {'error': repr(exc)}
        elif small_multi is not None:
            line['mts_bo_steps'] = small_multi
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
