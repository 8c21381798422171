"""Sweep step (BASELINE configs[4] shape family) against the oracle, and the
world-size invariance of the trial-indexed device RNG."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sweep_step_matches_oracle_small():
    import torch
    from ocbo_b200 import sweep
    from oracle import philox, sweep_step
    n, T, A, S = 256, 8, 128, 128
    step = sweep.SweepStep(n, T, A, S, 6, batch=2)
    trials = [5, 12]
    inputs = [sweep.make_trial_inputs(t, n, T, A) for t in trials]
    task, act, gain = step.step_host(trials, inputs)
    F = step.F.cpu().numpy()
    mean = step.mean.cpu().numpy()
    for b, t in enumerate(trials):
        X, y, Xs = inputs[b]
        Z = philox.normals(sweep.SEED, t, T * A, S)
        t_ref, a_ref, g_ref, F_ref, mu, var = sweep_step.step(X, y, Xs, T, S, Z=Z, return_samples=True)
        assert np.max(np.abs(mean[b] - mu)) <= 1e-8 * np.max(np.abs(mu))
        assert np.max(np.abs(F[b].T - F_ref)) <= 1e-8 * np.max(np.abs(F_ref))
        assert np.array_equal(task[b], t_ref) and np.array_equal(act[b], a_ref)
        assert np.max(np.abs(gain[b] - g_ref)) <= 1e-8 * np.max(np.abs(g_ref))


def test_sweep_step_matches_oracle_at_the_headline_shape():
    """BASELINE.json configs[4] / bench.py's workload itself: n=1024, m=8192 (64 x 128), S=256.
    Picks bit-exact for all 256 draws; mean, samples and gains to max(1e-8, 10 cond eps)
    relative (cond(Sigma) = 4.3e5 at this shape)."""
    from ocbo_b200 import sweep, workload
    from oracle import philox, sweep_step
    n, T, A, S = workload.N_OBS, workload.N_TASK, workload.N_ACT, workload.N_DRAW
    tol = max(1e-8, 10 * 4.3e5 * np.finfo(np.float64).eps)
    step = sweep.SweepStep(n, T, A, S, 6, batch=1)
    for trial in (0, 7):
        X, y, Xs = workload.make_trial_inputs(trial, n, T, A)
        task, act, gain = step.step_host([trial], [(X, y, Xs)])
        assert step.jitter[0] == (None, None)  # the headline configuration never walks the ladder
        Z = philox.normals(sweep.SEED, trial, T * A, S)
        t_ref, a_ref, g_ref, F_ref, mu, var = sweep_step.step(X, y, Xs, T, S, Z=Z, return_samples=True)
        assert np.array_equal(task[0], t_ref) and np.array_equal(act[0], a_ref)
        assert np.max(np.abs(step.mean[0].cpu().numpy() - mu)) <= tol * np.max(np.abs(mu))
        assert np.max(np.abs(step.samples_host(0) - F_ref)) <= tol * np.max(np.abs(F_ref))
        assert np.max(np.abs(gain[0] - g_ref)) <= tol * np.max(np.abs(g_ref))


def test_sweep_results_do_not_depend_on_batching_or_stream():
    """Trial r gives the same picks whether it runs alone, in a batch, or on a
    side stream -- the property that makes 1/2/4/8-GPU runs agree."""
    import torch
    from ocbo_b200 import sweep
    n, T, A, S = 128, 4, 128, 64
    inp = {t: sweep.make_trial_inputs(t, n, T, A) for t in (0, 1, 2)}
    solo = {}
    for t in inp:
        st = sweep.SweepStep(n, T, A, S, 6, batch=1)
        solo[t] = st.step_host([t], [inp[t]])
    st3 = sweep.SweepStep(n, T, A, S, 6, batch=3, stream=torch.cuda.Stream())
    task, act, gain = st3.step_host([2, 0, 1], [inp[2], inp[0], inp[1]])
    for b, t in enumerate([2, 0, 1]):
        assert np.array_equal(task[b], solo[t][0][0])
        assert np.array_equal(act[b], solo[t][1][0])
        assert np.array_equal(gain[b], solo[t][2][0])


def test_run_sweep_single_process_gather():
    from ocbo_b200 import dist as odist
    from ocbo_b200 import sweep
    n, T, A, S = 128, 4, 128, 64
    res = odist.run_sweep(5, lambda s: sweep.SweepStep(n, T, A, S, 6, batch=2, stream=s),
                          lambda t: sweep.make_trial_inputs(t, n, T, A), streams=2, batch=2)
    task, act, gain = res
    assert task.shape == (5, S) and (task >= 0).all() and (act >= 0).all() and np.isfinite(gain).all()
    st = sweep.SweepStep(n, T, A, S, 6, batch=1)
    t3 = st.step_host([3], [sweep.make_trial_inputs(3, n, T, A)])
    assert np.array_equal(task[3], t3[0][0]) and np.array_equal(act[3], t3[1][0])


def test_sweep_falls_back_to_the_jitter_ladder():
    """Isotropic 0.25 bandwidths make Sigma numerically singular (SURVEY 8d's own
    choice): the ladder-free fast path reports info != 0 and
    the trial is redone through stable_cholesky; picks stay valid and the mean
    (which does not depend on Sigma) still matches the oracle to 1e-8."""
    from ocbo_b200 import sweep
    from oracle import sweep_step
    n, T, A, S = 256, 8, 128, 64
    st = sweep.SweepStep(n, T, A, S, 6, batch=2, bandwidths=0.25)
    trials = [3, 4]
    inputs = [sweep.make_trial_inputs(t, n, T, A) for t in trials]
    task, act, gain = st.step_host(trials, inputs)
    assert any(j[1] is not None for j in st.jitter), st.jitter
    assert (task >= 0).all() and (task < T).all() and (act >= 0).all() and (act < A).all()
    assert np.isfinite(gain).all()
    mean = st.mean.cpu().numpy()
    for b, t in enumerate(trials):
        X, y, Xs = inputs[b]
        _, _, _, _, mu, _ = sweep_step.step(X, y, Xs, T, 2, bandwidths=0.25, return_samples=True)
        assert np.max(np.abs(mean[b] - mu)) <= 1e-8 * np.max(np.abs(mu))


def test_sweep_device_ladder_equals_the_host_walk():
    """SURVEY 8d's literal isotropic-0.25 configuration: Sigma is numerically singular and
    stable_cholesky's ladder is needed.  With ``ladder_rungs`` the rungs are walked on the device
    inside the step (no host round trip); rung, picks and gains must equal the host-walked redo
    (same kernels on the same data), and the oracle's ladder must land within a rung."""
    from ocbo_b200 import sweep, workload
    n, T, A, S = 256, 8, 128, 64
    trials = [3, 4]
    inputs = [workload.make_trial_inputs(t, n, T, A) for t in trials]
    host = sweep.SweepStep(n, T, A, S, 6, batch=2, bandwidths=0.25)
    t0, a0, g0 = host.step_host(trials, inputs)
    dev = sweep.SweepStep(n, T, A, S, 6, batch=2, bandwidths=0.25, ladder_rungs=4)
    t1, a1, g1 = dev.step_host(trials, inputs)
    assert all(j[1] is not None for j in dev.jitter), dev.jitter
    assert [j[1] for j in dev.jitter] == [j[1] for j in host.jitter]
    assert np.array_equal(t0, t1) and np.array_equal(a0, a1) and np.array_equal(g0, g1)
    # a step that needs more rungs than were enqueued falls back to the host walk
    short = sweep.SweepStep(n, T, A, S, 6, batch=2, bandwidths=0.25, ladder_rungs=1)
    t2, a2, g2 = short.step_host(trials, inputs)
    assert np.array_equal(t0, t2) and np.array_equal(a0, a2) and np.array_equal(g0, g2)
    assert [j[1] for j in short.jitter] == [j[1] for j in host.jitter]
