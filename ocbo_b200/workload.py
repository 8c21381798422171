"""Synthetic inputs of the sweep workload (BASELINE.json configs[4]).

numpy only -- no torch, no ``_lib``: ``bench.py --impl reference`` (the CPU arm) builds
its inputs from here without mapping the product's shared library."""
import numpy as np

SEED = 100826730  # the reference's global seed (src/ocbo.py:74-75)
# SE-ARD bandwidths of the sweep: 0.25 on the 4 context dims, 0.05 on the 2 action dims.
# (SURVEY 8d quotes an isotropic 0.25 "to keep cond(Sigma) <~ 1e6"; measured, that choice gives
# cond ~ 1e16 and fires the jitter ladder on every trial, so the headline action bandwidth is set
# to meet the stated conditioning goal instead: cond(Sigma) ~ 1e4-1e5.  The literal isotropic
# configuration is benched beside it as the "ladder" variant.  DESIGN.md "sweep configuration".)
SWEEP_BANDWIDTHS = (0.25, 0.25, 0.25, 0.25, 0.05, 0.05)
ISOTROPIC_BANDWIDTHS = (0.25,) * 6
N_OBS, N_TASK, N_ACT, N_DRAW, DIM = 1024, 64, 128, 256, 6


def step_flops(n, m, S):
    """Algorithmic FLOPs of one trial-step."""
    return (n ** 3 / 3.0 + float(n) ** 2 * m + float(n) * m ** 2 + m ** 3 / 3.0
            + float(m) ** 2 * S + 2.0 * m * n + 2.0 * n ** 2)


def hartmann6(x):
    A = np.asarray([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14],
                    [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]])
    P = 1e-4 * np.asarray([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                           [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])
    alpha = np.asarray([1.0, 1.2, 3.0, 3.2])
    x = np.atleast_2d(x)
    e = -np.einsum('ij,nij->ni', A, (x[:, None, :] - P[None]) ** 2)
    return (alpha * np.exp(e)).sum(axis=1)


def make_trial_inputs(trial, n=1024, T=64, A=128, ctx_dim=4, act_dim=2):
    """Synthetic inputs of trial ``trial``: X ~ U[0,1]^(n x 6),
    y = hartmann6(X) + N(0, 1e-4), candidates = task-major grid of T task
    locations ~ U[0,1]^4 times A actions ~ U[0,1]^2 (misc_util.py:104-116)."""
    g = np.random.default_rng(trial)
    D = ctx_dim + act_dim
    X = g.uniform(size=(n, D))
    y = (hartmann6(X) if D == 6 else np.sin(3.0 * X.sum(axis=1))) + 1e-2 * g.standard_normal(n)
    locs = g.uniform(size=(T, ctx_dim))
    acts = g.uniform(size=(T * A, act_dim))
    Xs = np.hstack([np.repeat(locs, A, axis=0), acts])
    return X, y, Xs


def workload_name(n=N_OBS, T=N_TASK, A=N_ACT, S=N_DRAW, dim=DIM):
    return ('sweep: n=%d obs, m=%d candidates (T=%d x A=%d), d=%d, S=%d draws/trial, SE-ARD '
            'bw ctx 0.25 / act 0.05, scale 1, noise 1e-2, mu0 0' % (n, T * A, T, A, dim, S))


def workload_config():
    """The ``config`` object of the bench line -- identical in both arms of bench.py."""
    return {'workload': workload_name(),
            'trial_inputs': 'ocbo_b200.workload.make_trial_inputs(trial): numpy default_rng(trial)',
            'trial_ids': 'step-major, trial r on rank r mod W',
            'l2': 'working set per trial (Sigma 512 MiB) exceeds the 126 MB L2; no flush needed'}
