"""CPU restatement of one joint-MTS posterior-sampling step at fixed
hyper-parameters -- the checker for ``ocbo_b200.sweep.SweepStep`` and the CPU
baseline ``bench.py`` times.  TEST INFRASTRUCTURE ONLY.

Follows JointMTS._find_best_query (src/strategies/joint_mts.py:24-57) through
DragonflyGP.draw_sample (src/gp/gp_interface.py:76-79) with the arithmetic of
SURVEY Appendix A; candidates are the task-major grid only (SURVEY 8d sweep),
so the per-task baseline is 0.
"""
import numpy as np

from . import gp_numpy, picks


def step(X, y, Xs, T, S, scale=1.0, bandwidths=(0.25, 0.25, 0.25, 0.25, 0.05, 0.05),
         noise_var=1e-2, mu0=0.0, Z=None,
         kernel_type='se', return_samples=False):
    """Returns (task[S], action[S], gain[S]) [+ samples (S, m), mean, diag var]."""
    D = X.shape[1]
    bws = np.broadcast_to(np.asarray(bandwidths, dtype=np.float64), (D,))
    gp = gp_numpy.make_gp(X, y, kernel_type, scale, bws, noise_var, mu0)
    mu, cov = gp.eval(Xs, include_covar=True)
    L = gp_numpy.stable_cholesky(cov)
    if Z is None:
        Z = np.random.normal(size=(len(Xs), S))
    F = L.dot(Z).T + mu  # (S, m)
    out = [picks.grid_only_pick(F[s], T) for s in range(S)]
    task = np.asarray([o[0] for o in out], dtype=np.int32)
    act = np.asarray([o[1] for o in out], dtype=np.int32)
    gain = np.asarray([o[2] for o in out])
    if return_samples:
        return task, act, gain, F, mu, np.diag(cov)
    return task, act, gain
