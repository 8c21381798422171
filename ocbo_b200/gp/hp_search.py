"""Hyper-parameter fitting on top of the batched LML kernel (SURVEY 8a row a7,
8f row f1).

Dragonfly's optimiser (DiRect / PDOO / slice sampling) is not reproducible here
(dependency absent), so the search is defined as a batched
random search: all candidates are evaluated in ONE batched device pass
(Gram -> Cholesky -> solve -> LML for every candidate at once).

  'ml'            arg-max of the LML over the candidates
  'post_sampling' sampling-importance-resampling of ``num_samples`` candidates
                  with weights exp(LML - max LML)
  'a-b'           the criteria alternate between successive fits

theta = (scale, noise_var, mu0, bw_1..bw_D); the box follows SURVEY Appendix A:
scale in Var(Y)*[0.1, 10], noise in Var(Y)*[5e-3, 2e-1], mu0 in
median(Y) +- 3 std(Y), bw_j in std(X_j)*[1e-2, 1e2] (log-uniform where positive).
"""
import numpy as np

from .. import engine
from ..engine import DevicePosterior


def hp_bounds(X, Y):
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    y_var = float(np.var(Y))
    if not np.isfinite(y_var) or y_var <= 1e-12:
        y_var = 1.0
    y_std = np.sqrt(y_var)
    spread = X.std(axis=0)
    spread = np.where(spread > 1e-8, spread, 1.0)
    med = float(np.median(Y))
    lo = np.concatenate([[np.log(0.1 * y_var), np.log(5e-3 * y_var), med - 3.0 * y_std],
                         np.log(1e-2 * spread)])
    hi = np.concatenate([[np.log(10.0 * y_var), np.log(2e-1 * y_var), med + 3.0 * y_std],
                         np.log(1e2 * spread)])
    return lo, hi


def hp_candidates(X, Y, num, uniforms=None):
    """Row 0 = box centre, rows 1.. uniform in the transformed box; consumes one
    ``np.random.uniform(size=(num-1, P))`` from the global stream."""
    lo, hi = hp_bounds(X, Y)
    P = len(lo)
    if uniforms is None:
        uniforms = np.random.uniform(size=(max(num - 1, 0), P))
    z = np.vstack([np.full((1, P), 0.5), np.asarray(uniforms, dtype=np.float64).reshape(-1, P)])
    t = lo + z * (hi - lo)
    theta = t.copy()
    for col in [0, 1] + list(range(3, P)):
        theta[:, col] = np.exp(t[:, col])
    return theta


def batched_lml(X, Y, kind, thetas):
    """LML of every theta row in one device batch.  Returns (lml ndarray, post or None).  The
    data set is uploaded once and shared by all candidates (batch stride 0); a candidate whose
    Gram matrix needs the jitter ladder gets it (stable_cholesky), like any posterior.  At
    fixed padded shapes the pass is a replayed CUDA graph (engine.StepPlan): hyper-parameters
    are re-fitted after every evaluation (``tune_every 1`` in all option files), and for the
    small data sets of set2d / rand6d the pass is a dozen ~10 us kernels."""
    X = np.asarray(X, dtype=np.float64)
    n, D = X.shape
    thetas = np.asarray(thetas, dtype=np.float64).reshape(-1, 3 + D)
    B = len(thetas)

    def stage():
        Xd, yd, _, nv, nv_host = DevicePosterior.pack([X], [Y], thetas[:1], D)
        return Xd, yd, engine._h2d(thetas), nv, nv_host

    def posterior(st):
        Xd, yd, th, nv, nv_host = st
        return DevicePosterior.from_device(Xd.expand(B, -1, -1), yd.expand(B, -1), th, nv.repeat(B),
                                           nv_host * B, kind, D)

    if engine.graphs_enabled():
        n_pad = engine.pad_to(max(n, 1))
        plan = engine.StepPlan.get(('lml', kind, D, B, n_pad),
                                   8 * (n_pad * (D + 1) + 3 + D + thetas.size) + 16 * engine.Staging.ALIGN)
        with plan.staging():
            st = stage()

        def body(st_, ext):
            post = posterior(st_)
            pend = post.build_device_ladder(engine.GRAPH_SPEC_RUNGS)
            return [post.lml, pend.info]

        lml, info = plan.run(st, [], body)
        if not (info > 0).any():
            return np.array(lml), None
    with engine.staging(16 * X.size + 8 * thetas.size + (1 << 16)):
        st = stage()
    post = posterior(st)
    post.build(need_alpha=False)
    return post.lml.cpu().numpy(), post


def select(lmls, criterion, num_samples):
    finite = np.where(np.isfinite(lmls), lmls, -np.inf)
    if criterion == 'ml' or not num_samples or num_samples < 1:
        return [int(np.argmax(finite))]
    if criterion in ('post_sampling', 'post_mean'):
        w = np.exp(finite - np.max(finite))
        w = w / w.sum()
        u = np.random.uniform(size=int(num_samples))
        return [int(i) for i in np.minimum(np.searchsorted(np.cumsum(w), u), len(w) - 1)]
    raise ValueError('Unknown hp_tune_criterion %s' % criterion)
