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
from collections import defaultdict

import numpy as np

from .. import engine, inflight
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
    med = _median(Y)
    P = 3 + X.shape[1]
    lo, hi = np.empty(P), np.empty(P)
    lo[0], lo[1], lo[2] = np.log(0.1 * y_var), np.log(5e-3 * y_var), med - 3.0 * y_std
    hi[0], hi[1], hi[2] = np.log(10.0 * y_var), np.log(2e-1 * y_var), med + 3.0 * y_std
    lo[3:] = np.log(1e-2 * spread)
    hi[3:] = np.log(1e2 * spread)
    return lo, hi


def _median(y):
    """np.median's value (middle element, or the mean of the two middle ones) without its
    ~40 us of Python dispatch -- a fit per BO step calls this."""
    s = np.sort(y, axis=None)
    n = s.size
    if n == 0:
        return float('nan')
    return float(s[n // 2]) if n % 2 else float(0.5 * (s[n // 2 - 1] + s[n // 2]))


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


def _to_theta(t):
    theta = np.array(t, dtype=np.float64, copy=True)
    P = theta.shape[1]
    for col in [0, 1] + list(range(3, P)):
        theta[:, col] = np.exp(theta[:, col])
    return theta


def _to_box(theta):
    t = np.array(theta, dtype=np.float64, copy=True).reshape(1, -1)
    for col in [0, 1] + list(range(3, t.shape[1])):
        t[:, col] = np.log(t[:, col])
    return t[0]


def refine_candidates(X, Y, num, rnd, crit, best_theta, uniforms=None):
    """Candidates of search round ``rnd`` >= 1 (``hp_search_rounds`` > 1: the device is idle during
    a 64-candidate fit -- 466 k LML evaluations/s at n = 55 -- so a fit can afford more of them).
      'ml'            zoom: ``num - 1`` uniform draws in a box of half-width (hi - lo) / 2^(rnd+1)
                      around the best candidate so far (transformed coordinates, clipped to the
                      search box);
      otherwise       ``num - 1`` more uniform draws in the whole box (importance resampling needs
                      draws from the prior box, not from a zoomed proposal).
    Consumes one ``np.random.uniform(size=(num-1, P))`` from the global stream."""
    lo, hi = hp_bounds(X, Y)
    P = len(lo)
    if uniforms is None:
        uniforms = np.random.uniform(size=(max(num - 1, 0), P))
    u = np.asarray(uniforms, dtype=np.float64).reshape(-1, P)
    if crit == 'ml':
        centre = _to_box(best_theta)
        half = (hi - lo) * 0.5 ** (rnd + 1)
        blo, bhi = np.maximum(lo, centre - half), np.minimum(hi, centre + half)
    else:
        blo, bhi = lo, hi
    return _to_theta(blo + u * (bhi - blo))


def search(X, Y, kind, num, rounds, crit, lml_fn=None):
    """All rounds of one fit.  Returns (candidates [rounds*num - (rounds-1), P], lmls)."""
    lml_fn = lml_fn or (lambda th: batched_lml(X, Y, kind, th)[0])
    cands = hp_candidates(X, Y, num)
    lmls = np.asarray(lml_fn(cands), dtype=np.float64)
    for rnd in range(1, max(int(rounds), 1)):
        finite = np.where(np.isfinite(lmls), lmls, -np.inf)
        more = refine_candidates(X, Y, num, rnd, crit, cands[int(np.argmax(finite))])
        cands = np.concatenate([cands, more], axis=0)
        lmls = np.concatenate([lmls, np.asarray(lml_fn(more), dtype=np.float64)])
    return cands, lmls


def batched_lml(X, Y, kind, thetas):
    """LML of every theta row in one device batch.  Returns (lml ndarray, post or None).  The
    data set is uploaded once and shared by all candidates (batch stride 0); a candidate whose
    Gram matrix needs the jitter ladder gets it (stable_cholesky), like any posterior.  At
    fixed padded shapes the pass is a replayed CUDA graph (engine.StepPlan): hyper-parameters
    are re-fitted after every evaluation (``tune_every 1`` in all option files), and for the
    small data sets of set2d / rand6d the pass is a dozen ~10 us kernels.  With trials in
    flight (ocbo_b200.inflight) the fits of all trials share the pass (``_lml_group``)."""
    X = np.asarray(X, dtype=np.float64)
    thetas = np.asarray(thetas, dtype=np.float64).reshape(-1, 3 + X.shape[1])
    return inflight.merge('lml', (X, np.asarray(Y, dtype=np.float64).ravel(), int(kind), thetas),
                          _lml_many)


def _lml_many(jobs):
    """Merged LML pass: jobs (X, Y, kind, thetas) with the same kernel kind, input dimension
    and candidate count form one batch of sum(B_j) problems."""
    out = [None] * len(jobs)
    groups = defaultdict(list)
    for j, (X, Y, kind, thetas) in enumerate(jobs):
        groups[(kind, X.shape[1], len(thetas))].append(j)
    for (kind, D, B), members in groups.items():
        if len(members) == 1:
            out[members[0]] = _lml_single(*jobs[members[0]])
            continue
        # factor + pristine copy + workspace: ~3 n_pad^2 doubles per problem in HBM
        n_pad = engine.pad_to(max(len(jobs[j][1]) for j in members))
        chunk = max(1, int(MERGED_LML_BYTES // (24 * B * n_pad * n_pad)))
        for c0 in range(0, len(members), chunk):
            part = members[c0:c0 + chunk]
            if len(part) == 1:
                out[part[0]] = _lml_single(*jobs[part[0]])
                continue
            for j, lml in zip(part, _lml_group([jobs[j] for j in part], kind, D, B)):
                out[j] = (lml, None)
    return out


MERGED_LML_BYTES = 32 << 30  # HBM budget of one merged LML pass (180 GB per B200)


def _lml_group(jobs, kind, D, B):
    """J data sets x B candidates each as ONE batch of J*B posteriors.  The data sets are
    uploaded once each; the device replicates them per candidate (the C ABI addresses a batch
    through ONE stride, so "stride 0 inside groups of B" is a copy -- J*B*n_pad*(D+1) doubles,
    device to device)."""
    J = len(jobs)
    Xs, Ys = [j[0] for j in jobs], [j[1] for j in jobs]
    thetas = np.concatenate([j[3] for j in jobs], axis=0)

    def stage():
        Xd, yd, _, nv, nv_host = DevicePosterior.pack(Xs, Ys, thetas[:J], D)
        return Xd, yd, engine._h2d(thetas), nv, nv_host

    def posterior(st):
        Xd, yd, th, nv, nv_host = st
        return DevicePosterior.from_device(Xd.repeat_interleave(B, dim=0), yd.repeat_interleave(B, dim=0),
                                           th, nv.repeat_interleave(B),
                                           [n for n in nv_host for _ in range(B)], kind, D)

    n_pad = engine.pad_to(max([len(y) for y in Ys] + [1]))
    if engine.graphs_enabled():
        plan = engine.StepPlan.get(('lml-group', kind, D, J, B, n_pad),
                                   8 * (J * (n_pad * (D + 1) + 3 + D) + thetas.size) + 4 * J
                                   + 16 * engine.Staging.ALIGN)
        with plan.staging():
            st = stage()

        def body(st_, ext):
            post = posterior(st_)
            pend = post.build_device_ladder(engine.GRAPH_SPEC_RUNGS)
            return [post.lml, pend.info]

        lml, info = plan.run(st, [], body)
        if not (info > 0).any():
            return list(np.array(lml).reshape(J, B))
    with engine.staging(16 * sum(x.size for x in Xs) + 8 * thetas.size + 16 * J * n_pad + (1 << 16)):
        st = stage()
    post = posterior(st)
    post.build(need_alpha=False)
    return list(post.lml.cpu().numpy().reshape(J, B))


def _lml_single(X, Y, kind, thetas):
    n, D = X.shape
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
    if not np.isfinite(finite).any():
        # every candidate's Gram matrix failed to factor / gave a non-finite likelihood: the
        # arg-max (or NaN importance weights) would silently hand back candidate 0
        raise ValueError('hyper-parameter search: no candidate has a finite log marginal likelihood')
    if criterion == 'ml' or not num_samples or num_samples < 1:
        return [int(np.argmax(finite))]
    if criterion in ('post_sampling', 'post_mean'):
        w = np.exp(finite - np.max(finite))
        w = w / w.sum()
        u = np.random.uniform(size=int(num_samples))
        return [int(i) for i in np.minimum(np.searchsorted(np.cumsum(w), u), len(w) - 1)]
    raise ValueError('Unknown hp_tune_criterion %s' % criterion)
