"""Batched posterior-sample-and-reduce over many GPs / many candidate sets.

This is where the B200 design departs from the reference's Python ``for`` loops
(src/strategies/mts.py:28, src/cstrats/profile_cts.py:31): every task (or
context) becomes one problem of a strided batch and the whole
posterior -> covariance -> Cholesky -> sample -> max/argmax chain is ONE
sequence of batched kernel launches.  Normals are passed in (drawn by the
caller in the reference's order) so a seeded run consumes the global numpy
stream exactly like the reference.
"""
import os
from collections import defaultdict

import numpy as np
import torch

from . import engine, inflight
from .engine import DevicePosterior, F64


def _thetas(cores, D):
    """[scale, noise, mu0, bandwidths] of every core as one (B, 3 + D) array (filled in place:
    ``np.asarray([c.theta() ...])`` builds two temporaries per core, 3 us of a merged pass's
    ~6 us per problem)."""
    th = np.empty((len(cores), 3 + D))
    for k, c in enumerate(cores):
        hp = c.kernel.hyperparams
        row = th[k]
        row[0] = hp['scale']
        row[1] = c.noise_var
        row[2] = c._mu0()
        row[3:] = hp['dim_bandwidths']
    return th


def _stage_candidates(B, D, pts_list, Z_list, seg_lists):
    """Uploads of the sampling tail (call inside a staging block)."""
    Xd, mv_dev, mv, m_pad = engine.upload_points(B, D, pts_list)
    Z = engine.pad_normals(Z_list, mv, m_pad, 1)
    seg = engine._h2d(np.asarray(seg_lists, dtype=np.int32), engine.I32)
    return Xd, mv_dev, mv, m_pad, Z, seg


def _enqueue_tail(post, staged, spec=None):
    """posterior at pts -> Sigma -> stable_cholesky (plain attempt + speculative rungs)
    -> mu + L z -> segment max/argmax, all enqueued, no host round trip."""
    Xd, mv_dev, mv, m_pad, Z, seg = staged
    V = post.solve_cross(Xd, mv_dev)
    mean = post.mean_from_v(V, mv_dev)
    Sigma = post.cov(Xd, mv_dev, V, True)
    pend = engine.PendingCholesky(Sigma, mv_dev, spec)

    def tail():
        Fm = engine.sample_from_factor(mean, Sigma, Z, pend.info)
        return (Fm,) + engine.segment_argmax(Fm, 1, seg)

    return mean, pend, tail


def _sample_reduce(post, staged):
    """Eager form: the first read-back is the step's one synchronisation; problems that
    need more than the speculative rungs continue the ladder from the host.
    Returns (seg_max [B, nseg], seg_arg [B, nseg], mean dev, F dev, mv, jit)."""
    mean, pend, tail = _enqueue_tail(post, staged)
    Fm, mx, ag = tail()
    status = pend.status()
    if getattr(post, 'unchecked', False) and post.info.cpu().numpy().any():
        # a prior Gram matrix needed the ladder (rare): rebuild with the host-walked ladder, redo
        post.build()
        return _sample_reduce(post, staged)
    jit = pend.resolve(status)
    if (status[0] > 0).any():
        Fm, mx, ag = tail()
    return mx[:, :, 0].cpu().numpy(), ag[:, :, 0].cpu().numpy(), mean, Fm, staged[2], jit


def _rungs(rung_host):
    return [None if r == engine.NO_RUNG else int(r) for r in rung_host]


def _arena_bytes(B, n_pad, m_pad, D, nseg, with_posterior):
    doubles = B * m_pad * (D + 1)
    if with_posterior:
        doubles += B * n_pad * (D + 1) + B * (3 + D)
    return 8 * doubles + 4 * B * (nseg + 3) + 16 * engine.Staging.ALIGN


def _many(flat):
    """Lifts ``flat(list_a, list_b, ...) -> list`` (one entry per problem) to a merged pass
    over several requests (ocbo_b200.inflight): the requests' lists are concatenated, the
    device pass runs once, the results are split back."""
    def run_many(requests):
        sizes = [len(r[0]) for r in requests]
        joined = [[x for r in requests for x in r[k]] for k in range(len(requests[0]))]
        res = flat(*joined)
        out, at = [], 0
        for sz in sizes:
            out.append(res[at:at + sz])
            at += sz
        return out
    return run_many


def mts_step(gps, samp_pts_list, n_old_list, Z_list):
    """Independent-GP MTS reduction (src/strategies/mts.py:28-43) for all tasks
    at once.  ``gps[f]`` are B200GP objects (only their data and hyper-parameters
    are read: the batched posterior is rebuilt here, as ``get_next_gp`` would).
    Returns per task (max_prev, best_new_idx_relative, best_new_val).  With trials in
    flight (ocbo_b200.inflight) the tasks of all trials share the device pass."""
    return inflight.merge('mts', (gps, samp_pts_list, n_old_list, Z_list), _many(_mts_flat))


PIPELINE_MIN = int(os.environ.get('OCBO_PIPELINE_MIN', '64'))  # problems per chunk of a split pass
PIPELINE_MAX_CHUNKS = int(os.environ.get('OCBO_PIPELINE_CHUNKS', '3'))


def _chunks(members):
    """A merged pass of hundreds of problems is host-bound (staging ~6 us per problem, then the
    device runs while the host waits): split it, enqueue chunk k and stage chunk k + 1 meanwhile."""
    n = min(PIPELINE_MAX_CHUNKS, len(members) // PIPELINE_MIN) if engine.graphs_enabled() else 1
    if n < 2:
        return [members]
    step = (len(members) + n - 1) // n
    return [members[i:i + step] for i in range(0, len(members), step)]


def _mts_flat(gps, samp_pts_list, n_old_list, Z_list):
    T = len(gps)
    groups = defaultdict(list)
    for f, gp in enumerate(gps):
        core = gp.gp_core
        groups[(core.dim, core.kernel.kind)].append(f)
    out = [None] * T
    launched = []
    for (D, kind), all_members in groups.items():
        for c, members in enumerate(_chunks(all_members)):
            launched.append(_mts_launch(gps, samp_pts_list, n_old_list, Z_list, members, D, kind, c))
    for members, D, kind, stage, plan in launched:
        mx = None
        if plan is not None:
            mx_h, ag_h, info_k, info_s, rung = plan.results()
            if not info_k.any() and not (info_s > 0).any():
                mx, ag, jit = mx_h[:, :, 0], ag_h[:, :, 0], _rungs(rung)
        if mx is None:  # graphs off, or a problem needs the host-continued ladder
            with engine.staging():
                packed, staged = stage()
            post = DevicePosterior.from_device(*packed, kind, D)
            post.build(check=False, need_alpha=False)
            mx, ag, _, _, _, jit = _sample_reduce(post, staged)
        mx_l, ag_l = np.asarray(mx).tolist(), np.asarray(ag).tolist()  # (python scalars in one go)
        for k, f in enumerate(members):
            out[f] = (mx_l[k][0], ag_l[k][1], mx_l[k][1])
            gps[f].gp_core.last_sample_jitter = jit[k]
    return out


def _mts_launch(gps, samp_pts_list, n_old_list, Z_list, members, D, kind, chunk):
    """Stage one chunk and enqueue its step plan (no synchronisation)."""
    cores = [gps[f].gp_core for f in members]
    Xs_ = [c.X_array() for c in cores]
    ys_ = [c.Y_array() for c in cores]
    thetas = _thetas(cores, D)
    pts = [samp_pts_list[f] for f in members]
    segs = [[0, n_old_list[f], len(samp_pts_list[f])] for f in members]
    Zs = [Z_list[f] for f in members]

    def stage():
        return (DevicePosterior.pack(Xs_, ys_, thetas, D),
                _stage_candidates(len(members), D, pts, Zs, segs))

    if not engine.graphs_enabled():
        return members, D, kind, stage, None
    B = len(members)
    n_pad = engine.pad_to(max([len(y) for y in ys_] + [1]))
    m_pad = engine.pad_to(max([len(q) for q in pts] + [1]))
    # (the chunk index keeps the plans -- arenas, result buffers -- of equal-sized chunks apart)
    plan = engine.StepPlan.get(('mts', kind, D, B, n_pad, m_pad, chunk),
                               _arena_bytes(B, n_pad, m_pad, D, 2, True))
    with plan.staging():
        st_now = stage()

    def body(st, ext):
        packed, staged_ = st
        post_ = DevicePosterior.from_device(*packed, kind, D)
        post_.build(check=False, need_alpha=False)
        _, pend, tail = _enqueue_tail(post_, staged_, engine.GRAPH_SPEC_RUNGS)
        _, mx_, ag_ = tail()
        return [mx_, ag_, post_.info, pend.info, pend.rung]

    # chunk k > 0 replays on its own stream: the chunks' device work overlaps too
    plan.run(st_now, [], body, wait=False, stream=engine.chunk_stream(chunk) if chunk else None)
    return members, D, kind, stage, plan


def joint_many(requests):
    """Joint-MTS steps (B200GPCore.sample_and_pick) of one or several trials.  One request runs
    on the GP's own posterior (the tuned batch-1 path); several are grouped by kernel, input
    dimension and task count and go through ONE batched pass that rebuilds the trials'
    posteriors side by side, like ``_mts_flat``."""
    out = [None] * len(requests)
    groups = defaultdict(list)
    for i, (core, Xs, seg, T, Zh) in enumerate(requests):
        groups[(core.kernel.kind, core.dim, T, len(seg))].append(i)
    for (kind, D, T, _), members in groups.items():
        res = None
        if len(members) > 1:
            res = _joint_group([requests[i] for i in members], kind, D, T)
        for k, i in enumerate(members):
            core, Xs, seg, T_, Zh = requests[i]
            out[i] = res[k] if res is not None else core._sample_and_pick_alone(Xs, seg, T_, Zh)
    return out


def _joint_group(reqs, kind, D, T):
    """-> [(task, action, gain)] per request, or None when a problem needs the host-continued
    ladder (the callers then run alone, each on its own posterior)."""
    if not engine.graphs_enabled():
        return None
    cores = [r[0] for r in reqs]
    B = len(reqs)
    Xs_, ys_ = [c.X_array() for c in cores], [c.Y_array() for c in cores]
    thetas = _thetas(cores, D)
    pts, Zs = [r[1] for r in reqs], [r[4] for r in reqs]
    segs = np.asarray([r[2] for r in reqs], dtype=np.int32)
    n_pad = engine.pad_to(max([len(y) for y in ys_] + [1]))
    m_pad = engine.pad_to(max([len(q) for q in pts] + [1]))
    plan = engine.StepPlan.get(('joint-group', kind, D, B, n_pad, m_pad, segs.shape[1], T),
                               _arena_bytes(B, n_pad, m_pad, D, segs.shape[1], True))
    with plan.staging():
        st_now = (DevicePosterior.pack(Xs_, ys_, thetas, D), _stage_candidates(B, D, pts, Zs, segs))

    def body(st, ext):
        packed, staged_ = st
        post_ = DevicePosterior.from_device(*packed, kind, D)
        post_.build(check=False, need_alpha=False)
        _, pend, tail = _enqueue_tail(post_, staged_, engine.GRAPH_SPEC_RUNGS)
        _, mx_, ag_ = tail()
        return list(engine.joint_pick(mx_, ag_, T, T)) + [post_.info, pend.info, pend.rung]

    task, act, gain, info_k, info_s, rung = plan.run(st_now, [], body)
    if info_k.any() or (info_s > 0).any():
        return None
    jit = _rungs(rung)
    for k, c in enumerate(cores):
        c.last_sample_jitter = jit[k]
    return [(int(task[k, 0]), int(act[k, 0]), float(gain[k, 0])) for k in range(B)]


def mean_many(requests):
    """Posterior means (B200GPCore.eval, mean only: the risk-neutral pick of every BO step) of
    one or several trials' GPs at their own point sets.  One request runs on the GP's own
    posterior; several share one batched pass (posteriors rebuilt side by side)."""
    out = [None] * len(requests)
    groups = defaultdict(list)
    for i, (core, Xs) in enumerate(requests):
        groups[(core.kernel.kind, core.dim)].append(i)
    for (kind, D), members in groups.items():
        res = None
        if len(members) > 1:
            cores = [requests[i][0] for i in members]
            pts = [requests[i][1] for i in members]
            with engine.staging(16 * sum(c.X_array().size + len(c.Y) for c in cores)
                                + 16 * sum(p.size for p in pts) + (1 << 16)):
                post = DevicePosterior([c.X_array() for c in cores], [c.Y_array() for c in cores],
                                       _thetas(cores, D), kind, D)
                Xd, mv_dev, mv, _ = post.upload_points(pts)
            post.build(check=False, need_alpha=True)
            mean = post.mean(Xd, mv_dev)
            if not post.info.cpu().numpy().any():  # (a prior Gram matrix needing the ladder: alone)
                mean_h = mean.cpu().numpy()
                res = [mean_h[k, :mv[k]].copy() for k in range(len(members))]
        for k, i in enumerate(members):
            out[i] = res[k] if res is not None else requests[i][0]._mean_alone(requests[i][1])
    return out


class SharedPosterior(object):
    """View of ONE built posterior replicated (stride 0) over B candidate sets --
    the continuous strategies evaluate the same GP at 50 contexts
    (src/cstrats/profile_cts.py:26-37)."""

    def __init__(self, post, B):
        self.B, self.D, self.kind, self.n_pad = B, post.D, post.kind, post.n_pad
        ex = lambda t: t.expand(B, *t.shape[1:])
        # operands addressed through an explicit batch stride are shared (stride 0) ...
        self.X, self.theta, self.L, self.w = ex(post.X), ex(post.theta), ex(post.L), ex(post.w)
        # ... the per-problem arrays the kernels index as a[b] must be real
        self.nv = post.nv.repeat(B)
        self.info = torch.zeros((B,), dtype=engine.I32, device=post.L.device)
        self.invd = post.invd.repeat(B, 1, 1, 1)  # batch stride of invdiag is implicit in the ABI
        self.nv_host = list(post.nv_host) * B

    upload_points = DevicePosterior.upload_points
    mean_from_v = DevicePosterior.mean_from_v
    solve_cross = DevicePosterior.solve_cross
    cov = DevicePosterior.cov


class _PostView(object):
    """The operands of ONE built posterior that the sampling tail reads (optionally
    re-pointed at a StepPlan's own copies); usable as a batch-1 posterior itself."""
    FIELDS = ('X', 'theta', 'nv', 'L', 'invd', 'w')
    B = 1

    def __init__(self, post, tensors=None):
        self.D, self.kind, self.n_pad, self.nv_host = post.D, post.kind, post.n_pad, post.nv_host
        for name, t in zip(self.FIELDS, tensors or [getattr(post, f) for f in self.FIELDS]):
            setattr(self, name, t)
        self.info = post.info if tensors is None else None

    def with_info(self):
        self.info = torch.zeros((1,), dtype=engine.I32, device=self.L.device)
        return self

    def tensors(self):
        return [getattr(self, f) for f in self.FIELDS]

    mean_from_v = DevicePosterior.mean_from_v
    solve_cross = DevicePosterior.solve_cross
    cov = DevicePosterior.cov


def profile_step(gp, act_sets, Z_list):
    """Continuous MTS inner loop (src/cstrats/profile_cts.py:76-105) for all
    contexts at once.  Returns (sample_max, sample_argmax, mean_max, mean_argmax,
    sample_at_mean_argmax) as arrays over contexts.  With trials in flight the steps of the
    trials are enqueued one after the other, each on its own stream, and collected afterwards:
    trial k's device pass runs while trial k + 1 is staged and its posterior built."""
    return inflight.merge('cts', (gp, act_sets, Z_list), _profile_many)


def _profile_many(requests):
    if len(requests) == 1 or not engine.graphs_enabled():
        return [_profile_finish(_profile_launch(*r)) for r in requests]
    launched = [_profile_launch(*r, slot=k) for k, r in enumerate(requests)]
    return [_profile_finish(h) for h in launched]


def _reduce_means(mean, Fm, seg):
    mmx, mag = engine.segment_argmax(mean.unsqueeze(2), 1, seg)
    s_at = Fm[:, :, 0].gather(1, mag[:, 0, :].long())[:, 0]
    return mmx[:, 0, 0], mag[:, 0, 0], s_at


def _profile_launch(gp, act_sets, Z_list, slot=None):
    """Stage one trial's profile step and enqueue its plan (slot k: plan and stream number k;
    None: the current stream).  Nothing is read back here."""
    core = gp.gp_core
    if core._post_raw is None:
        core.build_posterior()
    base = core._post_raw  # possibly unverified: checked with the results, in _profile_finish
    B, D = len(act_sets), base.D
    segs = [[0, len(a)] for a in act_sets]
    plan = None
    if engine.graphs_enabled():
        m_pad = engine.pad_to(max([len(a) for a in act_sets] + [1]))
        plan = engine.StepPlan.get(('cts', base.kind, D, B, base.n_pad, m_pad, slot),
                                   _arena_bytes(B, 0, m_pad, D, 1, False))
        view = _PostView(base)
        with plan.staging():
            staged = _stage_candidates(B, D, act_sets, Z_list, segs)

        def body(st, ext):
            post = SharedPosterior(_PostView(base, ext), B)
            mean, pend, tail = _enqueue_tail(post, st, engine.GRAPH_SPEC_RUNGS)
            Fm, mx, ag = tail()
            return [mx, ag] + list(_reduce_means(mean, Fm, st[5])) + [pend.info, pend.rung]

        plan.run(staged, view.tensors(), body, wait=False,
                 stream=engine.chunk_stream(slot + 1) if slot is not None else None)
    return gp, act_sets, Z_list, base, segs, plan


def _profile_finish(handle):
    gp, act_sets, Z_list, base, segs, plan = handle
    core = gp.gp_core
    B, D = len(act_sets), base.D
    if plan is not None:
        mx, ag, mmx, mag, s_at, info_s, rung = plan.results()
        if base.unchecked and not core._verify(base):
            # K needed the ladder: redo on the rebuilt factor
            return _profile_finish(_profile_launch(gp, act_sets, Z_list))
        if not (info_s > 0).any():
            core.last_sample_jitter = _rungs(rung)
            return tuple(np.array(v) for v in (mx[:, 0, 0], ag[:, 0, 0], mmx, mag, s_at))
    post = SharedPosterior(base, B)
    with engine.staging():
        staged = _stage_candidates(B, D, act_sets, Z_list, segs)
    mx, ag, mean, Fm, mv, jit = _sample_reduce(post, staged)
    if base.unchecked and not core._verify(base):
        return _profile_finish(_profile_launch(gp, act_sets, Z_list))
    mmx, mag, s_at = [t.cpu().numpy() for t in _reduce_means(mean, Fm, staged[5])]
    core.last_sample_jitter = jit
    return (mx[:, 0], ag[:, 0], mmx, mag, s_at)


# ---------------------------------------------------------------------------
# EI family: posterior mean + diagonal variance + EI + arg-max,
# never forming the m x m covariance (ocbo_post_var_ei).
# ---------------------------------------------------------------------------
def _ei_reduce(post, pts_list, best):
    """best: device tensor [B] or a callable(mean_max[B] tensor) -> tensor [B]."""
    Xd, mv_dev, mv, m_pad = post.upload_points(pts_list)
    V = post.solve_cross(Xd, mv_dev)
    mean = post.mean_from_v(V, mv_dev)
    segs = np.asarray([[0, m] for m in mv], dtype=np.int32)
    if callable(best):
        mmx, _ = engine.segment_argmax(mean.unsqueeze(2), 1, segs)
        best = best(mmx[:, 0, 0])
    ei = torch.empty((post.B, m_pad), dtype=F64, device=mean.device)
    engine.call('ocbo_post_var_ei', engine._p(V), post.n_pad, V.stride(1), V.stride(0),
                engine._p(post.nv), engine._p(mean), mean.stride(0), engine._p(post.theta),
                post.theta.stride(0), engine._p(best), m_pad, engine._p(mv_dev), None, 0,
                engine._p(ei), ei.stride(0), post.B, engine._stream())
    emx, eag = engine.segment_argmax(ei.unsqueeze(2), 1, segs)
    return emx[:, 0, 0].cpu().numpy(), eag[:, 0, 0].cpu().numpy()


def mei_step(gps, pts_list, best_list):
    """Independent-GP MEI (src/strategies/mei.py:29-46 via misc_util.py:368-385) for all
    tasks at once.  Returns per task (max EI, arg-max index)."""
    return inflight.merge('mei', (gps, pts_list, best_list), _many(_mei_flat))


def _mei_flat(gps, pts_list, best_list):
    T = len(gps)
    groups = defaultdict(list)
    for f, gp in enumerate(gps):
        groups[(gp.gp_core.dim, gp.gp_core.kernel.kind)].append(f)
    out = [None] * T
    for (D, kind), members in groups.items():
        cores = [gps[f].gp_core for f in members]
        post = DevicePosterior([c.X_array() for c in cores], [c.Y_array() for c in cores],
                               _thetas(cores, D), kind, D).build()
        best = engine._h2d(np.asarray([best_list[f] for f in members], dtype=np.float64))
        pts = [np.asarray(pts_list[f], dtype=np.float64).reshape(-1, D) for f in members]
        emx, eag = _ei_reduce(post, pts, best)
        for k, f in enumerate(members):
            out[f] = (float(emx[k]), int(eag[k]))
    return out


def shared_ei_step(gp, pts_list, best_list=None, cap=None):
    """EI over several candidate sets of ONE GP (joint-MEI: src/strategies/joint_mei.py:
    26-52; PEI: src/cstrats/profile_cts.py:54-67).  ``best_list`` gives the incumbent per
    set; with ``cap`` (PEI) the incumbent is min(max posterior mean of the set, cap)."""
    core = gp.gp_core
    if core._post is None:
        core.build_posterior()
    post = SharedPosterior(core._post, len(pts_list))
    if cap is not None:
        best = lambda mean_max: torch.clamp(mean_max, max=float(cap)).contiguous()
    else:
        best = engine._h2d(np.asarray(best_list, dtype=np.float64))
    return _ei_reduce(post, pts_list, best)


# ---------------------------------------------------------------------------
# REVI: knowledge gradient of every candidate over the judgement groups.
# ---------------------------------------------------------------------------
def revi_improvements(gp, cand_pts, judge_pts, num_groups, group_size):
    """improvement[c] = sum_g KG(judge_means[g], cov(c, judge[g]) / sqrt(noise + var_c))
    (src/strategies/corr_opt.py:92-119, src/cstrats/postmax_cts.py:139-160) for all C
    candidates in one device pass: V for both point sets, posterior means of the judgement
    points, the C x (G E) posterior cross-covariance (get_pt_relations, :112-119), the
    candidates' variances, then ocbo_kg on C x G CTAs.  The sum over groups runs in group
    order on the host (C x G doubles), like the reference's ``improvement +=``."""
    core = gp.gp_core
    if not core.has_posterior():
        core.build_posterior()
    post = core._post  # verified factor
    D = post.D
    cand = np.asarray(cand_pts, dtype=np.float64).reshape(-1, D)
    judge = np.asarray(judge_pts, dtype=np.float64).reshape(-1, D)
    C, G, E = len(cand), int(num_groups), int(group_size)
    if len(judge) != G * E:
        raise ValueError('judgement set must hold num_groups * group_size points')
    with engine.staging(16 * (cand.size + judge.size) + (1 << 16)):
        Xa, mav, _, _ = engine.upload_points(1, D, [cand])
        Xb, mbv, _, _ = engine.upload_points(1, D, [judge])
    Va, Vb = post.solve_cross(Xa, mav), post.solve_cross(Xb, mbv)
    means_j = post.mean_from_v(Vb, mbv)
    var_c = torch.empty((1, Xa.shape[1]), dtype=F64, device=Xa.device)
    engine.call('ocbo_post_var_ei', engine._p(Va), post.n_pad, Va.stride(1), Va.stride(0),
                engine._p(post.nv), None, 0, engine._p(post.theta), post.theta.stride(0), None,
                Xa.shape[1], engine._p(mav), engine._p(var_c), var_c.stride(0), None, 0, 1,
                engine._stream())
    inter = post.cross_cov(Xa, mav, Va, Xb, mbv, Vb)
    kg = engine.knowledge_gradients(means_j[0], E, inter[0], var_c[0], gp.get_estimated_noise(),
                                    C, G, E)
    return np.cumsum(kg.cpu().numpy(), axis=1)[:, -1]
