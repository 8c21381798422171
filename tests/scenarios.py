"""Seeded end-to-end MTS scenarios shared by the golden-vector generator (run
with the reference's own strategy classes on the numpy oracle, in the build
container) and the GPU parity test (run with ocbo_b200's strategies on the B200
engine).  Shapes follow the four benchmark option files, shortened so the CPU
side finishes in seconds."""
from argparse import Namespace

import numpy as np

SEED = 100826730  # src/ocbo.py:74-75

DISCRETE = {
    # set2d.txt: MTS, independent GPs, 2-D tasks, tuning ml
    'set2d': dict(kind='mts', functions='branin,parabaloid,parabaloid', capital=10, init_capital=5,
                  tuning_methods='ml', hp_samples=3, risk_neutral=False),
    # rand6d.txt: MTS, 6-D random mass tasks, ml-post_sampling with hp_samples draws
    'rand6d': dict(kind='mts', masses=(1, 1, 2), massf_dim=6, capital=10, init_capital=5,
                   tuning_methods='ml-post_sampling', hp_samples=3, risk_neutral=False),
    # jointbran.txt: joint-MTS over chopped Branin, risk-neutral bookkeeping on
    'jointbran': dict(kind='joint-mts', cts_f='branin', num_chops=4, capital=8, init_capital=5,
                      tuning_methods='ml', hp_samples=3, risk_neutral=True),
    # EI family: the other GP-driven methods of the same option files
    'set2d-mei': dict(kind='mei', functions='branin,parabaloid,parabaloid', capital=8, init_capital=5,
                      tuning_methods='ml', hp_samples=3, risk_neutral=False),
    'jointbran-mei': dict(kind='joint-mei', cts_f='branin', num_chops=4, capital=8, init_capital=5,
                          tuning_methods='ml', hp_samples=3, risk_neutral=True),
}
# Baselines of the same option files (agnostic_opt.py, joint_agnostic_opt.py, random_strat.py,
# rand_cts.py): host-logic parity only (tests/test_host_strategies.py) -- same loop, same
# consumption of the global numpy stream, same picks on the same oracle engine.
DISCRETE_BASELINES = {
    'set2d-agn-thomp': dict(kind='agn-thomp', functions='branin,parabaloid,parabaloid', capital=8,
                            init_capital=5, tuning_methods='ml', hp_samples=3, risk_neutral=False),
    'set2d-agn-ei': dict(kind='agn-ei', functions='branin,parabaloid,parabaloid', capital=8,
                         init_capital=5, tuning_methods='ml', hp_samples=3, risk_neutral=False),
    'set2d-random': dict(kind='Random', functions='branin,parabaloid,parabaloid', capital=8,
                         init_capital=5, tuning_methods='ml', hp_samples=3, risk_neutral=False),
    'jointbran-ja-thomp': dict(kind='ja-thomp', cts_f='branin', num_chops=4, capital=6, init_capital=5,
                               tuning_methods='ml', hp_samples=3, risk_neutral=True),
    'jointbran-ja-ei': dict(kind='ja-ei', cts_f='branin', num_chops=4, capital=6, init_capital=5,
                            tuning_methods='ml', hp_samples=3, risk_neutral=True),
    # REVI (corr_opt.py): knowledge gradient over the joint GP; EI round robin after run_ei_after
    'jointbran-revi': dict(kind='revi', cts_f='branin', num_chops=3, capital=5, init_capital=5,
                           tuning_methods='ml', hp_samples=3, risk_neutral=True, num_eval_pts=30,
                           num_candidates=12, run_ei_after=4),
    'jointbran-rand': dict(kind='joint-rand', cts_f='branin', num_chops=4, capital=6, init_capital=5,
                           tuning_methods='ml', hp_samples=3, risk_neutral=True),
}
CONTINUOUS_BASELINES = {
    'conth42-rand': dict(kind='rand', act_dim=2, capital=5, init_capital=5, num_profiles=10,
                         profile_evals=100, eval_ctx_fidel=8, eval_act_fidel=8),
    'conth42-agn-ts': dict(kind='ts', act_dim=2, capital=5, init_capital=5, num_profiles=10,
                           profile_evals=100, eval_ctx_fidel=8, eval_act_fidel=8),
    'conth42-revi': dict(kind='revi', act_dim=2, capital=4, init_capital=40, num_profiles=10,
                         profile_evals=100, eval_ctx_fidel=8, eval_act_fidel=8, cand_size=15,
                         judge_ctx_size=6, judge_act_size=20),
    'conth42-agn-ei': dict(kind='ei', act_dim=2, capital=5, init_capital=5, num_profiles=10,
                           profile_evals=100, eval_ctx_fidel=8, eval_act_fidel=8),
}
CONTINUOUS = {
    # conth42.txt: Hartmann-6 with 4 context + 2 action dims
    'conth42-cmts': dict(kind='cmts', act_dim=2, capital=6, init_capital=5, num_profiles=10,
                         profile_evals=100, eval_ctx_fidel=8, eval_act_fidel=8),
    'conth42-cmts-pm': dict(kind='cmts-pm', act_dim=2, capital=6, init_capital=5, num_profiles=10,
                            profile_evals=100, eval_ctx_fidel=8, eval_act_fidel=8),
    'conth42-pei': dict(kind='pei', act_dim=2, capital=6, init_capital=5, num_profiles=10,
                        profile_evals=100, eval_ctx_fidel=8, eval_act_fidel=8),
}


def discrete_options(cfg, gp_engine):
    return Namespace(tuning_methods=cfg['tuning_methods'], hp_samples=cfg['hp_samples'],
                     kernel_type='se', risk_neutral=cfg['risk_neutral'], rn_eval_size=50,
                     gp_engine=gp_engine, tune_every=1, max_opt_evals=100,
                     num_eval_pts=cfg.get('num_eval_pts', 100), num_candidates=cfg.get('num_candidates', 100),
                     run_ei_after=cfg.get('run_ei_after', float('inf')))


def continuous_options(cfg, gp_engine):
    return Namespace(tuning_methods='ml', tune_every=1, kernel_type='se', use_additive_gp=False,
                     add_max_group_size=6, hp_samples=3, eval_ctx_fidel=cfg['eval_ctx_fidel'],
                     eval_act_fidel=cfg['eval_act_fidel'], score_every=1, cand_size=cfg.get('cand_size', 100),
                     judge_ctx_size=cfg.get('judge_ctx_size', 50), judge_act_size=cfg.get('judge_act_size', 50),
                     judge_ctx_thresh=None, judge_act_thresh=None,
                     gp_engine=gp_engine, num_profiles=cfg['num_profiles'],
                     profile_evals=cfg['profile_evals'], agn_evals=100)


def run_discrete(cfg, strategy_cls, f_infos, options):
    """src/ocbo.py:128-144 for one method: shared init points, then optimize."""
    method = strategy_cls(f_infos, options)
    init_pts = []
    for wrap in f_infos:
        low, high = zip(*wrap.domain)
        init_pts.append([np.random.uniform(low, high) for _ in range(cfg['init_capital'])])
    method.set_clean()
    h = method.optimize(cfg['capital'], init_pts=init_pts)
    names = [fi.name for fi in f_infos]
    pts, vals = [], []
    # re-assemble the queries in time order from the per-task histories
    counters = {n: cfg['init_capital'] for n in names}
    for f_idx in h.choice_timeline:
        q = h.query_history[names[f_idx]][counters[names[f_idx]]]
        counters[names[f_idx]] += 1
        pts.append([float(v) for v in np.ravel(q.pt)])
        vals.append(float(q.val))
    return dict(choice_timeline=[int(c) for c in h.choice_timeline], points=pts, values=vals,
                total_best=[float(v) for v in h.total_best])


def run_continuous(cfg, strategy_cls, function, domain, options):
    """src/cts_ocbo.py:95-110 for one method."""
    ctx_dim = len(domain) - cfg['act_dim']
    method = strategy_cls(function, domain, ctx_dim, options)
    init_pts = list(np.random.uniform(*zip(*domain), (cfg['init_capital'], len(domain))))
    method.set_clean()
    h = method.optimize(cfg['capital'], init_pts=init_pts)
    return dict(points=[[float(v) for v in np.ravel(q.pt)] for q in h.query_history],
                rewards=[float(q.reward) for q in h.query_history],
                y_init=[float(v) for v in h.y_init],
                regret=[float(s.regret) for s in h.score_history])


# ---------------------------------------------------------------------------
# "Documented ties" (BASELINE.json north_star: picks bit-exact APART FROM
# DOCUMENTED TIES).  MTS samples at the queried points themselves, so the
# posterior covariance is numerically singular and Dragonfly's jitter ladder is
# exercised routinely.  With jitter j = 10^p * max(diag), a
# joint sample is only defined up to O(sqrt(j) * |z|): two correct fp64
# implementations (different summation orders, different rung of the ladder on
# knife-edge matrices) differ by that much.  Worse, on such matrices a tiny
# lucky pivot amplifies rounding noise (column = residual / sqrt(pivot)), and
# the reference's OWN sample moves by `rung_sens` when its jitter moves one rung
# (measured on the B200 box: up to 3.6 on jointbran, function range ~100; the
# GPU sample agrees with numpy's Cholesky of the same matrix at the same rung
# to 1e-7).  A pick whose margin over the runner-up is below
#     max( TIE_C * sqrt(10^p_eff * max_diag) * |z|_max ,  rung_sens )
# is a documented tie.
# ---------------------------------------------------------------------------
TIE_C = 10.0
P_NONE = -12  # effective rung when no jitter was needed


def _tol(draws):
    t = 0.0
    for d in draws:
        p = P_NONE if d['jitter_p'] is None else d['jitter_p']
        t = max(t, TIE_C * np.sqrt(10.0 ** p * max(d['max_diag'], 1e-300)) * d['z_max'],
                d.get('rung_sens', 0.0))
    return float(t)


def _top2_gap(v):
    v = np.sort(np.ravel(v))[::-1]
    return float(v[0] - v[1]) if len(v) > 1 else float('inf')


def discrete_margins(cfg, num_tasks, timeline, trace, A=100):
    """Per step: (margin of the reference's pick over the runner-up, tie tolerance)."""
    counts = [cfg['init_capital']] * num_tasks
    out = []
    if cfg['kind'] in ('mei', 'joint-mei'):
        return [(float('inf'), 0.0)] * len(timeline)  # no sampling: exact match required
    per_step = 1 if cfg['kind'] == 'joint-mts' else num_tasks
    for t, chosen in enumerate(timeline):
        draws = trace[t * per_step:(t + 1) * per_step]
        if cfg['kind'] == 'joint-mts':
            s = draws[0]['sample']
            tq = sum(counts)
            divs = np.concatenate([[0], np.cumsum(counts)])
            old = np.asarray([np.max(s[divs[i]:divs[i + 1]]) for i in range(num_tasks)])[:, None]
            margin = _top2_gap(s[tq:].reshape(num_tasks, -1) - old)
        else:
            gains, within = [], []
            for f, d in enumerate(draws):
                s = d['sample']
                gains.append(np.max(s[counts[f]:]) - np.max(s[:counts[f]]))
                within.append(_top2_gap(s[counts[f]:]))
            margin = min(_top2_gap(gains), within[int(np.argmax(gains))])
        out.append((margin, _tol(draws)))
        counts[chosen] += 1
    return out


def continuous_margins(cfg, trace, y_init, rewards):
    P = cfg['num_profiles']
    out = []
    if cfg['kind'] == 'pei':
        return pei_margins(cfg, trace, y_init, rewards)
    y = list(y_init)
    for t in range(cfg['capital']):
        draws = trace[t * P:(t + 1) * P]
        gains, within = [], []
        for d in draws:
            s, mu = d['sample'], d['mean']
            if cfg['kind'] == 'cmts':
                gains.append(np.max(s) - s[int(np.argmax(mu))])
            else:
                gains.append(np.max(s) - min(np.max(mu), np.max(y)))
            within.append(_top2_gap(s))
        out.append((min(_top2_gap(gains), within[int(np.argmax(gains))]), _tol(draws)))
        y.append(rewards[t])
    return out


EI_TIE_REL = 1e-6  # EI arg-max ties: relative gap of the two largest EI values


def pei_margins(cfg, evals, y_init, rewards):
    """PEI (profile_cts.py:54-67) draws no samples; its pick is an arg-max over EI
    values that underflow towards 0 when no candidate is likely to improve.  Margin =
    relative gap between the two largest EI values (across contexts and inside the
    winning context); below EI_TIE_REL the two candidates are tied in fp64.  The
    reference evaluates the score GP once more per step (cts_opt.py:312) -- those
    evals are 'none'-form and not traced."""
    from scipy.stats import norm
    P = cfg['num_profiles']
    out, y = [], list(y_init)
    for t in range(cfg['capital']):
        best_vals, within = [], []
        for d in evals[t * P:(t + 1) * P]:
            mu, sd = d['mean'], np.sqrt(d['var'])
            z = (mu - min(np.max(mu), np.max(y))) / sd
            ei = sd * (z * norm.cdf(z) + norm.pdf(z))
            best_vals.append(np.max(ei))
            within.append(_top2_gap(ei))
        top = max(np.max(best_vals), 1e-300)
        margin = min(_top2_gap(best_vals), within[int(np.argmax(best_vals))]) / top
        out.append((float(margin), EI_TIE_REL))
        y.append(rewards[t])
    return out


def first_divergence(got_pts, ref_pts, atol=1e-9):
    for k, (a, b) in enumerate(zip(got_pts, ref_pts)):
        if np.max(np.abs(np.asarray(a) - np.asarray(b))) > atol:
            return k
    return None


def assert_same_up_to_ties(got_pts, ref, atol=1e-9):
    """Selections must be identical; the only excuse for the first divergence is a
    documented tie at that step (after which the two runs are different runs)."""
    k = first_divergence(got_pts, ref['points'], atol)
    if k is None:
        return None
    margin, tol = ref['margins'][k]
    assert margin <= tol, ('selection differs at step %d but the reference pick led by %.3e '
                           '> tie tolerance %.3e' % (k, margin, tol))
    return k


# ---------------------------------------------------------------------------
# Option-file-length traces (src/options/{set2d,jointbran,rand6d,conth42}.txt at their real
# task counts and capitals) with the comparison CONTINUING past an accepted tie.
#
# The golden side is the reference's own strategy class, run freely; per step it stores the
# pick, the tie margin / tolerance, the jitter rung of every draw and a digest of every GP's
# hyper-parameters at decision time.  The checked side runs the same loop "teacher forced":
# its own decision is recorded and compared, then the REFERENCE's query is the one evaluated
# and added, so both runs hold the same data (and the same position in the global numpy
# stream) at every step -- one accepted tie no longer ends the comparison.
# ---------------------------------------------------------------------------
LONG_DISCRETE = {
    # set2d.txt:3-8: T=5, capital 100
    'set2d-full': dict(kind='mts', functions='branin,parabaloid,parabaloid,parabaloid,parabaloid',
                       capital=100, init_capital=5, tuning_methods='ml', hp_samples=3,
                       risk_neutral=False),
    # jointbran.txt:3-12: T=10 chops of Branin, capital 110, risk-neutral bookkeeping
    'jointbran-full': dict(kind='joint-mts', cts_f='branin', num_chops=10, capital=110, init_capital=5,
                           tuning_methods='ml', hp_samples=3, risk_neutral=True),
    # rand6d.txt:3-13: T=30 (15 easy / 10 medium / 5 hard), 6-D, ml-post_sampling, 5 HP draws;
    # 300 of the file's 1500 steps
    'rand6d-full': dict(kind='mts', masses=(15, 10, 5), massf_dim=6, capital=300, init_capital=5,
                        tuning_methods='ml-post_sampling', hp_samples=5, risk_neutral=False),
}
LONG_CONTINUOUS = {
    # conth42.txt:3-9: Hartmann-6 as 4 context + 2 action dims, 50 contexts x 100 actions per step
    # (profile_cts.py:13-18); 200 of the file's 750 steps; the policy score on an 8 x 8 grid
    'conth42-full-cmts': dict(kind='cmts', act_dim=2, capital=200, init_capital=5, num_profiles=50,
                              profile_evals=100, eval_ctx_fidel=8, eval_act_fidel=8),
    'conth42-full-cmts-pm': dict(kind='cmts-pm', act_dim=2, capital=200, init_capital=5, num_profiles=50,
                                 profile_evals=100, eval_ctx_fidel=8, eval_act_fidel=8),
}


def hp_digest(gp):
    """One number per GP: log scale + sum log length-scales + log noise (GPWrapper interface,
    src/gp/gp_interface.py:32-38) -- enough to see a hyper-parameter fit go a different way."""
    hps = gp.get_kernel_hps()
    ls = np.ravel(np.asarray(hps['lengthscale'], dtype=np.float64))
    return float(np.log(float(hps['scale'])) + np.sum(np.log(ls)) + np.log(gp.get_estimated_noise()))


def _discrete_gps(method):
    gps = list(method.gps)
    return gps[:1] if all(g is gps[0] for g in gps) else gps   # JointOpt aliases one GP T times


class Recorder(object):
    """Wraps a strategy object's decision method.  ``forced`` = the reference's per-step picks
    (teacher forcing) or None (free run: golden generation).  ``rungs(method)`` returns the
    jitter rungs of the step's draws if the engine exposes them (else None)."""

    def __init__(self, method, attr, forced=None, rungs=None, gps=_discrete_gps):
        self.method, self.attr, self.forced, self.rungs_fn, self.gps_fn = method, attr, forced, rungs, gps
        self.picks, self.rungs, self.hps = [], [], []
        self.inner = getattr(method, attr)
        setattr(method, attr, self)

    def __call__(self, *args, **kwargs):
        self.hps.append([hp_digest(g) for g in self.gps_fn(self.method)])
        out = self.inner(*args, **kwargs)
        self.picks.append(out)
        if self.rungs_fn is not None:
            self.rungs.append(self.rungs_fn(self.method))
        if self.forced is not None:
            return self.forced[len(self.picks) - 1]
        return out


def run_discrete_long(cfg, strategy_cls, f_infos, options, ref=None, rungs=None):
    """src/ocbo.py:128-144 for one method; with ``ref`` (a golden entry) the run is teacher
    forced.  Returns the per-step record."""
    method = strategy_cls(f_infos, options)
    init_pts = []
    for wrap in f_infos:
        low, high = zip(*wrap.domain)
        init_pts.append([np.random.uniform(low, high) for _ in range(cfg['init_capital'])])
    method.set_clean()
    forced = None
    if ref is not None:
        forced = [(int(c), np.asarray(p, dtype=np.float64)) for c, p in zip(ref['choices'], ref['points'])]
    rec = Recorder(method, 'decide_next_query', forced, rungs)
    h = method.optimize(cfg['capital'], init_pts=init_pts)
    return dict(choices=[int(p[0]) for p in rec.picks],
                points=[[float(v) for v in np.ravel(p[1])] for p in rec.picks],
                hps=rec.hps, rungs=rec.rungs, total_best=[float(v) for v in h.total_best])


def run_continuous_long(cfg, strategy_cls, function, domain, options, ref=None, rungs=None):
    """src/cts_ocbo.py:95-110 for one method, optionally teacher forced."""
    ctx_dim = len(domain) - cfg['act_dim']
    method = strategy_cls(function, domain, ctx_dim, options)
    init_pts = list(np.random.uniform(*zip(*domain), (cfg['init_capital'], len(domain))))
    method.set_clean()
    forced = [np.asarray(p, dtype=np.float64) for p in ref['points']] if ref is not None else None
    rec = Recorder(method, '_determine_next_query', forced, rungs, gps=lambda m: [m.gp])
    h = method.optimize(cfg['capital'], init_pts=init_pts)
    return dict(points=[[float(v) for v in np.ravel(p)] for p in rec.picks], hps=rec.hps,
                rungs=rec.rungs, rewards=[float(q.reward) for q in h.query_history],
                y_init=[float(v) for v in h.y_init],
                regret=[float(s.regret) for s in h.score_history])


def compare_long(got, ref, atol=1e-9, hp_atol=1e-6):
    """Per-step comparison of a teacher-forced run with its golden trace.  Every differing
    pick must be a documented tie (margin <= tolerance).  Returns the statistics DESIGN.md
    section 2 quotes."""
    steps = len(ref['points'])
    assert len(got['points']) == steps
    diff_pick, diff_rung, rung_steps, hp_diff, worst, bad = [], 0, 0, 0, 0.0, []
    draws, draw_diff, hist = 0, 0, {}
    for k in range(steps):
        if np.max(np.abs(np.asarray(got['hps'][k]) - np.asarray(ref['hps'][k]))) > hp_atol:
            hp_diff += 1
        if got['rungs'] and ref.get('rungs'):
            rung_steps += 1
            if list(got['rungs'][k]) != list(ref['rungs'][k]):
                diff_rung += 1
            for a, b in zip(got['rungs'][k], ref['rungs'][k]):
                d = (P_NONE if a is None else a) - (P_NONE if b is None else b)
                draws += 1
                draw_diff += d != 0
                hist[str(d)] = hist.get(str(d), 0) + 1
        same = np.max(np.abs(np.asarray(got['points'][k]) - np.asarray(ref['points'][k]))) <= atol
        if 'choices' in ref:
            same = same and got['choices'][k] == ref['choices'][k]
        if not same:
            margin, tol = ref['margins'][k]
            diff_pick.append(k)
            ratio = margin / tol if tol > 0 else float('inf')
            worst = max(worst, ratio)
            if not margin <= tol:
                bad.append((k, margin, tol))
    stats = dict(steps_compared=steps, steps_with_a_different_pick=len(diff_pick),
                 steps_with_a_different_rung=diff_rung, steps_with_rungs_compared=rung_steps,
                 steps_with_different_hyper_parameters=hp_diff,
                 draws_compared=draws, draws_with_a_different_rung=draw_diff,
                 rung_difference_histogram=hist,  # ours - reference's, "no jitter" counted as -12
                 worst_margin_over_tolerance_among_differing_picks=worst,
                 differing_steps=diff_pick[:50], undocumented=bad[:10])
    return stats
