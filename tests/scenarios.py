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
}
CONTINUOUS = {
    # conth42.txt: Hartmann-6 with 4 context + 2 action dims
    'conth42-cmts': dict(kind='cmts', act_dim=2, capital=6, init_capital=5, num_profiles=10,
                         profile_evals=100, eval_ctx_fidel=8, eval_act_fidel=8),
    'conth42-cmts-pm': dict(kind='cmts-pm', act_dim=2, capital=6, init_capital=5, num_profiles=10,
                            profile_evals=100, eval_ctx_fidel=8, eval_act_fidel=8),
}


def discrete_options(cfg, gp_engine):
    return Namespace(tuning_methods=cfg['tuning_methods'], hp_samples=cfg['hp_samples'],
                     kernel_type='se', risk_neutral=cfg['risk_neutral'], rn_eval_size=50,
                     gp_engine=gp_engine, tune_every=1, max_opt_evals=100)


def continuous_options(cfg, gp_engine):
    return Namespace(tuning_methods='ml', tune_every=1, kernel_type='se', use_additive_gp=False,
                     add_max_group_size=6, hp_samples=3, eval_ctx_fidel=cfg['eval_ctx_fidel'],
                     eval_act_fidel=cfg['eval_act_fidel'], score_every=1, cand_size=100,
                     gp_engine=gp_engine, num_profiles=cfg['num_profiles'],
                     profile_evals=cfg['profile_evals'])


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
                regret=[float(s.regret) for s in h.score_history])
