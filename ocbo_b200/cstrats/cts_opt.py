"""Continuous-context BO loop (one GP over context (+) action) on the B200 engine.

Plays the role of the reference's ``ContinuousOpt``
(/root/reference/src/cstrats/cts_opt.py): constructor ``(function, domain,
ctx_dim, options, eval_set=None, ctx_constraints=None)``, ``optimize(max_capital,
init_pts, pre_tune=False, hp_tune_samps=None)``, ``get_eval_set / set_eval_set /
set_clean / set_gp / suggest_next_eval / receive_feedback / get_strat_name`` and
the histories its driver stores (src/cts_ocbo.py:95-140).  The organisation is
our own: the scoring grid and the posterior-mean policy score live in
``PolicyScore``; the loop is "suggest -> evaluate -> absorb".

Behaviour kept: hyper-parameters are refit every ``tune_every`` evaluations;
every ``score_every`` evaluations a separate ML-tuned GP scores the greedy policy
on the evaluation grid (reference :297-323) -- on the device this is one fused
Gram x alpha pass over the 10^4 grid points; the global numpy stream is consumed
in the reference's order.
"""
from argparse import Namespace
from copy import deepcopy

import numpy as np

from ..gp.gp_util import get_best_prior, get_gp_fitter, get_tuned_gp
from ..options import euclidean_gp_args, get_option_specs, load_options
from ..util import uniform_draw

cts_opt_args = [
    get_option_specs('tuning_methods', False, 'ml', 'ml, post_sampling or a hyphenated alternation.'),
    get_option_specs('tune_every', False, 1, 'Refit hyper-parameters every this many evaluations.'),
    get_option_specs('kernel_type', False, 'se', 'GP kernel: se or matern.'),
    get_option_specs('use_additive_gp', False, False, 'Additive GP (not supported by the B200 engine).'),
    get_option_specs('add_max_group_size', False, 6, 'Unused (additive GP).'),
    get_option_specs('hp_samples', False, 3, 'Hyper-parameter draws per GP fit.'),
    get_option_specs('eval_ctx_fidel', False, 100, 'Contexts in the scoring grid.'),
    get_option_specs('eval_act_fidel', False, 100, 'Actions per context in the scoring grid.'),
    get_option_specs('score_every', False, 1, 'Score the policy every this many evaluations.'),
    get_option_specs('cand_size', False, 100, 'Candidate points per suggestion.'),
    get_option_specs('gp_engine', False, 'b200', 'GP engine (the B200 engine is the only one bundled).'),
]


class PolicyScore(object):
    """Evaluation grid (contexts x actions) and the greedy posterior-mean policy score."""

    def __init__(self, pts, grid):
        self.pts, self.grid = pts, grid

    @classmethod
    def draw(cls, ctx_domain, act_domain, ctx_fidel, act_fidel):
        """Random grid: every context paired with the SAME random actions (:332-350)."""
        ctxs = uniform_draw(ctx_domain, ctx_fidel)
        acts = uniform_draw(act_domain, act_fidel)
        pts = np.hstack([np.repeat(ctxs, act_fidel, axis=0), np.tile(acts, (ctx_fidel, 1))])
        return cls(pts, pts.reshape(ctx_fidel, act_fidel, pts.shape[1]))

    @property
    def shape(self):
        return self.grid.shape[:2]

    @staticmethod
    def _values(function, pts):
        """function at every point: through its vectorised form when it has one (bitwise equal
        values by contract, ocbo_b200/synth.py), else point by point like the reference."""
        batch = getattr(function, 'batch', None)
        if batch is not None:
            return np.asarray(batch(pts), dtype=np.float64)
        return np.asarray([function(p) for p in pts])

    def oracle_best(self, function):
        vals = self._values(function, self.pts).reshape(self.shape)
        return np.mean(np.max(vals, axis=1))

    def greedy(self, gp, function, pts=None, grid=None):
        pts = self.pts if pts is None else pts
        grid = self.grid if grid is None else grid
        mu = gp.eval(pts)[0].reshape(grid.shape[:2])
        chosen = grid[np.arange(grid.shape[0]), np.argmax(mu, axis=1), :]
        total = 0
        for v in self._values(function, chosen).tolist():  # same left-to-right sum
            total += v
        return total / grid.shape[0], chosen


class ContinuousOpt(object):

    def __init__(self, function, domain, ctx_dim, options, eval_set=None, ctx_constraints=None,
                 is_synthetic=True):
        self.function, self.domain = function, domain
        self.ctx_dim, self.dim = ctx_dim, len(domain)
        self.act_dim = self.dim - ctx_dim
        self.ctx_domain, self.act_domain = domain[:ctx_dim], domain[ctx_dim:]
        self.is_synthetic = is_synthetic
        self.ctx_constraints = ctx_constraints
        self.options = options
        if eval_set is None:
            self._install_score(PolicyScore.draw(self.ctx_domain, self.act_domain,
                                                 options.eval_ctx_fidel, options.eval_act_fidel))
        elif is_synthetic:
            self.set_eval_set(eval_set[0], eval_set[1])
        self.eval_ctx_fidel, self.eval_act_fidel = options.eval_ctx_fidel, options.eval_act_fidel
        self.score_every, self.tune_every = options.score_every, options.tune_every
        self.gp_engine, self.cand_size = options.gp_engine, options.cand_size
        gp_opts = load_options(euclidean_gp_args, cmd_line=False)
        for key, val in (('kernel_type', options.kernel_type),
                         ('hp_tune_criterion', options.tuning_methods),
                         ('hp_samples', options.hp_samples), ('dim', self.dim),
                         ('act_dim', self.act_dim), ('use_additive_gp', options.use_additive_gp),
                         ('add_max_group_size', options.add_max_group_size)):
            setattr(gp_opts, key, val)
        for key in ('hp_search_candidates', 'hp_search_rounds'):
            if hasattr(options, key):
                setattr(gp_opts, key, getattr(options, key))
        self.gp_options = gp_opts
        self.eval_options = deepcopy(gp_opts)
        self.eval_options.hp_tune_criterion = 'ml'  # the scoring GP is always ML-tuned (:99-100)
        self._child_set_up(function, domain, ctx_dim, options)
        self.set_clean()

    # ------------------------------------------------------------------ evaluation grid
    def _install_score(self, score):
        self._score = score
        self.eval_pts, self.eval_grid = score.pts, score.grid
        self.best_score = score.oracle_best(self.function)

    def get_eval_set(self):
        return self.eval_pts, self.eval_grid

    def set_eval_set(self, eval_pts, eval_grid):
        self._install_score(PolicyScore(eval_pts, eval_grid))
        self.eval_ctx_fidel, self.eval_act_fidel = eval_grid.shape[:2]

    # ------------------------------------------------------------------ run state
    def set_clean(self):
        self.gp = self.gp_fitter = None
        self.x_data, self.y_data, self.x_init, self.y_init = [], [], [], []
        self.query_history, self.score_history, self.hp_history = [], [], []
        self.pre_loaded_fit = False
        self.t = 0

    def get_histories(self):
        return Namespace(query_history=self.query_history, score_history=self.score_history,
                         hp_history=self.hp_history, x_init=self.x_init, y_init=self.y_init)

    def set_gp(self, gp):
        self.gp, self.pre_loaded_fit = gp, True

    # ------------------------------------------------------------------ main loop
    def optimize(self, max_capital, init_pts, pre_tune=False, hp_tune_samps=None):
        rewards = [self.function(pt) for pt in init_pts]
        if hp_tune_samps is not None:
            self.pre_tune_gp(hp_tune_samps, init_pts, rewards)
        elif pre_tune:
            self.pre_load_points(init_pts, rewards)
        else:
            self.load_init_pts(init_pts, rewards)
        for _ in range(max_capital):
            pt = self.suggest_next_eval()
            self.receive_feedback(pt, self.function(pt))
        return self.get_histories()

    def suggest_next_eval(self):
        self._draw_next_gp()
        return self._determine_next_query()

    def receive_feedback(self, eval_pt, reward):
        self.x_data.append(eval_pt)
        self.y_data.append(reward)
        if self.gp is not None:
            self.gp.add_data_single(eval_pt, reward)
        self.t += 1
        self._update_history(eval_pt, reward)
        if not self.pre_loaded_fit and self.t % self.tune_every == 0:
            self._make_new_gp_fitter()

    # three ways of starting a run (:170-214)
    def _remember_init(self, eval_pts, rewards):
        self.x_data += eval_pts
        self.y_data += rewards
        self.x_init, self.y_init = eval_pts, rewards

    def load_init_pts(self, eval_pts, rewards):
        self._remember_init(eval_pts, rewards)
        self._make_new_gp_fitter()

    def pre_load_points(self, eval_pts, rewards):
        self._remember_init(eval_pts, rewards)
        self.set_gp(get_tuned_gp(self.gp_engine, eval_pts, rewards,
                                 kernel_type=self.gp_options.kernel_type)[0])

    def pre_tune_gp(self, hp_tune_samps, eval_pts, rewards):
        self._remember_init(eval_pts, rewards)
        self.set_gp(get_best_prior(self.function, self.domain, kernel_type=self.gp_options.kernel_type,
                                   num_samples=hp_tune_samps, gp_engine=self.gp_engine))
        for pt, rew in zip(eval_pts, rewards):
            self.gp.add_data_single(pt, rew)

    # ------------------------------------------------------------------ to be specialised
    @staticmethod
    def get_strat_name():
        raise NotImplementedError('Abstract Method')

    def _determine_next_query(self):
        raise NotImplementedError('Abstract method')

    def _child_set_up(self, function, domain, ctx_dim, options):
        pass

    # ------------------------------------------------------------------ helpers
    def _get_ctx_candidates(self, num_candidates):
        if self.ctx_constraints is not None:
            picks = np.random.randint(0, len(self.ctx_constraints), num_candidates)
            return np.asarray([self.ctx_constraints[i] for i in picks])
        return uniform_draw(self.ctx_domain, num_candidates)

    def _make_candidate_set(self):
        ctxs = self._get_ctx_candidates(self.cand_size).reshape(self.cand_size, self.ctx_dim)
        return np.hstack([ctxs, uniform_draw(self.act_domain, self.cand_size)])

    def _draw_next_gp(self):
        if self.pre_loaded_fit:
            return self.gp
        assert self.gp_fitter is not None
        self.gp = self.gp_fitter.get_next_gp()

    def _make_new_gp_fitter(self):
        self.gp_fitter = get_gp_fitter(self.gp_engine, self.x_data, self.y_data, self.gp_options)
        self.gp_fitter.fit_gp()

    def _scoring_gp(self):
        if self.pre_loaded_fit:
            return self.gp
        fitter = get_gp_fitter(self.gp_engine, self.x_data, self.y_data, self.eval_options)
        fitter.fit_gp()
        return fitter.get_next_gp()

    def _get_current_score(self, return_max_pts=False, eval_gp=None, e_pts=None, e_grid=None):
        gp = self._scoring_gp() if eval_gp is None else eval_gp
        if e_pts is None or e_grid is None:
            e_pts = e_grid = None
        avg, chosen = self._score.greedy(gp, self.function, e_pts, e_grid)
        return (avg, chosen) if return_max_pts else avg

    def _find_best_score(self):
        self.best_score = self._score.oracle_best(self.function)

    def _update_history(self, pt, reward):
        self.query_history.append(Namespace(t=self.t, pt=pt, reward=reward))
        scoring = self.is_synthetic and self.score_every is not None and self.t % self.score_every == 0
        if scoring:
            score = self._get_current_score()
            self.score_history.append(Namespace(t=self.t, score=score, regret=self.best_score - score))
        if self.gp is not None:
            self.hp_history.append(Namespace(t=self.t, hps=self.gp.get_kernel_hps()))
