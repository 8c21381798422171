"""Continuous-context BO loop with one GP over (context, action) -- host mirror
of the reference's ``ContinuousOpt`` (/root/reference/src/cstrats/cts_opt.py) so
its driver (src/cts_ocbo.py:95-140) and option files run on the B200 engine.

Same constructor ``(function, domain, ctx_dim, options, eval_set=None,
ctx_constraints=None)``, ``get_eval_set / set_clean / optimize(max_capital,
init_pts, pre_tune=False, hp_tune_samps=None) / get_strat_name`` and histories.
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


class ContinuousOpt(object):

    def __init__(self, function, domain, ctx_dim, options, eval_set=None, ctx_constraints=None,
                 is_synthetic=True):
        self.function = function
        self.domain = domain
        self.ctx_domain, self.act_domain = domain[:ctx_dim], domain[ctx_dim:]
        self.dim, self.ctx_dim, self.act_dim = len(domain), ctx_dim, len(domain) - ctx_dim
        self.is_synthetic = is_synthetic
        if eval_set is None:
            self.eval_ctx_fidel = options.eval_ctx_fidel
            self.eval_act_fidel = options.eval_act_fidel
            self.eval_pts, self.eval_grid = self._get_eval_set()
            self._find_best_score()
        elif is_synthetic:
            self.set_eval_set(eval_set[0], eval_set[1])
        self.ctx_constraints = ctx_constraints
        self.options = options
        self.eval_ctx_fidel = options.eval_ctx_fidel
        self.eval_act_fidel = options.eval_act_fidel
        self.score_every = options.score_every
        self.tune_every = options.tune_every
        self.gp_engine = options.gp_engine
        self.cand_size = options.cand_size
        self.gp_options = load_options(euclidean_gp_args, cmd_line=False)
        self.gp_options.kernel_type = options.kernel_type
        self.gp_options.hp_tune_criterion = options.tuning_methods
        self.gp_options.hp_samples = options.hp_samples
        self.gp_options.dim = self.dim
        self.gp_options.act_dim = self.act_dim
        self.gp_options.use_additive_gp = options.use_additive_gp
        self.gp_options.add_max_group_size = options.add_max_group_size
        if hasattr(options, 'hp_search_candidates'):
            self.gp_options.hp_search_candidates = options.hp_search_candidates
        self.eval_options = deepcopy(self.gp_options)
        self.eval_options.hp_tune_criterion = 'ml'
        self._child_set_up(function, domain, ctx_dim, options)
        self.set_clean()

    def set_clean(self):
        self.gp = None
        self.gp_fitter = None
        self.x_data, self.x_init, self.y_data, self.y_init = [], [], [], []
        self.query_history, self.score_history, self.hp_history = [], [], []
        self.t = 0
        self.pre_loaded_fit = False

    def optimize(self, max_capital, init_pts, pre_tune=False, hp_tune_samps=None):
        init_rewards = [self.function(pt) for pt in init_pts]
        if hp_tune_samps is not None:
            self.pre_tune_gp(hp_tune_samps, init_pts, init_rewards)
        elif pre_tune:
            self.pre_load_points(init_pts, init_rewards)
        else:
            self.load_init_pts(init_pts, init_rewards)
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

    def load_init_pts(self, eval_pts, rewards):
        self.x_data += eval_pts
        self.y_data += rewards
        self.x_init, self.y_init = eval_pts, rewards
        self._make_new_gp_fitter()

    def pre_load_points(self, eval_pts, rewards):
        self.x_data += eval_pts
        self.y_data += rewards
        self.x_init, self.y_init = eval_pts, rewards
        self.pre_loaded_fit = True
        self.gp = get_tuned_gp(self.gp_engine, eval_pts, rewards,
                               kernel_type=self.gp_options.kernel_type)[0]

    def pre_tune_gp(self, hp_tune_samps, eval_pts, rewards):
        self.x_data += eval_pts
        self.y_data += rewards
        self.x_init, self.y_init = eval_pts, rewards
        self.pre_loaded_fit = True
        self.gp = get_best_prior(self.function, self.domain,
                                 kernel_type=self.gp_options.kernel_type,
                                 num_samples=hp_tune_samps, gp_engine=self.gp_engine)
        for pt, rew in zip(eval_pts, rewards):
            self.gp.add_data_single(pt, rew)

    def get_histories(self):
        return Namespace(query_history=self.query_history, score_history=self.score_history,
                         hp_history=self.hp_history, x_init=self.x_init, y_init=self.y_init)

    def get_eval_set(self):
        return self.eval_pts, self.eval_grid

    def set_eval_set(self, eval_pts, eval_grid):
        self.eval_pts, self.eval_grid = eval_pts, eval_grid
        self.eval_ctx_fidel, self.eval_act_fidel = eval_grid.shape[:2]
        self._find_best_score()

    def set_gp(self, gp):
        self.gp = gp
        self.pre_loaded_fit = True

    @staticmethod
    def get_strat_name():
        raise NotImplementedError('Abstract Method')

    def _determine_next_query(self):
        raise NotImplementedError('Abstract method')

    def _child_set_up(self, function, domain, ctx_dim, options):
        pass

    def _make_candidate_set(self):
        ctxs = self._get_ctx_candidates(self.cand_size).reshape(self.cand_size, self.ctx_dim)
        return np.hstack([ctxs, uniform_draw(self.act_domain, self.cand_size)])

    def _get_ctx_candidates(self, num_candidates):
        if self.ctx_constraints is None:
            lows, highs = zip(*self.ctx_domain)
            return np.random.uniform(lows, highs, (num_candidates, len(lows)))
        idxs = np.random.randint(0, len(self.ctx_constraints), num_candidates)
        return np.asarray([self.ctx_constraints[i] for i in idxs])

    def _update_history(self, pt, reward):
        self.query_history.append(Namespace(t=self.t, pt=pt, reward=reward))
        if self.is_synthetic and self.score_every is not None and self.t % self.score_every == 0:
            score = self._get_current_score()
            self.score_history.append(Namespace(t=self.t, score=score,
                                                regret=self.best_score - score))
        if self.gp is not None:
            self.hp_history.append(Namespace(t=self.t, hps=self.gp.get_kernel_hps()))

    def _draw_next_gp(self):
        if self.pre_loaded_fit:
            return self.gp
        assert self.gp_fitter is not None
        self.gp = self.gp_fitter.get_next_gp()

    def _make_new_gp_fitter(self):
        self.gp_fitter = get_gp_fitter(self.gp_engine, self.x_data, self.y_data, self.gp_options)
        self.gp_fitter.fit_gp()

    def _get_current_score(self, return_max_pts=False, eval_gp=None, e_pts=None, e_grid=None):
        """Posterior-mean policy score (cts_opt.py:297-323): fused Gram x alpha on
        the device, never materialising K(eval_pts, X)."""
        if eval_gp is None:
            if self.pre_loaded_fit:
                eval_gp = self.gp
            else:
                fitter = get_gp_fitter(self.gp_engine, self.x_data, self.y_data, self.eval_options)
                fitter.fit_gp()
                eval_gp = fitter.get_next_gp()
        if e_pts is None or e_grid is None:
            e_pts, e_grid = self.eval_pts, self.eval_grid
        mu = eval_gp.eval(e_pts)[0].reshape(e_grid.shape[:2])
        max_pts = e_grid[np.arange(e_grid.shape[0]), np.argmax(mu, axis=1), :]
        score = 0
        for max_pt in max_pts:
            score += self.function(max_pt)
        avg = score / e_grid.shape[0]
        return (avg, max_pts) if return_max_pts else avg

    def _find_best_score(self):
        evals = np.asarray([self.function(pt) for pt in self.eval_pts])
        self.best_score = np.mean(np.max(evals.reshape(self.eval_grid.shape[:2]), axis=1))

    def _get_eval_set(self, ctx_fidel=None, act_fidel=None):
        ctx_fidel = self.eval_ctx_fidel if ctx_fidel is None else ctx_fidel
        act_fidel = self.eval_act_fidel if act_fidel is None else act_fidel
        ctxs = np.repeat(uniform_draw(self.ctx_domain, ctx_fidel), act_fidel, axis=0)
        acts = uniform_draw(self.act_domain, act_fidel)
        acts = np.tile(acts.ravel(), ctx_fidel).reshape(ctx_fidel * act_fidel, self.act_dim)
        eval_pts = np.hstack([ctxs, acts])
        return eval_pts, eval_pts.reshape(ctx_fidel, act_fidel, self.ctx_dim + self.act_dim)
