"""Discrete multi-task BO loop with one GP per task -- host-side mirror of the
reference's ``MultiOpt`` (/root/reference/src/strategies/multi_opt.py) so that
its callers (src/ocbo.py:139-143,180-183) and its option files drive the B200
engine unchanged.  Same constructor, attributes, public methods and state
machine; the GP work goes through ``get_gp_fitter`` with ``gp_engine='b200'``.

State-machine facts honoured:
  * per step: draw GPs -> decide -> f(pt) -> add point -> history (-> refit)
    (reference :122-129);
  * the risk-neutral update sees the GP after ``add_data_single`` and before
    the refit (:294-296 precede :316-319);
  * the ``post_sampling`` cache: once a task holds ``hp_samples`` drawn GPs it
    keeps re-using the right-most one until that task is refit (:226-236,:286).
"""
from argparse import Namespace
from collections import deque
from copy import deepcopy

import numpy as np

from ..gp.gp_util import get_gp_fitter
from ..options import euclidean_gp_args, get_option_specs, load_options

multi_opt_args = [
    get_option_specs('tuning_methods', False, 'post_sampling',
                     'ml, post_sampling, or a comma/hyphen separated alternation of them.'),
    get_option_specs('hp_samples', False, 3, 'Hyper-parameter draws per GP fit.'),
    get_option_specs('kernel_type', False, 'se', 'GP kernel: se or matern.'),
    get_option_specs('risk_neutral', False, False, 'Track risk-neutral (posterior-mean) choices.'),
    get_option_specs('rn_eval_size', False, 10000, 'Grid size per task for risk-neutral evaluation.'),
    get_option_specs('gp_engine', False, 'b200', 'GP engine (the B200 engine is the only one bundled).'),
]


class MultiOpt(object):

    def __init__(self, f_infos, options, pre_tuned_gps=None, rn_info=None):
        self.fcns = [fi.function for fi in f_infos]
        self.domains = [fi.domain for fi in f_infos]
        self.dims = [len(d) for d in self.domains]
        self.f_names = [fi.name for fi in f_infos]
        self.pre_tuned_gps = pre_tuned_gps
        self.tune_every = options.tune_every
        self.options = options
        self.risk_neutral = options.risk_neutral
        if rn_info is not None:
            self.rn_grids, self.rn_bests = rn_info
        elif self.risk_neutral:
            self.rn_grids, self.rn_bests = self._create_rn_grids()
        else:
            self.rn_grids, self.rn_bests = None, None
        self.gp_options = load_options(euclidean_gp_args, cmd_line=False)
        self.gp_options.kernel_type = options.kernel_type
        self.gp_options.hp_tune_criterion = '-'.join(str(options.tuning_methods).split(','))
        self.gp_options.hp_samples = options.hp_samples
        if hasattr(options, 'hp_search_candidates'):
            self.gp_options.hp_search_candidates = options.hp_search_candidates
        self.hp_samples = options.hp_samples
        self.gp_engine = getattr(options, 'gp_engine', 'b200')
        self.gps = self.gp_fitters = self.sampled_gps = None
        self.query_history = self.curr_best = self.total_best = None
        self.choice_timeline = self.allocation_count = None
        self.rn_choice = self.rn_total = None
        self.t = 0
        self.set_clean()

    # ---- to be provided by the strategy ---------------------------------------
    def decide_next_query(self):
        raise NotImplementedError('Child implementation missing.')

    @staticmethod
    def get_opt_method_name():
        raise NotImplementedError('Child implementation missing.')

    def get_method_name_with_tuning(self):
        return '-'.join([self.gp_options.hp_tune_criterion, self.get_opt_method_name()])

    # ---- public loop ------------------------------------------------------------
    def optimize(self, capital, init_capital=5, init_pts=None):
        if init_pts is not None:
            self._handle_init_pts(init_pts)
        else:
            for f_idx in range(len(self.fcns)):
                self._make_random_queries(f_idx, init_capital)
        if self.pre_tuned_gps is None:
            self._update_models()
        for _ in range(capital):
            self.t += 1
            if self.pre_tuned_gps is None:
                self._draw_next_gps()
            f_idx, pt = self.decide_next_query()
            val = self.fcns[f_idx](pt)
            self._add_point_to_gp(f_idx, pt, val)
            self._update_history(f_idx, pt, val)
        return self.get_histories()

    def prep_for_optimization(self, init_evals):
        for q_info in init_evals:
            idx = self.f_names.index(q_info.f_name)
            q_info.init_pt = True
            q_info.t = 0
            self._update_history(idx, q_info.pt, q_info.val, init_pt=True, q_info=q_info)
        self._update_models()

    def draw_and_suggest_query(self):
        if self.pre_tuned_gps is None:
            self._draw_next_gps()
        return self.decide_next_query()

    def receive_feedback(self, q_info):
        self.t += 1
        idx = self.f_names.index(q_info.f_name)
        q_info.t = self.t
        self._update_history(idx, q_info.pt, q_info.val, q_info=q_info)

    def get_histories(self):
        return Namespace(query_history=self.query_history, curr_best=self.curr_best,
                         total_best=self.total_best, choice_timeline=self.choice_timeline,
                         rn_choice=self.rn_choice, rn_total=self.rn_total)

    def set_clean(self):
        if self.pre_tuned_gps is not None:
            self.gps = [deepcopy(gp) for gp in self.pre_tuned_gps]
        else:
            self.gps = [None] * len(self.f_names)
        self.gp_fitters = [None] * len(self.f_names)
        self.sampled_gps = [deque() for _ in self.f_names]
        self.query_history = {fn: [] for fn in self.f_names}
        self.curr_best = {fn: [] for fn in self.f_names}
        self.rn_choice = {fn: [] for fn in self.f_names}
        self.total_best, self.rn_total, self.choice_timeline = [], [], []
        self.allocation_count = [0] * len(self.f_names)
        self.t = 0

    # ---- hooks subclasses override ----------------------------------------------
    def _add_point_to_gp(self, f_idx, pt, val):
        self.gps[f_idx].add_data_single(pt, val)

    def _handle_init_pts(self, init_pts):
        for idx, pts in enumerate(init_pts):
            for pt in pts:
                val = self.fcns[idx](pt)
                if self.gps[idx] is not None:
                    self._add_point_to_gp(idx, pt, val)
                self._update_history(idx, pt, val, init_pt=True)

    def _make_random_queries(self, f_idx, queries):
        for _ in range(queries):
            low, high = zip(*self.domains[f_idx])
            pt = np.random.uniform(low, high)
            val = self.fcns[f_idx](pt)
            if self.gps[f_idx] is not None:
                self._add_point_to_gp(f_idx, pt, val)
            self._update_history(f_idx, pt, val, init_pt=True)

    def _draw_next_gps(self):
        assert None not in self.gp_fitters
        cycling = 'post_sampling' in self.gp_options.hp_tune_criterion and self.hp_samples > 1
        if not cycling:
            self.gps = [gpf.get_next_gp() for gpf in self.gp_fitters]
            return
        self.gps = []
        for f_idx in range(len(self.fcns)):
            cache = self.sampled_gps[f_idx]
            if len(cache) == self.hp_samples:
                nxt = cache.pop()
            else:
                nxt = self.gp_fitters[f_idx].get_next_gp()
            cache.append(nxt)
            self.gps.append(nxt)

    def _update_models(self):
        for idx in range(len(self.fcns)):
            self._update_model(idx)

    def _update_model(self, idx):
        hist = self.query_history[self.f_names[idx]]
        x_data = [h.pt for h in hist]
        y_data = [h.val for h in hist]
        gpf = get_gp_fitter(self.gp_engine, x_data, y_data, self.gp_options)
        gpf.fit_gp()
        self.gp_fitters[idx] = gpf
        self.sampled_gps[idx] = deque()

    def _rn_update(self, f_idx):
        f_name = self.f_names[f_idx]
        self.gps[f_idx].build_posterior()
        grid = self.rn_grids[f_idx]
        mean_vals = self.gps[f_idx].eval(grid)[0].ravel()
        rn_pt = grid[np.argmax(mean_vals)]
        self.rn_choice[f_name].append((self.fcns[f_idx](rn_pt), rn_pt))

    def _create_rn_grids(self):
        grids, bests = [], []
        for f_idx in range(len(self.f_names)):
            lows, highs = zip(*self.domains[f_idx])
            grid = np.random.uniform(lows, highs, (self.options.rn_eval_size, len(lows)))
            best = float('-inf')
            for pt in grid:
                best = max(best, self.fcns[f_idx](pt))
            grids.append(grid)
            bests.append(best)
        return grids, bests

    def _update_history(self, f_idx, pt, val, init_pt=False, q_info=None):
        self.allocation_count[f_idx] += 1
        if q_info is None:
            q_info = Namespace(pt=pt, val=val, init_pt=init_pt, time=self.t)
        f_name = self.f_names[f_idx]
        self.query_history[f_name].append(q_info)
        if self.risk_neutral:
            if init_pt:
                self.rn_choice[f_name].append((val, pt))
            else:
                self._rn_update(f_idx)
        bests = self.curr_best[f_name]
        if not bests or bests[-1][0] < val:
            bests.append((val, pt))
        else:
            bests.append(bests[-1])
        if self.t > 0:
            self.total_best.append(sum(b[-1][0] for b in self.curr_best.values()))
            self.choice_timeline.append(f_idx)
            if self.risk_neutral:
                self.rn_total.append(sum(r[-1][0] for r in self.rn_choice.values()))
        if self.tune_every > 0 and self.t > 0:
            if len(self.query_history[f_name]) % self.tune_every == 0:
                self._update_model(f_idx)
