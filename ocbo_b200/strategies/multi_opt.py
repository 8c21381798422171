"""Discrete multi-task BO loop (one GP per task) on the B200 engine.

Plays the role of the reference's ``MultiOpt``
(/root/reference/src/strategies/multi_opt.py): same constructor
``(f_infos, options, pre_tuned_gps=None, rn_info=None)``, same public methods
(``optimize, set_clean, get_histories, prep_for_optimization,
draw_and_suggest_query, receive_feedback, get_method_name_with_tuning``) and the
same history attributes, so the drivers (src/ocbo.py:139-143,180-183) and the
option files work unchanged.  The organisation is our own: bookkeeping lives in
``TaskLedger``; what differs between independent-GP and joint-GP strategies is
expressed through four small hooks (``_gp_input``, ``_refit``, ``_next_gps``,
``_rn_query``) instead of re-implemented loops.

Behaviour kept from the reference (SURVEY 8a "state-machine facts"):
  * per step: draw GPs -> decide -> f(pt) -> add point -> log (-> refit)   (:122-129)
  * the risk-neutral pick sees the GP after the new point, before the refit (:294-296,:316-319)
  * the ``post_sampling`` cache: once a task holds ``hp_samples`` drawn GPs the
    right-most one is re-used until that task is refit                     (:226-236,:286)
  * the global numpy stream is consumed in the same order
"""
from argparse import Namespace
from collections import deque
from copy import deepcopy

import numpy as np

from .ledger import TaskLedger
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


def _box(domain):
    lows, highs = zip(*domain)
    return lows, highs


class MultiOpt(object):

    def __init__(self, f_infos, options, pre_tuned_gps=None, rn_info=None):
        self.fcns = [fi.function for fi in f_infos]
        self.domains = [fi.domain for fi in f_infos]
        self.dims = [len(d) for d in self.domains]
        self.f_names = [fi.name for fi in f_infos]
        self.options = options
        self.pre_tuned_gps = pre_tuned_gps
        self.tune_every = options.tune_every
        self.risk_neutral = options.risk_neutral
        self.hp_samples = options.hp_samples
        self.gp_engine = getattr(options, 'gp_engine', 'b200')
        # risk-neutral evaluation grids are shared between methods through rn_info
        if rn_info is not None:
            self.rn_grids, self.rn_bests = rn_info
        elif self.risk_neutral:
            grids = [self._rn_grid(f) for f in range(len(self.fcns))]
            self.rn_grids = grids
            self.rn_bests = [self._grid_optimum(f, g) for f, g in enumerate(grids)]
        else:
            self.rn_grids = self.rn_bests = None
        gp_opts = load_options(euclidean_gp_args, cmd_line=False)
        gp_opts.kernel_type = options.kernel_type
        gp_opts.hp_tune_criterion = str(options.tuning_methods).replace(',', '-')
        gp_opts.hp_samples = options.hp_samples
        for key in ('hp_search_candidates', 'hp_search_rounds'):
            if hasattr(options, key):
                setattr(gp_opts, key, getattr(options, key))
        self.gp_options = gp_opts
        self.t = 0
        self.set_clean()

    # ------------------------------------------------------------------ strategy API
    def decide_next_query(self):
        raise NotImplementedError('Child implementation missing.')

    @staticmethod
    def get_opt_method_name():
        raise NotImplementedError('Child implementation missing.')

    def get_method_name_with_tuning(self):
        return '%s-%s' % (self.gp_options.hp_tune_criterion, self.get_opt_method_name())

    # ------------------------------------------------------------------ run state
    def set_clean(self):
        num = len(self.f_names)
        self.gps = self._fresh_pretuned() if self.pre_tuned_gps is not None else [None] * num
        self.gp_fitters = [None] * num
        self.sampled_gps = [deque() for _ in range(num)]
        self._ledger = TaskLedger(self.f_names, self.risk_neutral)
        # the reference's attribute names, bound to the ledger's containers
        led = self._ledger
        self.query_history, self.curr_best, self.rn_choice = led.query_history, led.curr_best, led.rn_choice
        self.total_best, self.rn_total = led.total_best, led.rn_total
        self.choice_timeline, self.allocation_count = led.choice_timeline, led.allocation_count
        self.t = 0

    def _fresh_pretuned(self):
        return [deepcopy(gp) for gp in self.pre_tuned_gps]

    def get_histories(self):
        return self._ledger.as_histories()

    # ------------------------------------------------------------------ main loop
    def optimize(self, capital, init_capital=5, init_pts=None):
        tuned = self.pre_tuned_gps is not None
        if init_pts is None:
            for f_idx in range(len(self.fcns)):
                for _ in range(init_capital):
                    self._evaluate_and_log(f_idx, np.random.uniform(*_box(self.domains[f_idx])), True)
        else:
            for f_idx, pts in enumerate(init_pts):
                for pt in pts:
                    self._evaluate_and_log(f_idx, pt, True)
        if not tuned:
            self._update_models()
        for _ in range(capital):
            self.t += 1
            if not tuned:
                self._draw_next_gps()
            f_idx, pt = self.decide_next_query()
            self._evaluate_and_log(f_idx, pt, False)
        return self.get_histories()

    def _evaluate_and_log(self, f_idx, pt, is_init):
        val = self.fcns[f_idx](pt)
        if self.gps[f_idx] is not None:
            self._add_point_to_gp(f_idx, pt, val)
        self._update_history(f_idx, pt, val, init_pt=is_init)

    # ask / tell interface (reference :133-167)
    def prep_for_optimization(self, init_evals):
        for q_info in init_evals:
            q_info.init_pt, q_info.t = True, 0
            self._update_history(self.f_names.index(q_info.f_name), q_info.pt, q_info.val,
                                 init_pt=True, q_info=q_info)
        self._update_models()

    def draw_and_suggest_query(self):
        if self.pre_tuned_gps is None:
            self._draw_next_gps()
        return self.decide_next_query()

    def receive_feedback(self, q_info):
        self.t += 1
        q_info.t = self.t
        self._update_history(self.f_names.index(q_info.f_name), q_info.pt, q_info.val, q_info=q_info)

    # ------------------------------------------------------------------ hooks
    def _gp_input(self, f_idx, pt):
        """The point as task ``f_idx``'s GP sees it."""
        return pt

    def _add_point_to_gp(self, f_idx, pt, val):
        self.gps[f_idx].add_data_single(self._gp_input(f_idx, pt), val)

    def _refit(self, f_idx):
        """New hyper-parameter fit for task ``f_idx`` from its own observations."""
        xs, ys = self._ledger.observations(f_idx)
        fitter = get_gp_fitter(self.gp_engine, xs, ys, self.gp_options)
        fitter.fit_gp()
        self.gp_fitters[f_idx] = fitter
        self.sampled_gps[f_idx] = deque()

    def _next_gps(self):
        """One GP per task for the coming decision."""
        cycling = 'post_sampling' in self.gp_options.hp_tune_criterion and self.hp_samples > 1
        out = []
        for f_idx, fitter in enumerate(self.gp_fitters):
            if not cycling:
                out.append(fitter.get_next_gp())
                continue
            cache = self.sampled_gps[f_idx]
            gp = cache.pop() if len(cache) == self.hp_samples else fitter.get_next_gp()
            cache.append(gp)
            out.append(gp)
        return out

    def _rn_grid(self, f_idx):
        return np.random.uniform(*_box(self.domains[f_idx]),
                                 (self.options.rn_eval_size, len(self.domains[f_idx])))

    def _rn_query(self, f_idx, grid_row):
        """Action to evaluate for a row of the risk-neutral grid."""
        return grid_row

    def _grid_optimum(self, f_idx, grid):
        best = float('-inf')
        for row in grid:
            best = max(best, self.fcns[f_idx](self._rn_query(f_idx, row)))
        return best

    # names the reference's subclasses / callers use
    def _update_models(self):
        for f_idx in range(len(self.fcns)):
            self._update_model(f_idx)

    def _update_model(self, idx):
        self._refit(idx)

    def _draw_next_gps(self):
        assert None not in self.gp_fitters
        self.gps = self._next_gps()

    def _rn_update(self, f_idx):
        """Risk-neutral pick: arg-max of the posterior mean on the task's grid (fused
        Gram x alpha on the device), evaluated on the true function."""
        gp = self.gps[f_idx]
        gp.build_posterior()
        grid = self.rn_grids[f_idx]
        row = grid[np.argmax(gp.eval(grid)[0].ravel())]
        pt = self._rn_query(f_idx, row)
        return self.fcns[f_idx](pt), pt

    def _update_history(self, f_idx, pt, val, init_pt=False, q_info=None):
        if q_info is None:
            q_info = Namespace(pt=pt, val=val, init_pt=init_pt, time=self.t)
        rn_entry = None
        if self.risk_neutral:
            rn_entry = (val, pt) if init_pt else self._rn_update(f_idx)
        self._ledger.record(f_idx, q_info, self.t, rn_entry)
        due = self.tune_every > 0 and self.t > 0 and \
            self._ledger.num_queries(f_idx) % self.tune_every == 0
        if due:
            self._update_model(f_idx)
