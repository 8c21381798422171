"""Multi-task BO with ONE joint GP over (task location, action) -- host mirror
of /root/reference/src/strategies/joint_opt.py:16-124: a single fitter shared by
all tasks, points prefixed with the task location, the same GP object aliased
for every task (:109)."""
import numpy as np

from .multi_opt import MultiOpt
from ..gp.gp_util import get_gp_fitter
from ..util import sample_grid


class JointOpt(MultiOpt):

    def __init__(self, f_infos, options, pre_tuned_gps=None, rn_info=None):
        self.f_locs = []
        for f_inf in f_infos:
            if not hasattr(f_inf, 'f_loc'):
                raise ValueError('Function %s does not have a location' % f_inf.name)
            self.f_locs.append(f_inf.f_loc)
        super(JointOpt, self).__init__(f_infos, options, pre_tuned_gps, rn_info)
        loc_dim = len(self.f_locs[0])
        self.gp_options.dim = self.dims[0] + loc_dim
        self.gp_options.act_dim = loc_dim  # (sic) the reference stores the location dim here (:43)

    def _add_point_to_gp(self, f_idx, pt, val):
        self.gps[f_idx].add_data_single(np.append(self.f_locs[f_idx], pt), val)

    def set_clean(self):
        if self.pre_tuned_gps is not None:
            self.gps = [gp for gp in self.pre_tuned_gps]  # shared, not copied (:52)
        else:
            self.gps = [None] * len(self.f_names)
        self.gp_fitters = [None] * len(self.f_names)
        self.query_history = {fn: [] for fn in self.f_names}
        self.curr_best = {fn: [] for fn in self.f_names}
        self.rn_choice = {fn: [] for fn in self.f_names}
        self.total_best, self.rn_total, self.choice_timeline = [], [], []
        self.allocation_count = [0] * len(self.f_names)
        self.t = 0
        self.next_explore_idx = 0
        self.num_explores = 0

    def _rn_update(self, f_idx):
        f_name = self.f_names[f_idx]
        self.gps[f_idx].build_posterior()
        grid = self.rn_grids[f_idx]
        mean_vals = self.gps[f_idx].eval(grid)[0].ravel()
        rn_pt = grid[np.argmax(mean_vals), len(self.f_locs[0]):]
        self.rn_choice[f_name].append((self.fcns[f_idx](rn_pt), rn_pt))

    def _create_rn_grids(self):
        grids, bests = [], []
        loc_dim = len(self.f_locs[0])
        for f_idx in range(len(self.f_names)):
            grid = sample_grid([self.f_locs[f_idx]], self.domains[0], self.options.max_opt_evals)
            best = float('-inf')
            for pt in grid:
                best = max(best, self.fcns[f_idx](pt[loc_dim:]))
            grids.append(grid)
            bests.append(best)
        return grids, bests

    def _draw_next_gps(self):
        assert None not in self.gp_fitters
        next_gp = self.gp_fitters[0].get_next_gp()
        if 'post_sampling' in self.gp_options.hp_tune_criterion and self.tune_every > 1:
            next_gp.build_posterior()
        self.gps = [next_gp for _ in self.fcns]

    def _update_models(self):
        x_data, y_data = [], []
        for f_idx, f_name in enumerate(self.f_names):
            for hist in self.query_history[f_name]:
                x_data.append(np.append(self.f_locs[f_idx], hist.pt))
                y_data.append(hist.val)
        gpf = get_gp_fitter(self.gp_engine, x_data, y_data, self.gp_options)
        gpf.fit_gp()
        self.gp_fitters = [gpf]

    def _update_model(self, idx):
        self._update_models()
