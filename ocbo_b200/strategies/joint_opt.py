"""Multi-task BO with ONE joint GP over (task location (+) action).

Plays the role of the reference's ``JointOpt``
(/root/reference/src/strategies/joint_opt.py:16-124).  Everything that differs
from the independent-GP loop is expressed through ``MultiOpt``'s hooks: points
are prefixed with the task location, a single fitter sees all observations, one
GP object is shared by every task (so ``add_data_single`` through any task
mutates the shared model, reference :45-48,:109), and the risk-neutral grids are
task-located ``sample_grid`` draws.
"""
import numpy as np

from .multi_opt import MultiOpt
from ..gp.gp_util import get_gp_fitter
from ..util import sample_grid


def located_rows(f_loc, pts):
    """[f_loc (+) pt for pt in pts] as one (len(pts), c + d) array (np.append per point was
    a third of a joint step's host time at n = 160)."""
    f_loc = np.asarray(f_loc, dtype=np.float64).ravel()
    c = f_loc.size
    if not len(pts):
        return np.zeros((0, c))
    P = np.asarray(pts, dtype=np.float64).reshape(len(pts), -1)
    rows = np.empty((len(pts), c + P.shape[1]))
    rows[:, :c] = f_loc
    rows[:, c:] = P
    return rows


class JointOpt(MultiOpt):

    def __init__(self, f_infos, options, pre_tuned_gps=None, rn_info=None):
        missing = [fi.name for fi in f_infos if not hasattr(fi, 'f_loc')]
        if missing:
            raise ValueError('Function %s does not have a location' % missing[0])
        self.f_locs = [fi.f_loc for fi in f_infos]  # needed by the grid hooks during base init
        super(JointOpt, self).__init__(f_infos, options, pre_tuned_gps, rn_info)
        loc_dim = len(self.f_locs[0])
        self.gp_options.dim = self.dims[0] + loc_dim
        self.gp_options.act_dim = loc_dim  # the reference stores the location size here (:41-43)

    def set_clean(self):
        super(JointOpt, self).set_clean()
        self.next_explore_idx = 0
        self.num_explores = 0

    # ------------------------------------------------------------------ hooks
    def _fresh_pretuned(self):
        return list(self.pre_tuned_gps)  # shared, not copied (:52)

    def _gp_input(self, f_idx, pt):
        return np.append(self.f_locs[f_idx], pt)

    def _refit(self, f_idx):
        blocks, ys = [], []
        for t_idx in range(len(self.f_names)):
            pts, vals = self._ledger.observations(t_idx)
            blocks.append(located_rows(self.f_locs[t_idx], pts))
            ys += vals
        xs = np.concatenate(blocks, axis=0)  # rows = _gp_input(task, pt), task by task
        fitter = get_gp_fitter(self.gp_engine, xs, ys, self.gp_options)
        fitter.fit_gp()
        self.gp_fitters = [fitter]

    def _update_models(self):
        self._refit(0)  # one joint fit, not one per task

    def _update_model(self, idx):
        # the reference's JointOpt routes the per-task hook through the plural one
        # (src/strategies/joint_opt.py:123-124), so a subclass that overrides ``_update_models``
        # (joint-rand: no model at all unless risk-neutral) changes both
        self._update_models()

    def _next_gps(self):
        shared = self.gp_fitters[0].get_next_gp()
        if 'post_sampling' in self.gp_options.hp_tune_criterion and self.tune_every > 1:
            shared.build_posterior()
        return [shared] * len(self.fcns)

    def _rn_grid(self, f_idx):
        return sample_grid([self.f_locs[f_idx]], self.domains[0], self.options.max_opt_evals)

    def _rn_query(self, f_idx, grid_row):
        return grid_row[len(self.f_locs[0]):]
