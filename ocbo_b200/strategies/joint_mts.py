"""Joint-GP Multi-task Thompson Sampling -- host mirror of
/root/reference/src/strategies/joint_mts.py:11-57.

One joint posterior sample over [all past points (task by task) ; T x A grid
(task-major)]; the per-task reduction (old best, gains, arg-max task then
action: :43-57) runs on the device (ocbo_segment_argmax + ocbo_joint_pick), so
only (task, action index, gain) comes back to the host."""
import numpy as np

from .joint_opt import JointOpt
from ..util import sample_grid


class JointMTS(JointOpt):

    def decide_next_query(self):
        self.gps[0].build_posterior()
        return self._find_best_query()

    @staticmethod
    def get_opt_method_name():
        return 'joint-mts'

    def _find_best_query(self):
        A = self.options.max_opt_evals
        T = len(self.fcns)
        grid_pts = sample_grid(self.f_locs, self.domains[0], A)
        past_pts, num_qs = [], []
        for f_idx, f_name in enumerate(self.f_names):
            for hist in self.query_history[f_name]:
                past_pts.append(np.append(self.f_locs[f_idx], hist.pt))
            num_qs.append(len(self.query_history[f_name]))
        total_qs = sum(num_qs)
        to_samp = np.vstack([np.asarray(past_pts).reshape(total_qs, -1), grid_pts])
        divs = np.concatenate([[0], np.cumsum(num_qs)])
        seg = np.concatenate([divs, total_qs + A * np.arange(1, T + 1)]).astype(np.int32)
        best_ctx, best_idx, _ = self.gps[0].gp_core.sample_and_pick(to_samp, seg, T)
        loc_dim = len(self.f_locs[0])
        best_pt = grid_pts.reshape((T, A, -1))[best_ctx, best_idx, :].ravel()
        return best_ctx, best_pt[loc_dim:]
