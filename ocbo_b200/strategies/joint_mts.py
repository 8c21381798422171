"""Joint-GP Multi-task Thompson Sampling on the B200 engine.

Plays the role of the reference's ``JointMTS``
(/root/reference/src/strategies/joint_mts.py:11-57): ONE joint posterior sample
over [every past query, task by task ; the T x A candidate grid, task-major],
then per task ``gain = sample(new) - max sample(old)``, arg-max over tasks, then
over that task's actions (np.argmax: first maximum wins).

Only the candidate layout is assembled on the host; posterior, covariance,
Cholesky (with Dragonfly's jitter ladder), sample and the whole reduction run on
the device (``B200GPCore.sample_and_pick`` -> ocbo_segment_argmax +
ocbo_joint_pick), so (task, action index, gain) is all that comes back.
"""
import numpy as np

from .joint_opt import JointOpt, located_rows
from .. import inflight
from ..util import sample_grid


def joint_layout(f_locs, per_task_points, grid):
    """Rows to sample and their segmentation: T segments of old points (in task
    order) followed by T grid segments of equal length."""
    T = len(f_locs)
    counts = [len(p) for p in per_task_points]
    n_old = sum(counts)
    width = grid.shape[1]
    blocks = [located_rows(f_locs[t], per_task_points[t]) for t in range(T) if counts[t]]
    rows = np.concatenate([b.reshape(-1, width) for b in blocks] + [grid], axis=0)
    per_task = len(grid) // T
    bounds = np.concatenate([[0], np.cumsum(counts), n_old + per_task * np.arange(1, T + 1)])
    return rows, bounds.astype(np.int32)


class JointMTS(JointOpt):

    @staticmethod
    def get_opt_method_name():
        return 'joint-mts'

    def decide_next_query(self):
        if not inflight.in_flight():  # (in flight the merged pass rebuilds all trials' posteriors)
            self.gps[0].build_posterior()
        return self._find_best_query()

    def _find_best_query(self):
        T, A = len(self.fcns), self.options.max_opt_evals
        grid = sample_grid(self.f_locs, self.domains[0], A)       # consumes the global stream first
        asked = [[q.pt for q in self.query_history[name]] for name in self.f_names]
        rows, bounds = joint_layout(self.f_locs, asked, grid)
        task, action, _ = self.gps[0].gp_core.sample_and_pick(rows, bounds, T)
        return task, grid[task * A + action, len(self.f_locs[0]):]
