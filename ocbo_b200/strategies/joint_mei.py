"""Joint-GP multi-task expected improvement -- host mirror of
/root/reference/src/strategies/joint_mei.py:11-52: per task, EI on a grid of
max_opt_evals actions at the task's location against the task's incumbent;
arg-max over tasks (np.argmax: first maximum).  The T grids are one strided
batch sharing the joint GP's Cholesky factor."""
import numpy as np

from .joint_opt import JointOpt
from .. import batched
from ..util import sample_grid


class JointMEI(JointOpt):

    def decide_next_query(self):
        self.gps[0].build_posterior()
        return self._find_best_query()

    @staticmethod
    def get_opt_method_name():
        return 'joint-mei'

    def _find_best_query(self):
        grids, bests = [], []
        for f_idx in range(len(self.fcns)):
            bests.append(self.curr_best[self.f_names[f_idx]][-1][0])
            grids.append(sample_grid([self.f_locs[f_idx]], self.domains[0],
                                     self.options.max_opt_evals))
        ei_max, ei_arg = batched.shared_ei_step(self.gps[0], grids, best_list=bests)
        query_idx = int(np.argmax(ei_max))
        loc_dim = len(self.f_locs[0])
        return query_idx, grids[query_idx][int(ei_arg[query_idx]), loc_dim:]
