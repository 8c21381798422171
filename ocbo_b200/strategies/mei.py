"""Multi-task expected improvement with independent GPs -- host mirror of
/root/reference/src/strategies/mei.py:13-46 (acquisition: src/util/misc_util.py:
368-385, random-search EI over max_opt_evals uniform points per task).  All tasks
are evaluated in one batched device pass; EI comes from the fused mean + diagonal
variance kernel, the m x m covariance the reference builds is never formed."""
from .multi_opt import MultiOpt
from .. import batched
from ..util import uniform_draw


class MEI(MultiOpt):

    def decide_next_query(self):
        pts, bests = [], []
        for idx in range(len(self.fcns)):
            pts.append(uniform_draw(self.domains[idx], self.options.max_opt_evals))
            bests.append(self.curr_best[self.f_names[idx]][-1][0])
        per_task = batched.mei_step(self.gps, pts, bests)
        best_idx, best_pt, most_gain = None, None, float('-inf')
        for idx, (ei_val, ei_arg) in enumerate(per_task):
            if ei_val > most_gain:
                most_gain, best_idx, best_pt = ei_val, idx, pts[idx][ei_arg]
        return best_idx, best_pt

    @staticmethod
    def get_opt_method_name():
        return 'mei'
