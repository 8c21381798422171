"""REVI with a joint GP over tasks -- host mirror of
/root/reference/src/strategies/corr_opt.py:23-145.

Per decision: a judgement grid (num_eval_pts actions per task) and a candidate grid
(num_candidates actions per task), both task-major ``sample_grid`` draws in that order; every
candidate is scored by the sum over tasks of the knowledge gradient of the task's judgement
means under an observation at the candidate (misc_util.py:180-212); the first best candidate
is queried.  The reference loops over T * num_candidates * T Python calls of
``knowledge_gradient``; here all of them are one device pass (batched.revi_improvements).
After ``run_ei_after`` steps the method plays round-robin EI (:66-80).
"""
import numpy as np

from .joint_opt import JointOpt
from .. import batched
from ..options import get_option_specs
from ..util import sample_grid

corr_args = [
    get_option_specs('num_candidates', False, 100, 'Candidate actions per task for REVI.'),
]


def first_best(values):
    """Index the reference's ``if v > best`` scan keeps: the first maximum, NaNs never win.
    If every value is NaN (a posterior so flat that all judgement lines coincide: 0/0 in the
    envelope) the reference dereferences ``None``; here that is a ValueError."""
    best_idx, best_val = None, float('-inf')
    for idx, v in enumerate(values):
        if v > best_val:
            best_idx, best_val = idx, v
    if best_idx is None:
        raise ValueError('REVI: no candidate has a finite knowledge gradient (degenerate posterior)')
    return best_idx


class CorrelatedOpt(JointOpt):

    def __init__(self, f_infos, options, pre_tuned_gps=None, rn_info=None):
        self.round_robin_idx = 0
        super(CorrelatedOpt, self).__init__(f_infos, options, pre_tuned_gps, rn_info)

    def _child_decide_query(self):
        raise NotImplementedError('To be implemented in child.')

    def decide_next_query(self):
        self.gps[0].build_posterior()
        run_ei_after = getattr(self.options, 'run_ei_after', None)
        if run_ei_after is None or self.t < float(run_ei_after):
            return self._child_decide_query()
        return self._play_ei()

    def _play_ei(self):
        f_idx = self.round_robin_idx
        self.round_robin_idx = (self.round_robin_idx + 1) % len(self.fcns)
        curr_best = self.curr_best[self.f_names[f_idx]][-1][0]
        ei_pts = sample_grid([self.f_locs[f_idx]], self.domains[0], self.options.max_opt_evals)
        _, ei_arg = batched.shared_ei_step(self.gps[0], [ei_pts], best_list=[curr_best])
        return f_idx, ei_pts[int(ei_arg[0]), len(self.f_locs[0]):]


class REVI(CorrelatedOpt):

    def _child_decide_query(self):
        num_eval = int(getattr(self.options, 'num_eval_pts', 100))
        num_cand = int(getattr(self.options, 'num_candidates', 100))
        judge_pts = sample_grid(self.f_locs, self.domains[0], num_eval)
        cand_pts = sample_grid(self.f_locs, self.domains[0], num_cand)
        imp = batched.revi_improvements(self.gps[0], cand_pts, judge_pts, len(self.fcns), num_eval)
        best = first_best(imp)
        return int(np.floor(best / num_cand)), cand_pts[best, len(self.f_locs[0]):]

    @staticmethod
    def get_opt_method_name():
        return 'revi'
