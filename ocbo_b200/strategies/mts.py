"""Multi-task Thompson Sampling with independent GPs -- host mirror of
/root/reference/src/strategies/mts.py:17-56.

``decide_next_query`` keeps the reference's semantics (per task: sample the
posterior at [queried points ; max_opt_evals uniform points], gain = best new
sample - best old sample, strict ``>`` keeps the first best task) and its
consumption of the global numpy stream (per task: one ``uniform`` for the
candidates, then one ``normal`` for the sample), but all tasks go through ONE
batched device pass (ocbo_b200.batched.mts_step) instead of a Python loop of
per-task Cholesky factorisations.
"""
import numpy as np

from .multi_opt import MultiOpt
from .. import batched, rng
from ..util import Box


class MTS(MultiOpt):

    def decide_next_query(self):
        queries = self._get_queried_pts()
        boxes = self.__dict__.get('_boxes')
        if boxes is None:  # (the tasks' bounds as arrays, converted once per strategy object)
            boxes = self._boxes = [Box(d) for d in self.domains]
        samp_pts, n_old, normals = [], [], []
        for f_idx, f_name in enumerate(self.f_names):
            f_qs = queries[f_name]
            rand_pts = boxes[f_idx].draw(self.options.max_opt_evals)
            pts = np.concatenate([f_qs.reshape(-1, rand_pts.shape[1]), rand_pts], axis=0)
            samp_pts.append(pts)
            n_old.append(len(f_qs))
            normals.append(rng.normals((len(pts), 1)))
        per_task = batched.mts_step(self.gps, samp_pts, n_old, normals)
        best_f, best_pt, best_gain = None, None, float('-inf')
        for f_idx, (max_prev, new_idx, new_val) in enumerate(per_task):
            gain = new_val - max_prev
            if gain > best_gain:
                best_f, best_gain = f_idx, gain
                best_pt = samp_pts[f_idx][n_old[f_idx] + new_idx]
        return best_f, best_pt

    def _get_queried_pts(self):
        return {fn: np.asarray([q.pt for q in self.query_history[fn]]) for fn in self.f_names}

    @staticmethod
    def get_opt_method_name():
        return 'mts'
