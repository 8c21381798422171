"""Continuous-context Thompson sampling (CMTS, CMTS-PM) -- host mirror of
/root/reference/src/cstrats/profile_cts.py:20-111.

The reference loops over ``num_profiles`` contexts, and for each evaluates the
posterior mean + 100 x 100 covariance on ``profile_evals`` random actions,
draws one sample and scores the context (:26-37, :76-105).  Here all contexts
are one strided batch sharing the GP's Cholesky factor (stride 0): one batched
cross-covariance solve, one batched block covariance, one batched Cholesky,
one batched sample, one segmented max/argmax.  The global numpy stream is
consumed in the reference's order (contexts; then per context: actions, normals).
"""
from argparse import Namespace

import numpy as np

from .cts_opt import ContinuousOpt
from .. import batched, rng
from ..options import get_option_specs
from ..util import Box, sample_grid

prof_args = [
    get_option_specs('num_profiles', False, 50, 'Contexts considered per suggestion.'),
    get_option_specs('profile_evals', False, 100, 'Random actions per context.'),
]


class ProfileOpt(ContinuousOpt):

    def _child_set_up(self, function, domain, ctx_dim, options):
        self.num_profiles = options.num_profiles
        self.profile_evals = options.profile_evals
        self._act_box = Box(self.act_domain)  # (bounds converted once: 50 grids per step)

    def _determine_next_query(self):
        ctxs = self._get_ctx_candidates(self.num_profiles)
        act_sets, normals = [], []
        for ctx in ctxs:
            act_set = sample_grid([ctx], self._act_box, self.profile_evals)
            act_sets.append(act_set)
            normals.append(rng.normals((len(act_set), 1)))
        s_max, s_arg, m_max, m_arg, s_at_m = batched.profile_step(self.gp, act_sets, normals)
        best_pt, best_imp = None, float('-inf')
        for i in range(len(ctxs)):
            imp = self._improvement(s_max[i], m_max[i], s_at_m[i])
            if imp > best_imp:
                best_pt, best_imp = act_sets[i][int(s_arg[i])], imp
        return best_pt

    def _improvement(self, sample_max, mean_max, sample_at_mean_argmax):
        raise NotImplementedError('Abstract Method')


class ProfileEI(ProfileOpt):
    """PEI (:47-67): per context, EI against min(max posterior mean, best reward seen);
    no sampling -- fused mean + diagonal variance + EI on the device."""

    @staticmethod
    def get_strat_name():
        return 'pei'

    def _determine_next_query(self):
        ctxs = self._get_ctx_candidates(self.num_profiles)
        act_sets = [sample_grid([ctx], self._act_box, self.profile_evals) for ctx in ctxs]
        ei_max, ei_arg = batched.shared_ei_step(self.gp, act_sets, cap=np.max(self.y_data))
        best_pt, best_imp = None, float('-inf')
        for i in range(len(ctxs)):
            if ei_max[i] > best_imp:
                best_pt, best_imp = act_sets[i][int(ei_arg[i])], ei_max[i]
        return best_pt


class CMTSPM(ProfileOpt):
    """gain = max(sample) - min(max(means), max(y_data))   (:76-86)."""

    @staticmethod
    def get_strat_name():
        return 'cmts-pm'

    def _improvement(self, sample_max, mean_max, sample_at_mean_argmax):
        return sample_max - np.min([mean_max, np.max(self.y_data)])


class ContinuousMultiTaskTS(ProfileOpt):
    """gain = max(sample) - sample[argmax(means)]   (:95-105)."""

    @staticmethod
    def get_strat_name():
        return 'cmts'

    def _improvement(self, sample_max, mean_max, sample_at_mean_argmax):
        return sample_max - sample_at_mean_argmax


prof_strats = [Namespace(impl=ProfileEI, name=ProfileEI.get_strat_name()),
               Namespace(impl=CMTSPM, name=CMTSPM.get_strat_name()),
               Namespace(impl=ContinuousMultiTaskTS, name=ContinuousMultiTaskTS.get_strat_name())]
