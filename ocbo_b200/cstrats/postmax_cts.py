"""REVI for continuous contexts -- host mirror of
/root/reference/src/cstrats/postmax_cts.py:22-193: ``cand_size`` candidate
points, a judgement set of judge_ctx_size random contexts x judge_act_size random actions, every
candidate scored by the summed knowledge gradient over the judgement contexts
(batched.revi_improvements: one device pass instead of cand_size * judge_ctx_size Python calls).

Only the set-evaluation form (:139-160) is offered: the thresholded variant (:162-190) is
reached with ``--judge_ctx_thresh / --judge_act_thresh`` only, which no option file sets, and
calls ``gp.eval(..., uncert_form=...)`` with a keyword the reference's own GP wrapper does not
accept."""
from .cts_opt import ContinuousOpt
from .. import batched
from ..options import get_option_specs
from ..strategies.corr_opt import first_best
from ..util import sample_grid, uniform_draw

pm_args = [
    get_option_specs('judge_act_size', False, 50, 'Actions per judgement context.'),
    get_option_specs('judge_ctx_size', False, 50, 'Judgement contexts.'),
    get_option_specs('judge_ctx_thresh', False, None, 'Not supported (see module docstring).'),
    get_option_specs('judge_act_thresh', False, None, 'Not supported (see module docstring).'),
]


class PosteriorMaximization(ContinuousOpt):

    def _child_set_up(self, function, domain, ctx_dim, options):
        self.judge_act_size = int(getattr(options, 'judge_act_size', None) or 50)
        self.judge_ctx_size = int(getattr(options, 'judge_ctx_size', None) or 50)
        if getattr(options, 'judge_ctx_thresh', None) is not None or \
                getattr(options, 'judge_act_thresh', None) is not None:
            raise NotImplementedError('thresholded judgement sets are not supported')

    def _make_judgement_set(self, cand_pt):
        ctxs = uniform_draw(self.ctx_domain, self.judge_ctx_size)
        return sample_grid([list(cx) for cx in ctxs], self.act_domain, self.judge_act_size)


class REVI(PosteriorMaximization):

    @staticmethod
    def get_strat_name():
        return 'revi'

    def _determine_next_query(self):
        cands = self._make_candidate_set()
        judge = self._make_judgement_set(cands[0])
        imp = batched.revi_improvements(self.gp, cands, judge, self.judge_ctx_size, self.judge_act_size)
        return cands[first_best(imp)]


pm_strats = [REVI]
