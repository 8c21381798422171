"""Context-agnostic and random baselines of the continuous drivers -- host mirrors of
/root/reference/src/cstrats/rand_cts.py:9-29 (rand) and agn_cts.py:18-76 (ts, ei): one random
context, then the action by posterior sample / EI-like score over ``agn_evals`` random actions.
Not on the MTS hot path; they complete the option files' method lists and use the
same device entry points (eval 'std', draw_sample)."""
import numpy as np

from .cts_opt import ContinuousOpt
from ..options import get_option_specs
from ..strategies.baselines import _std_normal_cdf, _std_normal_pdf
from ..util import sample_grid

agn_args = [
    get_option_specs('agn_evals', False, 100, 'Random actions per context to determine the max.'),
]


class RandOpt(ContinuousOpt):
    """Uniform context and action; no GP is fitted or drawn."""

    @staticmethod
    def get_strat_name():
        return 'rand'

    def _determine_next_query(self):
        ctx = self._get_ctx_candidates(1)[0]
        lows, highs = zip(*self.act_domain)
        return np.append(ctx, np.random.uniform(lows, highs))

    def _make_new_gp_fitter(self):
        return None

    def _draw_next_gp(self):
        return None


class _Agnostic(ContinuousOpt):

    def _child_set_up(self, function, domain, ctx_dim, options):
        self.agn_evals = getattr(options, 'agn_evals', 100)

    def _determine_next_query(self):
        ctx = self._get_ctx_candidates(1)[0]
        return self._best_action(ctx)


class AgnEI(_Agnostic):
    """:37-56.  Score z + Phi(z) + phi(z) with z = (mean - max mean) / VARIANCE, as the
    reference computes it (kept, not "fixed")."""

    @staticmethod
    def get_strat_name():
        return 'ei'

    def _best_action(self, ctx):
        act_set = sample_grid([list(ctx)], self.act_domain, self.agn_evals)
        means, stds = self.gp.gp_core.eval(act_set, 'std')
        z = (means - np.max(means)) / np.square(stds)
        return act_set[int(np.argmax(z + _std_normal_cdf(z) + _std_normal_pdf(z)))]


class AgnTS(_Agnostic):

    @staticmethod
    def get_strat_name():
        return 'ts'

    def _best_action(self, ctx):
        act_set = sample_grid([list(ctx)], self.act_domain, self.agn_evals)
        return act_set[int(np.argmax(self.gp.draw_sample(act_set).ravel()))]
