"""Task-agnostic and random baselines of the reference's ``methods`` lists -- host mirrors of
/root/reference/src/strategies/agnostic_opt.py:11-70 (agn-thomp, agn-ei),
joint_agnostic_opt.py:14-88 (ja-thomp, ja-ei) and random_strat.py:10-71 (random, joint-rand).

They are not part of the MTS hot path, but they complete the option files: the task
is drawn uniformly (``np.random.randint``), then a single-task acquisition runs on that task's
GP through the same device entry points as MTS / MEI (one posterior sample, or mean + diagonal
variance).  The global numpy stream is consumed in the reference's order.
"""
import numpy as np

from .joint_opt import JointOpt
from .multi_opt import MultiOpt
from ..util import sample_grid, uniform_draw


def _std_normal_cdf(z):
    from math import erfc, sqrt
    return np.asarray([0.5 * erfc(-v / sqrt(2.0)) for v in np.ravel(z)]).reshape(np.shape(z))


def _std_normal_pdf(z):
    return np.exp(-0.5 * np.square(z)) / np.sqrt(2.0 * np.pi)


def _random_search_ts(gp, domain, num):
    """misc_util.py:353-365: arg-max of one posterior sample at ``num`` uniform points."""
    pts = uniform_draw(domain, num)
    sample = gp.draw_sample(pts)
    return pts[int(np.argmax(sample))]


def _random_search_ei(gp, domain, num, incumbent):
    """misc_util.py:367-385: arg-max of EI at ``num`` uniform points (std from the diagonal;
    eval 'std' on the device never forms the covariance)."""
    pts = uniform_draw(domain, num)
    means, stds = gp.gp_core.eval(pts, 'std')
    z = (means - incumbent) / stds
    eis = stds * (z * _std_normal_cdf(z) + _std_normal_pdf(z))
    return pts[int(np.argmax(eis))]


class _Agnostic(MultiOpt):
    """agnostic_opt.py:11-28: uniformly random task, child picks the point."""

    def decide_next_query(self):
        idx = np.random.randint(len(self.fcns))
        self.gps[idx].build_posterior()
        return idx, self._point_for(idx)

    def _point_for(self, idx):
        raise NotImplementedError('To be implemented in child.')


class AgnEI(_Agnostic):

    def _point_for(self, idx):
        incumbent = self.curr_best[self.f_names[idx]][-1][0]
        return _random_search_ei(self.gps[idx], self.domains[idx], self.options.max_opt_evals, incumbent)

    @staticmethod
    def get_opt_method_name():
        return 'agn-ei'


class AgnThompson(_Agnostic):

    def _point_for(self, idx):
        return _random_search_ts(self.gps[idx], self.domains[idx], self.options.max_opt_evals)

    @staticmethod
    def get_opt_method_name():
        return 'agn-thomp'


class _JointAgnostic(JointOpt):
    """joint_agnostic_opt.py:14-31."""

    def decide_next_query(self):
        idx = np.random.randint(len(self.fcns))
        self.gps[idx].build_posterior()
        return idx, self._point_for(idx)

    def _point_for(self, idx):
        raise NotImplementedError('To be implemented in child.')


class JointAgnThompson(_JointAgnostic):

    def _point_for(self, idx):
        prefix = self.f_locs[idx]
        grid = sample_grid([prefix], self.domains[idx], self.options.max_opt_evals)
        sample = self.gps[idx].draw_sample(grid).ravel()
        return grid[int(np.argmax(sample)), len(prefix):]

    @staticmethod
    def get_opt_method_name():
        return 'ja-thomp'


class JointAgnEI(_JointAgnostic):
    """:62-83.  The reference's score here divides by the VARIANCE and adds the cdf
    (z + Phi(z) + phi(z) with z = (mu - best) / sigma^2): kept as is."""

    def _point_for(self, idx):
        incumbent = self.curr_best[self.f_names[idx]][-1][0]
        grid = sample_grid([self.f_locs[idx]], self.domains[0], self.options.max_opt_evals)
        mu, std = self.gps[idx].gp_core.eval(grid, 'std')
        z = (mu - incumbent) / np.square(std)
        score = z + _std_normal_cdf(z) + _std_normal_pdf(z)
        return grid[int(np.argmax(score)), len(self.f_locs[0]):]

    @staticmethod
    def get_opt_method_name():
        return 'ja-ei'


class _RandomMixin(object):
    """random_strat.py: uniform task and point; the GPs are only drawn and fed when the
    risk-neutral evaluation needs them.  (The per-task refit in ``_update_history`` is NOT
    switched off -- the reference only overrides the plural ``_update_models`` -- so a seeded
    run consumes the global stream exactly like the reference's.)"""

    def decide_next_query(self):
        idx = np.random.randint(len(self.fcns))
        low, high = zip(*self.domains[idx])
        return idx, np.random.uniform(low, high)

    def _draw_next_gps(self):
        if self.risk_neutral:
            return super(_RandomMixin, self)._draw_next_gps()

    def _update_models(self):
        if self.risk_neutral:
            return super(_RandomMixin, self)._update_models()

    def _add_point_to_gp(self, f_idx, pt, val):
        if self.risk_neutral:
            return super(_RandomMixin, self)._add_point_to_gp(f_idx, pt, val)


class RandomOpt(_RandomMixin, MultiOpt):

    @staticmethod
    def get_opt_method_name():
        return 'Random'


class JointRandom(_RandomMixin, JointOpt):

    @staticmethod
    def get_opt_method_name():
        return 'joint-rand'
