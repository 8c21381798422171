"""Grid / point generators whose LAYOUT is part of the hot-path contract
(/root/reference/src/util/misc_util.py:25-32,104-116,171-173)."""
import numpy as np


_FAST_BOX = [None]


def _fast_box_is_exact():
    """One-time check that ``low + (high - low) * random_sample(size)`` gives the bits (and
    leaves the stream where) ``uniform(low, high, size)`` does in this numpy build: numpy's
    uniform computes exactly that per element, but validates array bounds with several
    reductions per call (~10 us -- per task and per step, it was the largest item of an MTS
    decision's host time)."""
    if _FAST_BOX[0] is None:
        a, b = np.random.RandomState(7), np.random.RandomState(7)
        low, high = np.asarray([-1.5, 0.0, 2.25]), np.asarray([0.25, 1.0, 7.5])
        x = a.uniform(tuple(low), tuple(high), (64, 3))
        y = low + (high - low) * b.random_sample((64, 3))
        _FAST_BOX[0] = bool(np.array_equal(x, y) and a.normal() == b.normal())
    return _FAST_BOX[0]


def uniform_box(lows, highs, num_samps):
    """``np.random.uniform(lows, highs, (num_samps, d))`` from the global stream, same bits."""
    if _fast_box_is_exact():
        low = np.asarray(lows, dtype=np.float64)
        return low + (np.asarray(highs, dtype=np.float64) - low) * \
            np.random.random_sample((num_samps, len(low)))
    return np.random.uniform(lows, highs, (num_samps, len(lows)))


class Box(object):
    """A box domain [[low, high], ...] with its bounds converted once: ``draw(n)`` is
    ``uniform_box`` without the two tuple -> array conversions per call (the continuous
    strategies draw 50 action grids per step from the same box)."""

    def __init__(self, domain):
        self.lows, self.highs = zip(*domain)
        self.low = np.asarray(self.lows, dtype=np.float64)
        self.span = np.asarray(self.highs, dtype=np.float64) - self.low

    def draw(self, num_samps):
        if _fast_box_is_exact():
            return self.low + self.span * np.random.random_sample((num_samps, len(self.low)))
        return np.random.uniform(self.lows, self.highs, (num_samps, len(self.lows)))


def uniform_draw(domain, num_samps):
    """misc_util.py:25-32."""
    lows, highs = zip(*domain)
    return uniform_box(lows, highs, num_samps)


def sample_grid(f_locs, function_domain, num_pts):
    """misc_util.py:104-116 -- rows are TASK-MAJOR: [f_loc_0]*A, [f_loc_1]*A, ...
    hstack U(lows, highs, (T*A, d)); JointMTS reshapes the sample to (T, A)."""
    num_funcs = len(f_locs)
    if num_funcs == 1:  # one context (continuous strategies, 50 calls per step): no tile / hstack
        box = function_domain if isinstance(function_domain, Box) else Box(function_domain)
        c = len(f_locs[0])
        out = np.empty((num_pts, c + len(box.low)))
        out[:, :c] = f_locs[0]
        out[:, c:] = box.draw(num_pts)
        return out
    if isinstance(function_domain, Box):
        lows, highs = function_domain.lows, function_domain.highs
    else:
        lows, highs = zip(*function_domain)
    locs = np.tile(np.asarray(f_locs), num_pts).reshape(-1, len(f_locs[0]))
    f_pts = uniform_box(lows, highs, num_funcs * num_pts)
    return np.hstack([locs, f_pts])


def build_gp_posterior(gp):
    """misc_util.py:171-173."""
    gp.build_posterior()
