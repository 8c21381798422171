"""Grid / point generators whose LAYOUT is part of the hot-path contract
(/root/reference/src/util/misc_util.py:25-32,104-116,171-173)."""
import numpy as np


def uniform_draw(domain, num_samps):
    """misc_util.py:25-32."""
    lows, highs = zip(*domain)
    return np.random.uniform(lows, highs, (num_samps, len(lows)))


def sample_grid(f_locs, function_domain, num_pts):
    """misc_util.py:104-116 -- rows are TASK-MAJOR: [f_loc_0]*A, [f_loc_1]*A, ...
    hstack U(lows, highs, (T*A, d)); JointMTS reshapes the sample to (T, A)."""
    lows, highs = zip(*function_domain)
    num_funcs = len(f_locs)
    locs = np.tile(np.asarray(f_locs), num_pts).reshape(-1, len(f_locs[0]))
    f_pts = np.random.uniform(lows, highs, (num_funcs * num_pts, len(lows)))
    return np.hstack([locs, f_pts])


def build_gp_posterior(gp):
    """misc_util.py:171-173."""
    gp.build_posterior()
