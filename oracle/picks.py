"""numpy restatement of the reference's sample->pick reductions.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  These follow the
in-tree strategy code line by line and are pinned by running the reference's
own unmodified classes (``tests/golden/make_golden.py``).
"""
import numpy as np


def sample_grid(f_locs, function_domain, num_pts, uniforms=None):
    """src/util/misc_util.py:104-116 -- task-major candidate grid
    ``[f_loc_0]*A, [f_loc_1]*A, ...`` hstack U(lows, highs, (T*A, d))."""
    lows, highs = zip(*function_domain)
    num_funcs = len(f_locs)
    locs = np.tile(np.asarray(f_locs), num_pts).reshape(-1, len(f_locs[0]))
    if uniforms is None:
        f_pts = np.random.uniform(lows, highs, (num_funcs * num_pts, len(lows)))
    else:
        f_pts = np.asarray(lows) + np.asarray(uniforms) * (np.asarray(highs) - np.asarray(lows))
    return np.hstack([locs, f_pts])


def mts_pick(samples_per_task, n_old_per_task):
    """src/strategies/mts.py:37-43.  ``samples_per_task[f]`` is the posterior
    sample at [queried pts ; new pts] of task f.  Returns (task, index into the
    new points, gain); strict ``>`` keeps the first best task."""
    best_f, best_idx, best_gain = None, None, float('-inf')
    for f_idx, (vals, n_old) in enumerate(zip(samples_per_task, n_old_per_task)):
        vals = np.asarray(vals).ravel()
        max_prev = np.max(vals[:n_old])
        best_new = int(np.argmax(vals[n_old:])) + n_old
        gain = vals[best_new] - max_prev
        if gain > best_gain:
            best_f, best_idx, best_gain = f_idx, best_new - n_old, gain
    return best_f, best_idx, best_gain


def joint_mts_pick(sampled, num_qs, num_tasks):
    """src/strategies/joint_mts.py:43-53.  ``sampled`` = one joint sample over
    [past pts (task by task) ; T*A grid (task-major)].  Returns (task, action
    index, gain)."""
    sampled = np.asarray(sampled).ravel()
    total_qs = int(sum(num_qs))
    old_pts, new_pts = sampled[:total_qs], sampled[total_qs:]
    new_pts = new_pts.reshape(num_tasks, -1)
    divs = [0] + [sum(num_qs[:idx]) for idx in range(1, len(num_qs) + 1)]
    old_bests = np.asarray([np.max(old_pts[divs[i]:divs[i + 1]])
                            for i in range(len(divs) - 1)]).reshape(-1, 1)
    gains = new_pts - old_bests
    best_gains = np.max(gains, axis=1).ravel()
    best_ctx = int(np.argmax(best_gains))
    best_idx = int(np.argmax(gains[best_ctx]))
    return best_ctx, best_idx, float(gains[best_ctx, best_idx])


def grid_only_pick(sampled, num_tasks):
    """Sweep variant (SURVEY 8d "sweep"): candidates are the grid only, the
    per-task baseline is 0, so gain = max_a sample[t, a]; argmax over tasks
    then actions with numpy first-index ties (joint_mts.py:50-53 with
    old_bests = 0)."""
    g = np.asarray(sampled).reshape(num_tasks, -1)
    best_gains = np.max(g, axis=1)
    t = int(np.argmax(best_gains))
    a = int(np.argmax(g[t]))
    return t, a, float(g[t, a])


def cmts_gain(sample, means):
    """src/cstrats/profile_cts.py:101-104 (cmts)."""
    sample = np.asarray(sample).ravel()
    best_post = int(np.argmax(means))
    return int(np.argmax(sample)), float(np.max(sample) - sample[best_post])


def cmts_pm_gain(sample, means, y_data):
    """src/cstrats/profile_cts.py:82-85 (cmts-pm)."""
    sample = np.asarray(sample).ravel()
    best_post = np.min([np.max(means), np.max(y_data)])
    return int(np.argmax(sample)), float(np.max(sample) - best_post)


def profile_pick(gains):
    """src/cstrats/profile_cts.py:30-35: strict ``>`` across contexts."""
    best, best_imp = None, float('-inf')
    for i, g in enumerate(gains):
        if g > best_imp:
            best, best_imp = i, g
    return best, best_imp
