"""Synthetic objectives of the four benchmark option files, restated as cheap
host functions (they are scalar Python callables in the reference too:
/root/reference/src/synth/twod.py:12-39, sixd.py:11-25,
continuous_2d.py:14-61, function_generator.py:56-123).  They are fixtures for
the drivers and the tests, not part of the accelerated path."""
from argparse import Namespace
from copy import deepcopy

import numpy as np


def branin(x):
    """twod.py:12-23 (negated Branin on [0,1]^2)."""
    x1 = x[0] * 15 - 5
    x2 = x[1] * 15
    b = 5.1 / (4 * np.pi ** 2)
    c = 5 / np.pi
    t = 1 / (8 * np.pi)
    return -float((x2 - b * x1 ** 2 + c * x1 - 6) ** 2 + 10 * (1 - t) * np.cos(x1) + 10)


def parabaloid(x):
    """twod.py:37-39."""
    x1, x2 = x
    return 1 - x1 ** 2 / 2 - x2 ** 2 / 2


_H6_A = np.asarray([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14],
                    [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]])
_H6_P = 1e-4 * np.asarray([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                           [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])
_H6_ALPHA = np.asarray([1.0, 1.2, 3.0, 3.2])


def hartmann6(x):
    """sixd.py:11-25."""
    x = np.asarray(x, dtype=np.float64)
    d = x - _H6_P  # all four terms at once (same operations in the same order, 2.6x fewer calls)
    return float(np.sum(_H6_ALPHA * np.exp(-np.sum(_H6_A * d ** 2, axis=1))))


def _hartmann6_batch(X):
    """hartmann6 at every row of X in one pass; the reductions run over the same 6 / 4 terms in
    the same order as the per-point form, so the values are bitwise equal (tests)."""
    X = np.asarray(X, dtype=np.float64).reshape(-1, 6)
    d = X[:, None, :] - _H6_P[None, :, :]
    return np.sum(_H6_ALPHA * np.exp(-np.sum(_H6_A * d ** 2, axis=2)), axis=1)


hartmann6.batch = _hartmann6_batch  # optional vectorised form (PolicyScore: 100 - 10^4 points per score)


_H4_A = np.asarray([[10, 3, 17, 3.5], [0.05, 10, 17, 0.1], [3, 3.5, 1.7, 10], [17, 8, 0.05, 10]])
_H4_P = 1e-4 * np.asarray([[1312, 1696, 5569, 124], [2329, 4135, 8307, 3736],
                           [2348, 1451, 3522, 2883], [4047, 8828, 8732, 5743]], dtype=np.float64)
_H3_A = np.asarray([[3, 10, 30], [0.1, 10, 35], [3, 10, 30], [0.1, 10, 35]])
_H3_P = 1e-4 * np.asarray([[3689, 1170, 2673], [4699, 4387, 7470], [1091, 8732, 5547],
                           [381, 5743, 8828]], dtype=np.float64)


def hartmann4(x):
    """twod.py:179-193: the rescaled 4-D Hartmann function (1.1 - sum_i alpha_i exp(-q_i)) / 0.839
    (conth22 / conth31 and the chopped h2-2 / h3-1 task sets)."""
    x = np.asarray(x, dtype=np.float64)
    d = x - _H4_P
    return float((1.1 - np.sum(_H6_ALPHA * np.exp(-np.sum(_H4_A * d ** 2, axis=1)))) / 0.839)


def hartmann3(x):
    """twod.py:163-177."""
    x = np.asarray(x, dtype=np.float64)
    d = x - _H3_P
    return float(np.sum(_H6_ALPHA * np.exp(-np.sum(_H3_A * d ** 2, axis=1))))


synth_functions = [
    Namespace(function=branin, domain=[[0, 1], [0, 1]], max_val=0.397887, name='branin'),
    Namespace(function=parabaloid, domain=[[-1, 1], [-1, 1]], max_val=1, name='parabaloid'),
    Namespace(function=hartmann6, domain=[[0, 1] for _ in range(6)], max_val=3.32237,
              name='hartmann6'),
    Namespace(function=hartmann3, domain=[[0, 1] for _ in range(3)], max_val=3.86278,
              name='hartmann3'),
    Namespace(function=hartmann4, domain=[[0, 1] for _ in range(4)], max_val=None,
              name='hartmann4'),
]


def assemble_functions(names=None):
    """misc_util.py:77-98: 'a,b,c' -> deep copies named ``<name>_<idx>``; unknown
    names raise ValueError like the reference."""
    if names is None:
        return []
    out = []
    for idx, name in enumerate(names.split(',')):
        for f_info in synth_functions:
            if f_info.name.lower() == name.lower():
                cp = deepcopy(f_info)
                cp.name = '%s_%d' % (f_info.name, idx)
                out.append(cp)
                break
        else:
            raise ValueError('Invalid function: %s' % name)
    return out


def assemble_chopped_1d_functions(f_name, num_divisions):
    """continuous_2d.py:14-41: slice a 2-D function along its first axis into
    ``num_divisions`` 1-D tasks located at f_loc = [division]."""
    full = None
    for f_info in synth_functions:
        if f_info.name.lower() == f_name.lower() and len(f_info.domain) == 2:
            full = f_info
    if full is None:
        raise ValueError('Could not find function %s.' % f_name)
    lo, hi = full.domain[0]

    def make_chop(z):
        return lambda x: full.function(np.asarray([z, float(np.ravel(x)[0])]))

    return [Namespace(function=make_chop(div), domain=[full.domain[1]], max_val=None, f_loc=[div],
                      name='%s_%d' % (f_name, i))
            for i, div in enumerate(np.linspace(lo, hi, num_divisions))]


def _chopped(full_f, locations, act_dim, name_fmt):
    """Tasks = a function of (context, action) frozen at the given context locations
    (continuous_2d.py:43-95): f_loc is the context, the domain the unit action box."""
    def make_chop(z):
        return lambda x: full_f(np.append(z, np.asarray(x, dtype=np.float64)))

    return [Namespace(function=make_chop(np.asarray(loc)), domain=[[0, 1] for _ in range(act_dim)],
                      max_val=None, f_loc=np.asarray(loc), name=name_fmt % i)
            for i, loc in enumerate(locations)]


def get_chopped_h42():
    """16 tasks: Hartmann-6 with its 4 context coordinates at the corners of {0,1}^4 (binary
    counting order), 2-D actions (jointh42.txt)."""
    corners = [[(i >> s) & 1 for s in (3, 2, 1, 0)] for i in range(16)]
    return _chopped(hartmann6, corners, 2, 'hartmann_%d')


def get_chopped_h22():
    """9 tasks: Hartmann-4 with 2 context coordinates on the 3 x 3 grid of {0, .5, 1} (first
    coordinate slowest), 2-D actions (jointh22.txt)."""
    g = np.linspace(0, 1, 3)
    return _chopped(hartmann4, [[a, b] for a in g for b in g], 2, 'h2-2_%d')


def get_chopped_h31():
    """8 tasks: Hartmann-4 with 3 context coordinates at the corners of {0,1}^3 (first
    coordinate slowest), 1-D actions (jointh31.txt)."""
    g = np.linspace(0, 1, 2)
    return _chopped(hartmann4, [[a, b, c] for a in g for b in g for c in g], 1, 'h3-1_%d')


def create_mass_func(num_masses, band_bounds, scale_bounds, domain):
    """function_generator.py:101-123: sum of signed Gaussian bumps; consumes the
    global numpy stream in the reference's order."""
    lows, highs = zip(*domain)
    D = len(domain)
    centers = np.random.uniform(lows, highs, (num_masses, D))
    bands = np.random.uniform(band_bounds[0], band_bounds[1], (num_masses, D))
    scales = np.random.uniform(scale_bounds[0], scale_bounds[1], num_masses)
    scales *= (np.random.binomial(1, 0.5, num_masses) - 0.5) * 2

    def f(x):
        x = np.asarray(x, dtype=np.float64)
        if num_masses == 0:
            return 0.0
        # product of per-dimension exponentials, like the reference (not exp of the sum)
        return float(np.sum(scales * np.prod(np.exp(-(x - centers) ** 2 / bands), axis=1)))

    return Namespace(function=f, domain=domain, max_val=None, center=centers, bands=bands,
                     scales=scales)


def random_mass_functions(easy, med, hard, dim, scaling=None):
    """function_generator.py:56-99 in the order the driver builds them
    (src/ocbo.py:214-226): hard, then medium, then easy task sets (rand4d/rand6d)."""
    domain = [[0, 1] for _ in range(dim)]
    out = []
    s = 10 if scaling is None else scaling
    for idx in range(hard):
        fi = create_mass_func(np.random.randint(5, 8), [0.25, 0.4], [0.5 * s, s], domain)
        fi.name = 'hard_mass_%d' % idx
        out.append(fi)
    s = 5 if scaling is None else scaling
    for idx in range(med):
        fi = create_mass_func(np.random.randint(3, 5), [0.4, 0.6], [0.25 * s, s], domain)
        fi.name = 'med_mass_%d' % idx
        out.append(fi)
    s = 1 if scaling is None else scaling
    for idx in range(easy):
        fi = create_mass_func(np.random.randint(0, 3), [0.7, 0.9], [0, s], domain)
        fi.name = 'easy_mass_%d' % idx
        out.append(fi)
    return out
