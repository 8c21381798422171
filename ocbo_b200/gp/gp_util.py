"""Engine factory: the reference's plug-in seam for GP engines.

Mirrors /root/reference/src/gp/gp_util.py:16-30 -- ``get_gp_fitter(gp_engine,
x_data, y_data, options)`` dispatches on the engine name; the B200 engine
registers as ``'b200'``.  Unknown engines raise ``ValueError`` exactly like the
reference (:30).  The Dragonfly / GPyTorch CPU engines are not bundled: asking
for them is an error, never a silent fallback.
"""
import numpy as np

from .gp_core import B200GPCore
from .gp_interface import B200GP, B200GPFitter
from ..options import euclidean_gp_args, load_options

ENGINE_NAME = 'b200'


def get_gp_fitter(gp_engine, x_data, y_data, options):
    """reference gp_util.py:16-30."""
    if str(gp_engine).lower() == ENGINE_NAME:
        return B200GPFitter(x_data, y_data, options)
    raise ValueError('GP Engine %s not found.' % gp_engine)


def get_tuned_gp(gp_engine, x_data, y_data, kernel_type='se'):
    """reference gp_util.py:102-117: ML-tuned GP and the options used."""
    options = load_options(euclidean_gp_args, cmd_line=False)
    options.kernel_type = kernel_type
    options.hp_tune_criterion = 'ml'
    options.hp_samples = 0
    fitter = get_gp_fitter(gp_engine, x_data, y_data, options)
    fitter.fit_gp()
    return fitter.get_next_gp(), options


def gp_regression(gp_engine, x_data, y_data, options=None):
    """reference gp_util.py:32-47."""
    if options is None:
        options = load_options(euclidean_gp_args, cmd_line=False)
        options.hp_tune_criterion = 'ml'
        options.kernel_type = 'se'
        options.hp_samples = 0
    gpf = get_gp_fitter(gp_engine, x_data, y_data, options)
    gpf.fit_gp()
    return gpf.get_next_gp()


def get_best_prior(f, domain, kernel_type='se', num_samples=100, gp_engine=ENGINE_NAME):
    """reference gp_util.py:49-71: tune on random samples of ``f`` and return an
    EMPTY-data prior GP carrying the tuned kernel / mean / noise."""
    low_b, high_b = zip(*domain)
    x_data, y_data = [], []
    for _ in range(num_samples):
        pt = np.random.uniform(low_b, high_b)
        y_data.append(f(pt))
        x_data.append(pt)
    tuned, options = get_tuned_gp(gp_engine, x_data, y_data, kernel_type)
    core = tuned.gp_core
    empty = B200GPCore([], [], core.kernel, core.mean_func, core.noise_var, build_posterior=False)
    return B200GP(empty, options)


def get_best_joint_prior(f_infos, kernel_type='se', samps_per=100, gp_engine=ENGINE_NAME):
    """reference gp_util.py:73-100: joint (task location + action) prior."""
    act_low, act_high = zip(*f_infos[0].domain)
    act_dim = len(act_low)
    x_rows, y_data = [], []
    for f_info in f_infos:
        prefixes = np.tile(f_info.f_loc, samps_per).reshape(samps_per, -1)
        rand_acts = np.random.uniform(act_low, act_high, (samps_per, act_dim))
        y_data += [f_info.function(a) for a in rand_acts]
        x_rows.append(np.hstack([prefixes, rand_acts]))
    x_data = np.vstack(x_rows)
    tuned, options = get_tuned_gp(gp_engine, x_data, y_data, kernel_type)
    core = tuned.gp_core
    empty = B200GPCore([], [], core.kernel, core.mean_func, core.noise_var, build_posterior=False)
    return B200GP(empty, options)
