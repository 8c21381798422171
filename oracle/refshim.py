"""Import shim that lets the UNMODIFIED reference strategy code under
``/root/reference/src`` run on top of the numpy oracle.

TEST INFRASTRUCTURE ONLY; used in the build container by
``tests/golden/make_golden.py`` and the ``refsrc``-marked tests (the GPU box has
no /root/reference).  It injects throw-away modules for the reference's absent
third-party imports into ``sys.modules``:

  dragonfly.gp.euclidean_gp, dragonfly.utils.{general_utils,option_handler},
  dragonfly.exd.{exd_utils,domains}, dragonfly.opt.gpb_acquisitions,
  matplotlib(.pyplot), mpl_toolkits.mplot3d, gpytorch(.models,.kernels,...), tqdm

Nothing from the reference tree is copied; it is imported where it lies.
"""
import os
import sys
import types
from argparse import Namespace

from . import gp_numpy

REFERENCE_SRC = '/root/reference/src'


def reference_available():
    return os.path.isdir(REFERENCE_SRC)


# -- dragonfly.utils.option_handler (call sites: src/ocbo.py:24-71 etc.) --------
def get_option_specs(name, required, default, help_str, **kwargs):
    return {'name': name, 'required': required, 'default': default, 'help': help_str}


def load_options(list_of_options, descr='', cmd_line=False, partial_options=None):
    opts = Namespace()
    for spec in list_of_options:
        setattr(opts, spec['name'], spec['default'])
    if partial_options is not None:
        for k, v in vars(partial_options).items():
            setattr(opts, k, v)
    return opts


EUCLIDEAN_GP_ARGS = [
    get_option_specs('kernel_type', False, 'se', 'Kernel type.'),
    get_option_specs('hp_tune_criterion', False, 'ml', 'ml / post_sampling / a-b.'),
    get_option_specs('hp_samples', False, 0, 'Number of HP samples.'),
    get_option_specs('hp_search_candidates', False, 64, 'Candidates per HP fit round.'),
    get_option_specs('hp_search_rounds', False, 1, 'Search rounds per fit.'),
    get_option_specs('matern_nu', False, 2.5, 'Matern smoothness.'),
    get_option_specs('use_additive_gp', False, False, 'Additive GP (unsupported).'),
    get_option_specs('add_max_group_size', False, 6, 'Unused.'),
    get_option_specs('dim', False, None, 'Input dimension.'),
    get_option_specs('act_dim', False, None, 'Action dimension.'),
    get_option_specs('mean_func_type', False, 'tune', 'Constant mean, tuned.'),
    get_option_specs('noise_var_type', False, 'tune', 'Noise variance, tuned.'),
]


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _unavailable(*a, **k):
    raise NotImplementedError('not part of the MTS hot path; stubbed by oracle.refshim')


def install(gp_cls=None, fitter_cls=None):
    """Install the stubs and put the reference ``src`` on ``sys.path``.
    ``gp_cls`` / ``fitter_cls`` override the numpy core (tests swap in the
    CUDA-backed core to drive the reference strategies with it)."""
    if not reference_available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_SRC)
    sys.dont_write_bytecode = True
    gp_numpy.euclidean_gp_args = EUCLIDEAN_GP_ARGS
    _module('dragonfly')
    _module('dragonfly.gp')
    _module('dragonfly.utils')
    _module('dragonfly.exd')
    _module('dragonfly.opt')
    _module('dragonfly.gp.euclidean_gp',
            EuclideanGP=gp_cls or gp_numpy.EuclideanGP,
            EuclideanGPFitter=fitter_cls or gp_numpy.EuclideanGPFitter,
            euclidean_gp_args=EUCLIDEAN_GP_ARGS)
    _module('dragonfly.utils.general_utils',
            solve_lower_triangular=gp_numpy.solve_lower_triangular)
    _module('dragonfly.utils.option_handler',
            get_option_specs=get_option_specs, load_options=load_options)
    _module('dragonfly.exd.exd_utils', maximise_with_method=_unavailable)
    _module('dragonfly.exd.domains', EuclideanDomain=_unavailable)
    _module('dragonfly.opt.gpb_acquisitions',
            asy_ucb=_unavailable, asy_ei=_unavailable, asy_ttei=_unavailable,
            get_gp_sampler_for_parallel_strategy=_unavailable,
            maximise_acquisition=_unavailable, _get_ucb_beta_th=_unavailable,
            _get_gp_eval_for_parallel_strategy=_unavailable,
            _get_gp_ucb_dim=_unavailable,
            _expected_improvement_for_norm_diff=_unavailable)
    # plotting / alternate engine: never executed on this path
    for name in ('matplotlib', 'matplotlib.pyplot', 'mpl_toolkits', 'mpl_toolkits.mplot3d'):
        if name not in sys.modules:
            _module(name, use=lambda *a, **k: None, cm=None, Axes3D=None)
    if 'gpytorch' not in sys.modules:
        class _Any(object):
            def __init__(self, *a, **k):
                pass
        g = _module('gpytorch')
        g.models = _module('gpytorch.models', ExactGP=_Any)
        g.kernels = _module('gpytorch.kernels', ScaleKernel=_Any, RBFKernel=_Any,
                            LinearKernel=_Any, SpectralMixtureKernel=_Any)
        g.means = _module('gpytorch.means', ConstantMean=_Any)
        g.distributions = _module('gpytorch.distributions', MultivariateNormal=_Any)
        g.likelihoods = _module('gpytorch.likelihoods', GaussianLikelihood=_Any)
        g.mlls = _module('gpytorch.mlls', ExactMarginalLogLikelihood=_Any)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
