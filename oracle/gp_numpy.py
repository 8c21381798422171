"""numpy restatement of the Dragonfly 0.1.4 GP arithmetic OCBO's MTS path uses.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``); PARITY UNPINNED: nothing
in the reference pins these numbers, Dragonfly is not in the image.

Every function cites the reference call site it stands behind
(``/root/reference/src/...``) and, where the arithmetic lives in Dragonfly, the
SURVEY.md Appendix A line it restates ([RECALL] items are assumptions that
this file *defines* for the build).
"""
from __future__ import annotations

from argparse import Namespace
from copy import deepcopy

import numpy as np
from scipy.linalg import solve_triangular

# --- random-number hook ------------------------------------------------------
# The reference draws its normals from the *global* numpy stream inside
# Dragonfly's draw_gaussian_samples.  Tests that inject
# identical normals into the oracle and the CUDA path replace this hook.
NORMAL_SOURCE = None  # callable(size)->ndarray, or None for np.random.normal
TRACE = None  # list collecting every draw (jitter rung, sample) while golden vectors are made
TRACE_EVAL = None  # list collecting (mean, diag covar) of every eval(..., 'covar') likewise


def _normals(size):
    if NORMAL_SOURCE is not None:
        return np.asarray(NORMAL_SOURCE(size), dtype=np.float64)
    return np.random.normal(size=size)


# --- kernels (dist_squared / SEKernel / MaternKernel) ------
def dist_squared(X1, X2):
    """Expanded-form squared distances, clipped at 0 (Appendix A line 1)."""
    n1 = (X1 ** 2).sum(axis=1)[:, None]
    n2 = (X2 ** 2).sum(axis=1)[None, :]
    return np.clip(n1 + n2 - 2.0 * X1.dot(X2.T), 0.0, np.inf)


def _as2d(X, dim):
    X = np.asarray(X, dtype=np.float64)
    if X.size == 0:
        return X.reshape(0, dim)
    if X.ndim == 1:
        X = X.reshape(-1, dim)
    return X


class _StationaryKernel(object):
    """scale * phi(|| (x - x') / dim_bandwidths ||); reached from the reference
    as ``gp_core.kernel(A, B)`` (src/gp/gp_interface.py:114-117) and through
    ``kernel.hyperparams['scale'|'dim_bandwidths']`` (:107-110)."""

    def __init__(self, dim, scale, dim_bandwidths):
        self.dim = int(dim)
        bws = np.asarray(dim_bandwidths, dtype=np.float64).ravel()
        if bws.size == 1 and self.dim > 1:
            bws = np.repeat(bws, self.dim)
        self.hyperparams = {'scale': float(scale), 'dim_bandwidths': bws}

    def _scaled(self, X):
        return _as2d(X, self.dim) / self.hyperparams['dim_bandwidths']

    def _phi(self, d2):
        raise NotImplementedError

    def __call__(self, X1, X2=None):
        X2 = X1 if X2 is None else X2
        d2 = dist_squared(self._scaled(X1), self._scaled(X2))
        return self.hyperparams['scale'] * self._phi(d2)


class SEKernel(_StationaryKernel):
    """scale * exp(-0.5 * dist_squared(X1/bw, X2/bw))  (Appendix A line 2)."""
    kernel_type = 'se'

    def _phi(self, d2):
        return np.exp(-0.5 * d2)


class MaternKernel(_StationaryKernel):
    """Matern nu in {1/2, 3/2, 5/2} on r = sqrt(dist_squared) (Appendix A)."""
    kernel_type = 'matern'

    def __init__(self, dim, nu, scale, dim_bandwidths):
        super(MaternKernel, self).__init__(dim, scale, dim_bandwidths)
        if nu not in (0.5, 1.5, 2.5):
            raise ValueError('Matern nu must be 0.5, 1.5 or 2.5.')
        self.nu = nu
        self.hyperparams['nu'] = nu

    def _phi(self, d2):
        r = np.sqrt(d2)
        if self.nu == 0.5:
            return np.exp(-r)
        if self.nu == 1.5:
            s = np.sqrt(3.0) * r
            return (1.0 + s) * np.exp(-s)
        s = np.sqrt(5.0) * r
        return (1.0 + s + 5.0 * d2 / 3.0) * np.exp(-s)


def make_kernel(kernel_type, dim, scale, dim_bandwidths, matern_nu=2.5):
    kt = str(kernel_type).lower()
    if kt in ('se', 'default'):
        return SEKernel(dim, scale, dim_bandwidths)
    if kt == 'matern':
        return MaternKernel(dim, matern_nu, scale, dim_bandwidths)
    raise ValueError('Unknown kernel_type %s' % kernel_type)


class ConstMean(object):
    """Constant prior mean (Appendix A ``mean_func``)."""

    def __init__(self, value):
        self.value = float(value)

    def __call__(self, X):
        return np.full((len(X),), self.value, dtype=np.float64)


# --- linear algebra -----------------------------------------------------------
def solve_lower_triangular(A, b):
    """dragonfly.utils.general_utils.solve_lower_triangular as imported by the
    reference at src/gp/gp_interface.py:7 and used at :116,:118."""
    return solve_triangular(A, b, lower=True)


def solve_upper_triangular(A, b):
    return solve_triangular(A, b, lower=False)


def stable_cholesky(M, return_jitter=False):
    """Appendix A ``stable_cholesky``: plain Cholesky, then the jitter ladder
    10^p * max(diag M) for p = -11, -10, ..., 4; finally ValueError."""
    M = np.asarray(M, dtype=np.float64)
    if M.size == 0:
        return (M, None) if return_jitter else M
    try:
        L = np.linalg.cholesky(M)
        return (L, None) if return_jitter else L
    except np.linalg.LinAlgError:
        pass
    diag_max = float(np.max(np.diag(M)))
    p = -11
    while p < 5:
        try:
            L = np.linalg.cholesky(M + (10.0 ** p) * diag_max * np.eye(len(M)))
            return (L, p) if return_jitter else L
        except np.linalg.LinAlgError:
            p += 1
    raise ValueError('stable_cholesky: matrix could not be made PD.')


def draw_gaussian_samples(num_samples, mu, K):
    """Appendix A ``draw_gaussian_samples``: L = stable_cholesky(K);
    U ~ N(0,1)^(m x S); return (L U)^T + mu, shape (S, m)."""
    mu = np.asarray(mu, dtype=np.float64).ravel()
    L, p = stable_cholesky(K, return_jitter=True)
    U = _normals((len(mu), num_samples))
    out = L.dot(U).T + mu
    if TRACE is not None:  # golden-vector generation only
        # How well-defined is this sample?  Re-factor one jitter rung higher and see how
        # far the SAME normals land: on numerically semidefinite covariances a tiny lucky
        # pivot amplifies rounding noise (column = residual / sqrt(pivot)), so the
        # reference's own sample can move by O(1) between adjacent rungs.
        md = float(np.max(np.diag(K))) if len(mu) else 0.0
        p_next = (-12 if p is None else p) + 1
        try:
            L2 = np.linalg.cholesky(np.asarray(K) + (10.0 ** p_next) * md * np.eye(len(mu)))
            sens = float(np.max(np.abs(L2.dot(U) - L.dot(U))))
        except np.linalg.LinAlgError:
            sens = float('inf')
        TRACE.append(dict(jitter_p=p, max_diag=md, rung_sens=sens, mean=mu.copy(),
                          sample=out[0].copy(), z_max=float(np.max(np.abs(U)))))
    return out


# --- the GP (Appendix A build_posterior / eval / draw_samples / add_data) -----
class EuclideanGP(object):
    """Stands in for ``dragonfly.gp.euclidean_gp.EuclideanGP`` with exactly the
    surface the reference touches: ctor signature (src/gp/gp_util.py:69-70),
    attributes X, Y, L, alpha, kernel, mean_func, noise_var
    (src/gp/gp_interface.py:87-118), eval / draw_samples / add_data_single /
    build_posterior (:72-88)."""

    def __init__(self, X, Y, kernel, mean_func, noise_var, build_posterior=True):
        self.kernel = kernel
        self.mean_func = mean_func
        self.noise_var = float(noise_var)
        self.dim = kernel.dim
        self.X = [np.asarray(x, dtype=np.float64).ravel() for x in X]
        self.Y = [float(y) for y in Y]
        self.L = None
        self.alpha = None
        self.chol_jitter = None
        if build_posterior:
            self.build_posterior()

    @property
    def num_tr_data(self):
        return len(self.Y)

    def _Xmat(self):
        return _as2d(np.asarray(self.X, dtype=np.float64), self.dim)

    def build_posterior(self):
        X = self._Xmat()
        n = len(X)
        if n == 0:
            self.L = np.zeros((0, 0))
            self.alpha = np.zeros((0,))
            return
        K = self.kernel(X, X) + self.noise_var * np.eye(n)
        self.L, self.chol_jitter = stable_cholesky(K, return_jitter=True)
        resid = np.asarray(self.Y) - self.mean_func(X)
        self.alpha = solve_upper_triangular(self.L.T,
                                            solve_lower_triangular(self.L, resid))

    def eval(self, X_test, uncert_form='none'):
        Xs = _as2d(X_test, self.dim)
        if self.alpha is None:
            self.build_posterior()
        X = self._Xmat()
        mu = self.mean_func(Xs)
        if len(X) > 0:
            K_sx = self.kernel(Xs, X)
            mu = mu + K_sx.dot(self.alpha)
        if uncert_form == 'none':
            return mu, None
        K_ss = self.kernel(Xs, Xs)
        if len(X) > 0:
            V = solve_lower_triangular(self.L, K_sx.T)
            covar = K_ss - V.T.dot(V)
        else:
            covar = K_ss
        if uncert_form == 'covar':
            if TRACE_EVAL is not None:  # golden-vector generation only (EI-family margins)
                TRACE_EVAL.append(dict(mean=mu.copy(), var=np.diag(covar).copy()))
            return mu, covar
        if uncert_form == 'std':
            return mu, np.sqrt(np.diag(covar))
        raise ValueError('uncert_form must be none, std or covar.')

    def draw_samples(self, num_samples, X_test=None, mean_vals=None, covar=None):
        if X_test is not None:
            mean_vals, covar = self.eval(X_test, 'covar')
        return draw_gaussian_samples(num_samples, mean_vals, covar)

    def add_data_single(self, x, y):
        self.X.append(np.asarray(x, dtype=np.float64).ravel())
        self.Y.append(float(y))
        self.build_posterior()

    def compute_log_marginal_likelihood(self):
        """-1/2 (Y-mu0)^T alpha - sum log L_ii - n/2 log 2 pi (Appendix A)."""
        if self.alpha is None:
            self.build_posterior()
        n = self.num_tr_data
        resid = np.asarray(self.Y) - self.mean_func(self._Xmat())
        return float(-0.5 * resid.dot(self.alpha)
                     - np.sum(np.log(np.diag(self.L)))
                     - 0.5 * n * np.log(2.0 * np.pi))


def log_marginal_likelihood(X, Y, kernel_type, theta, matern_nu=2.5):
    """LML of one hyper-parameter vector theta = (scale, noise_var, mu0,
    bw_1..bw_D) -- the evaluator behind ``fit_gp``."""
    X = np.asarray(X, dtype=np.float64)
    D = X.shape[1]
    kern = make_kernel(kernel_type, D, theta[0], theta[3:3 + D], matern_nu)
    gp = EuclideanGP(X, Y, kern, ConstMean(theta[2]), theta[1])
    return gp.compute_log_marginal_likelihood()


# --- hyper-parameter handling (assumptions documented in DESIGN.md) -----------
def hp_bounds(X, Y):
    """Search box for theta in the transformed space
    (log scale, log noise, mu0, log bw_1..D).  SURVEY Appendix A [RECALL]:
    scale in Var(Y)*[0.1,10]; noise in Var(Y)*[5e-3,2e-1]; mean in
    median(Y) +- 3 std(Y); bandwidth_j in spread_j*[1e-2,1e2]."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    y_var = float(np.var(Y))
    if not np.isfinite(y_var) or y_var <= 1e-12:
        y_var = 1.0
    y_std = np.sqrt(y_var)
    spread = X.std(axis=0)
    spread = np.where(spread > 1e-8, spread, 1.0)
    lo = [np.log(0.1 * y_var), np.log(5e-3 * y_var), float(np.median(Y)) - 3.0 * y_std]
    hi = [np.log(10.0 * y_var), np.log(2e-1 * y_var), float(np.median(Y)) + 3.0 * y_std]
    lo += list(np.log(1e-2 * spread))
    hi += list(np.log(1e2 * spread))
    return np.asarray(lo), np.asarray(hi)


def hp_candidates(X, Y, num, uniforms=None):
    """``num`` theta candidates: row 0 is the box centre, the rest are uniform
    in the transformed box.  Consumes ONE ``np.random.uniform(size=(num-1, P))``
    call from the global stream unless ``uniforms`` is given."""
    lo, hi = hp_bounds(X, Y)
    P = len(lo)
    if uniforms is None:
        uniforms = np.random.uniform(size=(max(num - 1, 0), P))
    z = np.vstack([np.full((1, P), 0.5), np.asarray(uniforms).reshape(-1, P)])
    t = lo + z * (hi - lo)
    theta = t.copy()
    theta[:, 0] = np.exp(t[:, 0])
    theta[:, 1] = np.exp(t[:, 1])
    theta[:, 3:] = np.exp(t[:, 3:])
    return theta


def _to_theta(t):
    theta = np.array(t, dtype=np.float64, copy=True)
    theta[:, 0] = np.exp(theta[:, 0])
    theta[:, 1] = np.exp(theta[:, 1])
    theta[:, 3:] = np.exp(theta[:, 3:])
    return theta


def refine_candidates(X, Y, num, rnd, crit, best_theta, uniforms=None):
    """Round ``rnd`` >= 1 of a multi-round search (``hp_search_rounds``): 'ml' zooms on the best
    candidate so far (box half-width (hi - lo) / 2^(rnd+1), clipped), other criteria draw ``num-1``
    more uniform candidates.  Same definition as ocbo_b200.gp.hp_search.refine_candidates."""
    lo, hi = hp_bounds(X, Y)
    P = len(lo)
    if uniforms is None:
        uniforms = np.random.uniform(size=(max(num - 1, 0), P))
    u = np.asarray(uniforms, dtype=np.float64).reshape(-1, P)
    if crit == 'ml':
        centre = np.array(best_theta, dtype=np.float64, copy=True)
        centre[0], centre[1], centre[3:] = np.log(centre[0]), np.log(centre[1]), np.log(centre[3:])
        half = (hi - lo) * 0.5 ** (rnd + 1)
        blo, bhi = np.maximum(lo, centre - half), np.minimum(hi, centre + half)
    else:
        blo, bhi = lo, hi
    return _to_theta(blo + u * (bhi - blo))


euclidean_gp_args = None  # filled in by oracle.dragonfly_stub (option specs)


class EuclideanGPFitter(object):
    """Stands in for ``dragonfly.gp.euclidean_gp.EuclideanGPFitter`` as driven
    from src/gp/gp_interface.py:123-138.

    The real optimiser (DiRect / PDOO / slice sampling) is unknowable here
    (SURVEY section 7 "hard parts"), so the search is *defined* as:
      'ml'            -> arg-max of the LML over ``hp_candidates``;
      'post_sampling' -> sampling-importance-resampling of ``num_samples``
                         candidates with weights exp(LML - max LML);
      'a-b' hyphenated -> the criteria alternate between successive fits.
    Parity for this row is on LML values at given candidates, not on a
    trajectory (BASELINE.json north_star (b))."""

    _fit_counter = 0

    def __init__(self, X, Y, options):
        self.options = options
        self.dim = None
        X = [np.asarray(x, dtype=np.float64).ravel() for x in X]
        if len(X) > 0:
            self.dim = len(X[0])
        elif getattr(options, 'dim', None) not in (None, 'None'):
            self.dim = int(options.dim)
        self.X = X
        self.Y = [float(y) for y in Y]
        self.hp_list = []
        self._next = 0
        self.last_lmls = None
        self.last_candidates = None

    def _criterion(self):
        crits = str(self.options.hp_tune_criterion).split('-')
        crit = crits[EuclideanGPFitter._fit_counter % len(crits)]
        EuclideanGPFitter._fit_counter += 1
        return crit

    def _build(self, theta):
        D = self.dim
        kern = make_kernel(self.options.kernel_type, D, theta[0], theta[3:3 + D],
                           getattr(self.options, 'matern_nu', 2.5))
        return EuclideanGP(self.X, self.Y, kern, ConstMean(theta[2]), theta[1],
                           build_posterior=False)

    def fit_gp_for_gp_bandit(self, num_samples=1):
        num_cand = int(getattr(self.options, 'hp_search_candidates', 64))
        Xm = _as2d(np.asarray(self.X), self.dim)
        rounds = int(getattr(self.options, 'hp_search_rounds', 1) or 1)
        crit = self._criterion()
        lml_of = lambda ths: np.asarray([log_marginal_likelihood(Xm, self.Y, self.options.kernel_type, th,
                                                                 getattr(self.options, 'matern_nu', 2.5))
                                         for th in ths])
        cands = hp_candidates(Xm, self.Y, num_cand)
        lmls = lml_of(cands)
        for rnd in range(1, max(rounds, 1)):
            fin = np.where(np.isfinite(lmls), lmls, -np.inf)
            more = refine_candidates(Xm, self.Y, num_cand, rnd, crit, cands[int(np.argmax(fin))])
            cands = np.concatenate([cands, more], axis=0)
            lmls = np.concatenate([lmls, lml_of(more)])
        self.last_lmls, self.last_candidates = lmls, cands
        finite = np.where(np.isfinite(lmls), lmls, -np.inf)
        if crit == 'ml' or not num_samples or num_samples < 1:
            self.hp_list = [cands[int(np.argmax(finite))]]
        elif crit in ('post_sampling', 'post_mean'):
            w = np.exp(finite - np.max(finite))
            w = w / w.sum()
            u = np.random.uniform(size=int(num_samples))
            idx = np.minimum(np.searchsorted(np.cumsum(w), u), len(w) - 1)
            self.hp_list = [cands[i] for i in idx]
        else:
            raise ValueError('Unknown hp_tune_criterion %s' % crit)
        self._next = 0
        return crit

    def fit_gp(self):
        crit = self.fit_gp_for_gp_bandit(getattr(self.options, 'hp_samples', 0))
        gp = self._build(self.hp_list[0])
        gp.build_posterior()
        return crit, gp, None

    def get_next_gp(self):
        theta = self.hp_list[self._next % len(self.hp_list)]
        self._next += 1
        return ('fitted_gp', theta, self._build(theta))


# --- the reference's wrapper semantics, restated for box-side tests -----------
class OracleGP(object):
    """Same surface as the reference's ``DragonflyGP`` (src/gp/gp_interface.py:
    67-119) over the numpy core; used by tests on the GPU box where
    /root/reference is absent."""

    def __init__(self, gp_core, options=None):
        self.gp_core = gp_core
        self.options = options

    def eval(self, pts, include_covar=False):
        if include_covar:
            return self.gp_core.eval(pts, uncert_form='covar')
        return self.gp_core.eval(pts)

    def draw_sample(self, samp_pts=None, means=None, covar=None):
        return self.gp_core.draw_samples(1, X_test=samp_pts, mean_vals=means,
                                         covar=covar).ravel()

    def add_data_single(self, pt, val):
        self.gp_core.add_data_single(pt, val)

    def build_posterior(self):
        if self.gp_core.alpha is None:
            self.gp_core.build_posterior()

    def get_estimated_noise(self):
        return self.gp_core.noise_var

    def get_kernel_hps(self):
        hps = deepcopy(self.gp_core.kernel.hyperparams)
        hps['lengthscale'] = hps['dim_bandwidths']
        hps.pop('dim_bandwidths', None)
        return hps

    def get_pt_relations(self, pt_set1, pt_set2):
        g = self.gp_core
        prior = g.kernel(pt_set1, pt_set2)
        X = g._Xmat()
        V1 = solve_lower_triangular(g.L, g.kernel(pt_set1, X).T)
        V2 = solve_lower_triangular(g.L, g.kernel(pt_set2, X).T)
        return prior - V1.T.dot(V2)


class OracleGPFitter(object):
    """Same surface as ``DragonflyGPFitter`` (src/gp/gp_interface.py:121-138)."""

    def __init__(self, x_data, y_data, options):
        self.x_data, self.y_data, self.options = x_data, y_data, options
        self.gpf_core = EuclideanGPFitter(x_data, y_data, options)

    def fit_gp(self):
        self.gpf_core.fit_gp_for_gp_bandit(num_samples=self.options.hp_samples)

    def get_next_gp(self):
        nxt = self.gpf_core.get_next_gp()[2]
        nxt.build_posterior()
        return OracleGP(nxt, self.options)


def make_gp(X, Y, kernel_type, scale, bandwidths, noise_var, mu0, matern_nu=2.5,
            build=True):
    """Convenience: a built OracleGP at fixed hyper-parameters."""
    X = np.asarray(X, dtype=np.float64)
    D = X.shape[1] if X.ndim == 2 and X.size else len(np.atleast_1d(bandwidths))
    kern = make_kernel(kernel_type, D, scale, bandwidths, matern_nu)
    core = EuclideanGP(list(X), list(np.asarray(Y, dtype=np.float64).ravel()), kern,
                       ConstMean(mu0), noise_var, build_posterior=build)
    return OracleGP(core, Namespace(use_additive_gp=False))
