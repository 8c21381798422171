"""B200 GP core: the object the reference reaches as ``gp_core``.

Plays the role Dragonfly's ``EuclideanGP`` plays for the reference
(/root/reference/src/gp/gp_interface.py:67-119, src/gp/gp_util.py:66-70): same
constructor signature, same attributes (``X, Y, kernel, mean_func, noise_var,
L, alpha``) and methods (``eval, draw_samples, add_data_single,
build_posterior``), but every number is computed by ``libocbo_b200.so`` on the
GPU.  Semantics follow SURVEY.md Appendix A.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, engine, rng
from ..engine import DevicePosterior, F64


class _Kernel(object):
    """scale * phi(||(x - x') / dim_bandwidths||); callable like Dragonfly's
    kernel objects (gp_interface.py:113-117) with ``hyperparams`` (:107-110)."""
    kernel_type = None

    def __init__(self, dim, scale, dim_bandwidths):
        self.dim = int(dim)
        bws = np.asarray(dim_bandwidths, dtype=np.float64).ravel()
        if bws.size == 1 and self.dim > 1:
            bws = np.repeat(bws, self.dim)
        if bws.size != self.dim:
            raise ValueError('dim_bandwidths must have %d entries' % self.dim)
        self.hyperparams = {'scale': float(scale), 'dim_bandwidths': bws}

    @property
    def kind(self):
        raise NotImplementedError

    def theta(self, noise_var=0.0, mu0=0.0):
        return np.concatenate([[self.hyperparams['scale'], noise_var, mu0],
                               self.hyperparams['dim_bandwidths']])

    def __call__(self, X1, X2=None):
        """Prior Gram matrix, evaluated on the GPU (ocbo_gram)."""
        X1 = np.asarray(X1, dtype=np.float64).reshape(-1, self.dim)
        X2 = X1 if X2 is None else np.asarray(X2, dtype=np.float64).reshape(-1, self.dim)
        n1, n2 = len(X1), len(X2)
        if n1 == 0 or n2 == 0:
            return np.zeros((n1, n2))
        p1, p2 = engine.pad_to(n1), engine.pad_to(n2)
        A = np.zeros((1, p1, self.dim))
        B = np.zeros((1, p2, self.dim))
        A[0, :n1], B[0, :n2] = X1, X2
        Ad, Bd = engine._h2d(A), engine._h2d(B)
        th = engine._h2d(self.theta()[None, :])
        out = torch.empty((1, p1, p2), dtype=F64, device=Ad.device)
        nv1 = engine._h2d(np.asarray([n1], dtype=np.int32), engine.I32)
        nv2 = engine._h2d(np.asarray([n2], dtype=np.int32), engine.I32)
        engine.gram(self.kind, Ad, nv1, Bd, nv2, th, out, 0)
        return out[0, :n1, :n2].cpu().numpy()


class SEKernel(_Kernel):
    kernel_type = 'se'

    @property
    def kind(self):
        return _lib.KERNEL_SE


class MaternKernel(_Kernel):
    kernel_type = 'matern'

    def __init__(self, dim, nu, scale, dim_bandwidths):
        super(MaternKernel, self).__init__(dim, scale, dim_bandwidths)
        self.nu = float(nu)
        self.hyperparams['nu'] = self.nu
        engine.kernel_kind('matern', self.nu)  # validates nu

    @property
    def kind(self):
        return engine.kernel_kind('matern', self.nu)


def make_kernel(kernel_type, dim, scale, dim_bandwidths, matern_nu=2.5):
    kt = str(kernel_type).lower()
    if kt in ('se', 'default'):
        return SEKernel(dim, scale, dim_bandwidths)
    if kt == 'matern':
        return MaternKernel(dim, matern_nu, scale, dim_bandwidths)
    raise ValueError('Unknown kernel_type %s' % kernel_type)


class ConstMean(object):
    def __init__(self, value):
        self.value = float(value)

    def __call__(self, X):
        return np.full((len(X),), self.value, dtype=np.float64)


class B200GPCore(object):
    """One GP resident on the GPU (a DevicePosterior with batch 1)."""

    def __init__(self, X, Y, kernel, mean_func, noise_var, build_posterior=True):
        self.kernel = kernel
        self.mean_func = mean_func
        self.noise_var = float(noise_var)
        self.dim = kernel.dim
        self._xarr = self._yarr = None
        if isinstance(X, np.ndarray) and X.ndim == 2 and X.dtype == np.float64 and \
                isinstance(Y, np.ndarray) and Y.dtype == np.float64:
            # a fitter hands its own arrays to every GP it draws (one per task and step):
            # rows as views and the arrays themselves as the caches, nothing is copied
            # (``add_data_single`` appends to the lists and drops the caches)
            self.X, self.Y = list(X), Y.tolist()
            if X.shape[1] == self.dim:
                self._xarr, self._yarr = X, Y.ravel()
        else:
            self.X = [np.asarray(x, dtype=np.float64).ravel() for x in X]
            self.Y = [float(y) for y in Y]
        self._post_raw = None
        self.chol_jitter = None
        if build_posterior:
            self.build_posterior()

    # -- the device posterior; its Cholesky status is verified lazily -----------
    @property
    def _post(self):
        p = self._post_raw
        if p is not None and getattr(p, 'unchecked', False):
            self._verify(p)
        return p

    @_post.setter
    def _post(self, value):
        self._post_raw = value

    def _verify(self, post):
        """First host read of a posterior built without a round trip: if the plain Cholesky
        of K + noise I failed, walk stable_cholesky's ladder now.  Returns False in that
        (rare) case, i.e. when whatever was enqueued on the unverified factor must be redone."""
        failed = bool(post.info.cpu().numpy().any())
        if failed:
            post.build()
        post.unchecked = False
        self.chol_jitter = post.jitter[0]
        return not failed

    def has_posterior(self):
        return self._post_raw is not None

    # -- attributes the reference reads -------------------------------------
    @property
    def num_tr_data(self):
        return len(self.Y)

    @property
    def alpha(self):
        if self._post is None:
            return None
        return self._post.ensure_alpha()[0, :self.num_tr_data].cpu().numpy()

    @property
    def L(self):
        if self._post is None:
            return None
        n = self.num_tr_data
        return np.tril(self._post.L[0, :n, :n].cpu().numpy())

    def X_array(self):
        """Training inputs as one (n, dim) array (cached until a point is added)."""
        n = len(self.X)
        if self._xarr is None or self._xarr.shape[0] != n:
            self._xarr = np.asarray(self.X, dtype=np.float64).reshape(n, self.dim)
        return self._xarr

    def Y_array(self):
        """Observations as one (n,) array (cached until a point is added)."""
        n = len(self.Y)
        if self._yarr is None or self._yarr.shape[0] != n:
            self._yarr = np.asarray(self.Y, dtype=np.float64)
        return self._yarr

    def _mu0(self):
        mf = self.mean_func
        if type(mf) is ConstMean:
            return mf.value
        return float(mf(np.zeros((1, self.dim)))[0])

    def theta(self):
        return self.kernel.theta(self.noise_var, self._mu0())

    def __deepcopy__(self, memo):
        # device state is rebuilt lazily (pre-tuned empty GPs are deep-copied:
        # src/strategies/multi_opt.py:180-181)
        from copy import deepcopy
        new = B200GPCore([x.copy() for x in self.X], list(self.Y), deepcopy(self.kernel, memo),
                         deepcopy(self.mean_func, memo), self.noise_var, build_posterior=False)
        return new

    # -- build_posterior / add_data_single (gp_interface.py:81-88) -----------
    def build_posterior(self):
        with engine.staging():
            post = DevicePosterior([self.X_array()], [self.Y],
                                   self.theta()[None, :], self.kernel.kind, self.dim)
        post.build(check=False, need_alpha=False)  # verified at the first host read (_verify)
        self._post_raw = post
        self.chol_jitter = None

    def add_data_single(self, x, y):
        """Append one observation.  Dragonfly rebuilds the posterior here; the
        rebuild is deferred until something reads the posterior (``alpha`` is
        None meanwhile, which is what the reference's wrapper tests, :87)."""
        self.X.append(np.asarray(x, dtype=np.float64).ravel())
        self.Y.append(float(y))
        self._post = None
        self._xarr = self._yarr = None

    def sample_and_pick(self, to_samp, seg_ptr, num_tasks, _normals=None):
        """One joint posterior sample at ``to_samp`` reduced ON THE DEVICE to the
        joint-MTS pick (src/strategies/joint_mts.py:42-57): ``seg_ptr`` holds the
        row offsets of the T old-point segments followed by the T grid segments.
        Returns (task, action index, gain)."""
        Xs = np.asarray(to_samp, dtype=np.float64).reshape(-1, self.dim)
        # the global normal stream is consumed exactly once per call, whatever path is taken
        Zh = rng.normals((len(Xs), 1)) if _normals is None else _normals
        if _normals is None:
            # with trials in flight the joint steps of all trials share one device pass
            from .. import batched, inflight
            return inflight.merge('joint', (self, Xs, np.asarray(seg_ptr, dtype=np.int32),
                                            int(num_tasks), Zh), batched.joint_many)
        return self._sample_and_pick_alone(Xs, seg_ptr, num_tasks, Zh)

    def _sample_and_pick_alone(self, Xs, seg_ptr, num_tasks, Zh):
        if self._post_raw is None:
            self.build_posterior()
        base = self._post_raw  # possibly unverified: checked with the results, below
        to_samp = Xs
        seg_h = np.asarray(seg_ptr, dtype=np.int32)[None, :]
        D, T = self.dim, int(num_tasks)

        def stage():
            Xd, mv_dev, mv, m_pad = engine.upload_points(1, D, [Xs])
            Z = engine.pad_normals([Zh], mv, m_pad, 1)
            return Xd, mv_dev, mv, m_pad, Z, engine._h2d(seg_h, engine.I32)

        def verified():
            return not base.unchecked or self._verify(base)

        def enqueue(post, staged, spec=None):
            Xd, mv_dev, mv, m_pad, Z, seg = staged
            V = post.solve_cross(Xd, mv_dev)
            mean = post.mean_from_v(V, mv_dev)
            Sigma = post.cov(Xd, mv_dev, V, True)
            pend = engine.PendingCholesky(Sigma, mv_dev, spec)

            def tail():
                Fm = engine.sample_from_factor(mean, Sigma, Z, pend.info)
                mx, ag = engine.segment_argmax(Fm, 1, seg)
                return engine.joint_pick(mx, ag, T, T)
            return pend, tail

        if engine.graphs_enabled():
            from ..batched import _PostView, _rungs
            m_pad = engine.pad_to(len(Xs))
            plan = engine.StepPlan.get(('joint', base.kind, D, base.n_pad, m_pad, seg_h.shape[1], T),
                                       8 * m_pad * (D + 1) + 4 * seg_h.shape[1] + 4096)
            view = _PostView(base)
            with plan.staging():
                staged = stage()

            def body(st, ext):
                pend, tail = enqueue(_PostView(base, ext).with_info(), st, engine.GRAPH_SPEC_RUNGS)
                return list(tail()) + [pend.info, pend.rung]

            task, act, gain, info_s, rung = plan.run(staged, view.tensors(), body)
            if not verified():
                return self.sample_and_pick(to_samp, seg_ptr, num_tasks, Zh)
            if not (info_s > 0).any():
                self.last_sample_jitter = _rungs(rung)[0]
                return int(task[0, 0]), int(act[0, 0]), float(gain[0, 0])
            # the ladder needs more than the speculative rungs: eager path below
        with engine.staging(max(4 << 20, 16 * Xs.size + (1 << 16))):
            staged = stage()
        pend, tail = enqueue(base, staged)
        task, act, gain = tail()
        status = pend.status()
        if not verified():
            return self.sample_and_pick(to_samp, seg_ptr, num_tasks, Zh)
        self.last_sample_jitter = pend.resolve(status)[0]
        if status[0][0] > 0:
            task, act, gain = tail()
        return int(task[0, 0].item()), int(act[0, 0].item()), float(gain[0, 0].item())

    def compute_log_marginal_likelihood(self):
        if self._post is None:
            self.build_posterior()
        return float(self._post.lml[0].item())

    # -- eval (gp_interface.py:69-74) ----------------------------------------
    def _posterior_at(self, X_test, want_cov, lower_only):
        if self._post is None:
            self.build_posterior()
        post = self._post
        Xs = np.asarray(X_test, dtype=np.float64).reshape(-1, self.dim)
        Xd, mv_dev, mv, m_pad = post.upload_points([Xs])
        mean = post.mean(Xd, mv_dev)
        if not want_cov:
            return Xd, mv_dev, mv, mean, None, None
        V = post.solve_cross(Xd, mv_dev)
        Sigma = post.cov(Xd, mv_dev, V, lower_only)
        return Xd, mv_dev, mv, mean, V, Sigma

    def _mean_alone(self, Xs):
        _, _, mv, mean, _, _ = self._posterior_at(Xs, False, False)
        return mean[0, :mv[0]].cpu().numpy()

    def eval(self, X_test, uncert_form='none'):
        if uncert_form not in ('none', 'std', 'covar'):
            raise ValueError('uncert_form must be none, std or covar.')
        if uncert_form == 'std':
            # diagonal only: fused scale - |V[:, i]|^2, the m x m covariance is never formed
            Xd, mv_dev, mv, mean, _, _ = self._posterior_at(X_test, False, False)
            post = self._post
            V = post.solve_cross(Xd, mv_dev)
            var = torch.empty_like(mean)
            engine.call('ocbo_post_var_ei', engine._p(V), post.n_pad, V.stride(1), V.stride(0),
                        engine._p(post.nv), None, 0, engine._p(post.theta), post.theta.stride(0), None,
                        mean.shape[1], engine._p(mv_dev), engine._p(var), var.stride(0), None, 0, 1,
                        engine._stream())
            m = mv[0]
            return mean[0, :m].cpu().numpy(), np.sqrt(var[0, :m].cpu().numpy())
        if uncert_form == 'none':
            from .. import batched, inflight  # with trials in flight, mean-only evaluations merge
            Xs = np.asarray(X_test, dtype=np.float64).reshape(-1, self.dim)
            return inflight.merge('mean', (self, Xs), batched.mean_many), None
        _, _, mv, mean, _, Sigma = self._posterior_at(X_test, True, False)
        m = mv[0]
        mu = mean[0, :m].cpu().numpy()
        cov = Sigma[0, :m, :m].cpu().numpy()
        if uncert_form == 'covar':
            return mu, cov
        return mu, np.sqrt(np.diag(cov))

    # -- draw_samples (gp_interface.py:76-79) ---------------------------------
    def draw_samples(self, num_samples, X_test=None, mean_vals=None, covar=None):
        """(S, m) joint samples: mu + L_Sigma z with Dragonfly's jitter ladder."""
        S = int(num_samples)
        if X_test is not None:
            Xd, mv_dev, mv, mean, V, Sigma = self._posterior_at(X_test, True, True)
            post = self._post

            def reassemble(b0, b1):
                post.cov(Xd, mv_dev, V, True, out=Sigma, b0=b0, b1=b1)
        else:
            mean_h = np.asarray(mean_vals, dtype=np.float64).ravel()
            cov_h = np.asarray(covar, dtype=np.float64)
            m = len(mean_h)
            m_pad = engine.pad_to(m)
            Sh = np.eye(m_pad)[None, :, :].copy()
            Sh[0, :m, :m] = cov_h
            Mh = np.zeros((1, m_pad))
            Mh[0, :m] = mean_h
            Sigma, mean = engine._h2d(Sh), engine._h2d(Mh)
            mv = [m]
            mv_dev = engine._h2d(np.asarray(mv, dtype=np.int32), engine.I32)
            def reassemble(b0, b1):
                Sigma.copy_(torch.from_numpy(Sh))  # ladder retry only
        m, m_pad = mv[0], Sigma.shape[1]
        _, info, jit = engine.factor_covariance(None, Sigma, mv_dev, mv, reassemble)
        self.last_sample_jitter = jit[0]
        Z = engine.pad_normals([rng.normals((m, S))], mv, m_pad, S)
        Fm = engine.sample_from_factor(mean, Sigma, Z, info)
        return Fm[0, :m, :S].t().contiguous().cpu().numpy()

    # -- get_pt_relations (gp_interface.py:112-119) ---------------------------
    def cross_covariance(self, pts1, pts2):
        if self._post is None:
            self.build_posterior()
        post = self._post
        Xa, mav, ma, _ = post.upload_points([np.asarray(pts1, dtype=np.float64).reshape(-1, self.dim)])
        Xb, mbv, mb, _ = post.upload_points([np.asarray(pts2, dtype=np.float64).reshape(-1, self.dim)])
        Va, Vb = post.solve_cross(Xa, mav), post.solve_cross(Xb, mbv)
        C = post.cross_cov(Xa, mav, Va, Xb, mbv, Vb)
        return C[0, :ma[0], :mb[0]].cpu().numpy()
