"""Host-side engine: padded, batched GP posteriors on one B200.

PyTorch is plumbing only (device memory, streams, pinned staging); every
floating-point operation of the path runs in ``libocbo_b200.so``.  There is no
CPU fallback: constructing anything here without CUDA raises.

Layout in HBM (all fp64, row-major, one slab per batched quantity):
  X      [B, n_pad, D]        training inputs, rows >= nv[b] are padding
  y      [B, n_pad]
  theta  [B, 3 + D]           scale, noise_var, mu0, bandwidths
  L      [B, n_pad, n_pad]    lower Cholesky factor of K + noise I (+ jitter)
  invd   [B, n_pad/128, 128, 128]  inverted diagonal blocks of L
  alpha  [B, n_pad]
with n_pad a multiple of 128.  Padding is exact (see include/ocbo_b200.h).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import TILE, call

F64 = torch.float64
I32 = torch.int32

KIND_OF = {'se': _lib.KERNEL_SE, 'default': _lib.KERNEL_SE,
           ('matern', 0.5): _lib.KERNEL_MATERN12, ('matern', 1.5): _lib.KERNEL_MATERN32,
           ('matern', 2.5): _lib.KERNEL_MATERN52}


def kernel_kind(kernel_type, matern_nu=2.5):
    kt = str(kernel_type).lower()
    if kt in ('se', 'default'):
        return _lib.KERNEL_SE
    if kt == 'matern':
        try:
            return KIND_OF[('matern', float(matern_nu))]
        except KeyError:
            raise ValueError('Matern nu must be 0.5, 1.5 or 2.5.')
    raise ValueError('Unknown kernel_type %s' % kernel_type)


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('ocbo_b200 needs a CUDA device (B200, sm_100a); there is no CPU path.')


def device():
    require_cuda()
    return torch.device('cuda', torch.cuda.current_device())


def pad_to(n, mult=TILE):
    return max(mult, ((int(n) + mult - 1) // mult) * mult)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


try:
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # older torch: the slow public path
    _raw_stream = None


def _stream():
    """cudaStream_t of torch's current stream (the raw getter costs ~0.3 us; the public
    ``current_stream()`` object ~13 us, which is a visible share of a small MTS step)."""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class Staging(object):
    """One host->device copy for all the small operands of a step.

    Inside ``with staging():`` every ``_h2d`` bump-allocates from a persistent pinned host
    arena and returns a view of ONE device arena at the same offset; leaving the block
    issues a single asynchronous copy of the used prefix.  The uploads of a small MTS step
    (training sets, hyper-parameters, candidates, normals, segment offsets: 8 arrays) cost
    ~35 us each as separate pageable copies -- more than the step's kernels.  Nothing may be
    launched on the staged tensors before the block is left."""
    ALIGN = 256
    _host = None     # persistent pinned arena (grown on demand)
    _done = None     # event: the last copy out of the arena has finished

    def __init__(self, capacity=4 << 20):
        self.capacity = int(capacity)
        self.used = 0
        self.dev = None

    def __enter__(self):
        require_cuda()
        cls = Staging
        if cls._host is None or cls._host.numel() < self.capacity:
            cls._host = torch.empty(self.capacity, dtype=torch.uint8).pin_memory()
        elif cls._done is not None:
            cls._done.synchronize()  # the previous step's copy must have left the arena
        self.capacity = cls._host.numel()
        self.host = cls._host
        self.host_np = cls._host.numpy()
        self.dev = torch.empty(self.capacity, dtype=torch.uint8, device=device())
        _ACTIVE.append(self)
        return self

    def alloc(self, shape, dtype):
        """Reserve a region: (numpy view of the pinned host side, device view) or None if it
        does not fit.  The host side is a numpy view of the pinned arena (a slice assignment,
        ~1 us; the torch spelling of the same copy is four tensor ops, ~8 us -- a dozen
        operands per step made that a fifth of a small step's host time)."""
        npdt = _NP_OF[dtype]
        nb = int(np.prod(shape)) * np.dtype(npdt).itemsize
        off = self.used
        if off + nb > self.capacity:
            return None
        self.used = off + ((nb + self.ALIGN - 1) // self.ALIGN) * self.ALIGN
        return (self.host_np[off:off + nb].view(npdt).reshape(shape),
                self._dev_view(off, nb, tuple(shape), dtype))

    def take(self, arr, dtype):
        """Stage contiguous ndarray ``arr``; returns its device view (torch ``dtype``) or None."""
        r = self.alloc(arr.shape, dtype)
        if r is None:
            return None
        if arr.size:
            r[0][...] = arr
        return r[1]

    def _dev_view(self, off, nb, shape, dtype):
        return self.dev[off:off + nb].view(dtype).view(shape)

    def __exit__(self, *exc):
        _ACTIVE.pop()
        if self.used and exc[0] is None:
            self.dev[:self.used].copy_(Staging._host[:self.used], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            Staging._done = ev
        return False


_ACTIVE = []


def staging(capacity=4 << 20):
    return Staging(capacity)


class _PlanStaging(Staging):
    """Staging into the persistent arenas of a StepPlan: no copy on exit (the copy is the
    first node of the plan's graph) and a deterministic layout (same calls, same offsets)."""

    def __init__(self, plan):
        self.plan, self.capacity, self.used = plan, plan.capacity, 0
        self.host, self.dev, self.host_np = plan.host, plan.dev, plan.host_np

    def __enter__(self):
        _ACTIVE.append(self)
        return self

    def alloc(self, shape, dtype):
        r = Staging.alloc(self, shape, dtype)
        if r is None:
            raise RuntimeError('StepPlan arena too small (%d bytes)' % self.capacity)
        return r

    def _dev_view(self, off, nb, shape, dtype):
        # same calls, same offsets: the device views of a plan are made once
        key = (off, shape, dtype)
        d = self.plan.views.get(key)
        if d is None:
            d = self.plan.views[key] = Staging._dev_view(self, off, nb, shape, dtype)
        return d

    def __exit__(self, *exc):
        _ACTIVE.pop()
        self.plan.used = self.used
        return False


def graphs_enabled():
    import os
    return os.environ.get('OCBO_GRAPHS', '1') != '0'


class StepPlan(object):
    """One decision step at fixed padded shapes as a CUDA graph.

    A small MTS step is ~30 kernels of 5-40 us: launched one by one from Python its wall
    time is the host's (0.9 ms for set2d against 0.38 ms of kernels).  The padded shapes
    (batch, n_pad, m_pad) -- the only launch parameters -- stay fixed for ~100 BO steps, and
    everything that changes (data, valid sizes, normals, segment offsets) lives in device
    memory, so the whole step is captured once per shape key and replayed:
        fill pinned arena -> [graph: H2D arena, kernels incl. the speculative ladder rungs,
        D2H results] -> one stream synchronise.
    ``ext`` tensors (a posterior built elsewhere) are copied into plan-owned buffers before
    the replay.  Results are numpy views of pinned memory, valid until the next run."""
    MAX_PLANS = 24
    _plans = {}
    _order = []
    last_run = None  # the plan replayed most recently (measurement hooks: device_ms_per_replay)

    @classmethod
    def get(cls, key, capacity):
        plan = cls._plans.get(key)
        if plan is None:
            if len(cls._order) >= cls.MAX_PLANS:
                cls._plans.pop(cls._order.pop(0), None)
            plan = cls._plans[key] = StepPlan(capacity)
            cls._order.append(key)
        return plan

    @classmethod
    def clear(cls):
        cls._plans.clear()
        del cls._order[:]

    def __init__(self, capacity):
        require_cuda()
        self.capacity = ((int(capacity) + 4095) // 4096) * 4096
        self.host = torch.empty(self.capacity, dtype=torch.uint8).pin_memory()
        self.dev = torch.empty(self.capacity, dtype=torch.uint8, device=device())
        self.host_np = self.host.numpy()
        self.views = {}
        self.used = 0
        self.graph = None
        self.replays = 0

    def staging(self):
        return _PlanStaging(self)

    def run(self, staged, ext, body, wait=True, stream=None):
        """body(staged, ext) -> list of device tensors; returns them as numpy arrays.  With
        ``wait=False`` the replay is only enqueued (the caller stages its next chunk while this
        one runs) and ``results()`` synchronises later; ``stream`` replays the graph on another
        stream than the current one (chunks of one pass overlap on the device)."""
        if self.graph is None:
            self.staged, self.used0 = staged, self.used
            self.ext = [t.clone() for t in ext]
            self.dev[:self.used].copy_(self.host[:self.used], non_blocking=True)
            outs = body(self.staged, self.ext)  # eager warm-up: lazy kernel attributes, shapes
            torch.cuda.synchronize()
            self.outs_host = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
            g = torch.cuda.CUDAGraph()
            launches0 = int(_lib.LIB.ocbo_launch_count())
            with torch.cuda.graph(g):
                self.dev[:self.used].copy_(self.host[:self.used], non_blocking=True)
                outs = body(self.staged, self.ext)
                for h, o in zip(self.outs_host, outs):
                    h.copy_(o, non_blocking=True)
            self.graph = g
            self.kernels = int(_lib.LIB.ocbo_launch_count()) - launches0  # library kernels in the graph
        else:
            if self.used != self.used0:
                raise RuntimeError('StepPlan: staged layout changed under a fixed shape key')
            for own, t in zip(self.ext, ext):
                own.copy_(t, non_blocking=True)
        self._stream = stream if stream is not None else torch.cuda.current_stream()
        if stream is not None:
            if ext:  # (the copies into the plan's own buffers were enqueued on the current stream)
                stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(stream):
                self.graph.replay()
        else:
            self.graph.replay()
        self.replays += 1
        StepPlan.last_run = self
        return self.results() if wait else None

    def device_ms_per_replay(self, reps=20):
        """Device time of one replay of this plan's graph (H2D of the arena, every kernel, D2H of
        the results), replays back to back with CUDA events around them -- no host work between."""
        if self.graph is None:
            return None
        st = torch.cuda.current_stream()
        self.graph.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            self.graph.replay()
        e1.record(st)
        st.synchronize()
        return e0.elapsed_time(e1) / reps

    def results(self):
        self._stream.synchronize()
        return [h.numpy() for h in self.outs_host]


_NP_OF = {F64: np.float64, I32: np.int32}


def _h2d(arr, dtype=F64):
    """Host ndarray -> device tensor (through the active ``staging()`` arena if there is one)."""
    arr = np.ascontiguousarray(arr, dtype=_NP_OF[dtype])
    if _ACTIVE:
        d = _ACTIVE[-1].take(arr, dtype)
        if d is not None:
            return d
    t = torch.from_numpy(arr)
    if t.numel() * t.element_size() < (1 << 20):
        # small operands (the option-file shapes): a pageable copy costs ~10 us, pinning
        # a fresh buffer costs a cudaHostAlloc (~100 us) per call
        return t.to(device())
    return t.pin_memory().to(device(), non_blocking=True)


LAUNCHES = [0]  # kernels launched by this process (bench.py reports it)


def _count(k):
    LAUNCHES[0] += k


# ---------------------------------------------------------------------------
# thin wrappers over the C ABI on torch tensors
# ---------------------------------------------------------------------------
def gram(kind, X1, nv1, X2, nv2, theta, out, flags):
    """out[b] = K(X1[b], X2[b]); X: [B, n_pad, D]; out: [B, n1_pad, ld]."""
    B, n1, D = X1.shape
    n2 = X2.shape[1]
    call('ocbo_gram', kind, D, _p(X1), n1, X1.stride(0), _p(nv1), _p(X2), n2, X2.stride(0), _p(nv2),
         _p(theta), theta.stride(0), _p(out), out.stride(1), out.stride(0), B, flags, _stream())
    _count(1)


def potrf(A, invd, info, nv=None, ws=None):
    """Batched in-place Cholesky; ``nv`` (device int32 [B]) marks identity padding whose
    elimination steps the diagonal-block kernel may skip; ``ws`` (uint8 device tensor of
    ocbo_potrf_ws_bytes, full-size problems only): trailing updates on the int8 tensor cores."""
    B, n, _ = A.shape
    if ws is not None and nv is None:
        call('ocbo_potrf_ws', _p(A), n, A.stride(1), A.stride(0), _p(invd), _p(info), B, _p(ws), ws.numel(),
             _stream())
    else:
        call('ocbo_potrf_nv', _p(A), n, A.stride(1), A.stride(0), _p(nv), _p(invd), _p(info), B,
             _stream())
    _count(3 * (n // TILE) - 2)


def trsm_lower(L, invd, Bm, info):
    B, n, _ = L.shape
    call('ocbo_trsm_lower', _p(L), n, L.stride(1), L.stride(0), _p(invd), _p(Bm), Bm.shape[2],
         Bm.stride(1), Bm.stride(0), _p(info), B, _stream())
    _count(2 * (n // TILE) - 1)


def chol_with_ladder(assemble, A, invd, info, nv_dev, nv_host):
    """Dragonfly's stable_cholesky on a batch: factor once;
    every problem whose plain factorisation fails is re-assembled and retried
    alone with jitter 10^p * max(diag) for p = -11 .. 4, then ValueError.
    ``assemble(b0, b1)`` must (re)write A[b0:b1].  Returns the jitter exponent
    used per problem (None = none needed)."""
    B = A.shape[0]
    info.zero_()
    potrf(A, invd, info, nv_dev)
    status = info.cpu().numpy()
    jit = [None] * B
    if not status.any():
        return jit
    dmax = torch.empty(1, dtype=F64, device=A.device)
    val = torch.empty(1, dtype=F64, device=A.device)
    for b in np.nonzero(status)[0]:
        b = int(b)
        Ab, ib, inf_b = A[b:b + 1], invd[b:b + 1], info[b:b + 1]
        nvb = nv_dev[b:b + 1] if nv_dev is not None else None
        n = A.shape[1]
        done = False
        for p in range(-11, 5):
            assemble(b, b + 1)
            call('ocbo_diag_max', _p(Ab), n, Ab.stride(1), Ab.stride(0), _p(nvb), _p(dmax), 1, _stream())
            val.copy_(dmax * (10.0 ** p))
            call('ocbo_add_diag', _p(Ab), n, Ab.stride(1), Ab.stride(0), _p(nvb), _p(val), 1, _stream())
            _count(2)
            inf_b.zero_()
            potrf(Ab, ib, inf_b, nvb)
            if int(inf_b.item()) == 0:
                jit[b] = p
                done = True
                break
        if not done:
            raise ValueError('stable_cholesky: matrix could not be made PD.')
    return jit


def _load_hostpack():
    """The CPython staging helper built by ocbo_b200/build.py (csrc/hostpack.c), or None."""
    import importlib.util
    from .build import hostpack_path
    path = hostpack_path()
    if not os.path.exists(path):
        return None
    try:
        spec = importlib.util.spec_from_file_location('ocbo_b200.lib._hostpack', path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception:  # staging falls back to the numpy loop: data movement only
        return None


_HOSTPACK = _load_hostpack() if os.environ.get('OCBO_HOSTPACK', '1') != '0' else None


def pack_rows(arrs, width, n_pad=None, stage=False):
    """Ragged list of arrays holding n_b rows of ``width`` values -> zero-padded
    [B, n_pad, width], the row counts, n_pad.  With ``stage`` the block is written straight into
    the active staging arena and the DEVICE view is returned in its place (no intermediate
    array: with R trials in flight a pass stages hundreds of problems, several MB).  The copies
    run in C (csrc/hostpack.c: a memcpy and a memset per problem, no per-problem numpy call)
    when the helper is built and the inputs are contiguous float64 arrays; anything else
    (lists, strided or integer arrays) is converted and copied by a loop of slice assignments
    (~1 us each; a concatenate + fancy-index scatter measured 2.7x slower than that loop)."""
    B = len(arrs)
    lens = [len(a) for a in arrs]
    pad = pad_to(max(lens + [1])) if n_pad is None else n_pad
    arena = _ACTIVE[-1] if stage and _ACTIVE else None
    mark = arena.used if arena is not None else 0
    region = arena.alloc((B, pad, width), F64) if arena is not None else None
    out = region[0] if region is not None else np.empty((B, pad, width))
    done = False
    if _HOSTPACK is not None and B:
        try:
            done = _HOSTPACK.pack_rows(arrs, width, pad, out) == lens
        except (TypeError, ValueError):
            done = False
    if not done:
        arrs = [np.asarray(a, dtype=np.float64).reshape(-1, width) for a in arrs]
        lens = [len(a) for a in arrs]
        if n_pad is None and pad_to(max(lens + [1])) != pad:  # (row counts only known now)
            pad = pad_to(max(lens + [1]))
            if arena is not None:
                arena.used = mark
                region = arena.alloc((B, pad, width), F64)
            out = region[0] if region is not None else np.empty((B, pad, width))
        out[...] = 0.0
        for b, a in enumerate(arrs):
            if lens[b]:
                out[b, :lens[b]] = a
    if stage:
        return (region[1] if region is not None else _h2d(out)), lens, pad
    return out, lens, pad


def upload_points(B, D, pts_list):
    """Candidate sets -> (Xs [B, m_pad, D], mv device, mv host, m_pad)."""
    Xd, mv, m_pad = pack_rows(pts_list, D, stage=True)
    return Xd, _h2d(np.asarray(mv, dtype=np.int32), I32), mv, m_pad


class DevicePosterior(object):
    """B independent GP posteriors, padded to a common n_pad, resident in HBM."""

    def __init__(self, X_list, y_list, thetas, kind, dim):
        require_cuda()
        X, y, theta, nv, nv_host = self.pack(X_list, y_list, thetas, dim)
        self._attach(X, y, theta, nv, nv_host, kind, dim)

    @staticmethod
    def pack(X_list, y_list, thetas, dim):
        """Pad and upload (through the active staging arena) -> X, y, theta, nv, nv_host."""
        B, D = len(X_list), int(dim)
        if D < 1 or D > _lib.MAX_DIM:
            raise ValueError('input dimension must be in [1, %d]' % _lib.MAX_DIM)
        Xd, nv_host, n_pad = pack_rows(X_list, D, stage=True)
        yd = pack_rows(y_list, 1, n_pad, stage=True)[0].view(B, n_pad)
        theta_host = np.asarray(thetas, dtype=np.float64).reshape(B, 3 + D)
        return (Xd, yd, _h2d(theta_host), _h2d(np.asarray(nv_host, dtype=np.int32), I32), nv_host)

    @classmethod
    def from_device(cls, X, y, theta, nv, nv_host, kind, dim):
        """Posterior over operands that already live on the device (a StepPlan's arena)."""
        self = cls.__new__(cls)
        self._attach(X, y, theta, nv, nv_host, kind, dim)
        return self

    def _attach(self, X, y, theta, nv, nv_host, kind, dim):
        self.B, self.n_pad, self.D = X.shape[0], X.shape[1], int(dim)
        self.kind = int(kind)
        self.X, self.y, self.theta, self.nv, self.nv_host = X, y, theta, nv, nv_host
        B, n_pad, dev = self.B, self.n_pad, X.device
        self.L = torch.empty((B, n_pad, n_pad), dtype=F64, device=dev)
        self.invd = torch.empty((B, n_pad // TILE, TILE, TILE), dtype=F64, device=dev)
        self.alpha = torch.empty((B, n_pad), dtype=F64, device=dev)
        self.w = torch.empty((B, n_pad), dtype=F64, device=dev)
        self.lml = torch.empty((B,), dtype=F64, device=dev)
        self.info = torch.zeros((B,), dtype=I32, device=dev)
        self.jitter = [None] * B
        self.built = False
        self.unchecked = False
        self.has_alpha = False

    # -- build_posterior: K + noise I -> L -> alpha, LML  (gp_interface.py:85-88) --
    def _assemble_K(self, b0, b1):
        gram(self.kind, self.X[b0:b1], self.nv[b0:b1], self.X[b0:b1], self.nv[b0:b1],
             self.theta[b0:b1], self.L[b0:b1],
             _lib.GRAM_SYM | _lib.GRAM_LOWER | _lib.GRAM_ADD_NOISE)

    def build(self, check=True, need_alpha=True):
        """K + noise I -> L (stable_cholesky) -> alpha, LML.  ``check=False`` enqueues the plain
        factorisation only and leaves ``unchecked`` set: the caller reads ``info`` together with
        its results and calls ``build()`` again (host-walked ladder) if any problem failed --
        with noise on the diagonal that is the rare case."""
        self._assemble_K(0, self.B)
        self.unchecked = not check
        if check:
            self.jitter = chol_with_ladder(self._assemble_K, self.L, self.invd, self.info, self.nv,
                                           self.nv_host)
        else:
            self.info.zero_()
            potrf(self.L, self.invd, self.info, self.nv)
        # w = L^{-1}(y - mu0) and the LML always (mean = mu0 + V^T w wherever V = L^{-1} K(X, Xs) is
        # at hand: one pass over V, no kernel evaluations); alpha = L^{-T} w, the back-substitution
        # half of the single-CTA solve, only when asked for or on first use (ensure_alpha)
        self._solve(need_alpha)
        self.built = True
        return self

    def build_device_ladder(self, spec=None, need_alpha=False):
        """``build`` with stable_cholesky walked on the device (PendingCholesky: plain attempt and
        ``spec`` speculative rungs, no host round trip).  Returns the pending factorisation: the
        caller reads its status with its results and falls back to ``build()`` if a problem
        still fails."""
        self._assemble_K(0, self.B)
        # prior Gram matrices carry the noise term and factor at the plain attempt: speculating
        # the first rung side by side would only double the work (LML batches fill the GPU)
        pend = PendingCholesky(self.L, self.nv, spec, concurrent_first=False)
        self.invd, self.info = pend.invd, pend.info
        self.unchecked = True
        self._solve(need_alpha)
        self.built = True
        return pend

    # -- candidates -------------------------------------------------------------
    def upload_points(self, pts_list):
        return upload_points(self.B, self.D, pts_list)

    def _solve(self, with_alpha):
        call('ocbo_solve_alpha_w_lml', _p(self.L), self.n_pad, self.L.stride(1), self.L.stride(0),
             _p(self.invd), _p(self.y), self.y.stride(0), _p(self.nv), _p(self.theta),
             self.theta.stride(0), _p(self.alpha) if with_alpha else None, self.alpha.stride(0),
             _p(self.w), self.w.stride(0), _p(self.lml), _p(self.info), self.B, _stream())
        _count(1)
        self.has_alpha = bool(with_alpha)

    def ensure_alpha(self):
        if not self.has_alpha:
            self._solve(True)
        return self.alpha

    def mean(self, Xs, mv_dev):
        self.ensure_alpha()
        B, m_pad, _ = Xs.shape
        out = torch.empty((B, m_pad), dtype=F64, device=Xs.device)
        call('ocbo_post_mean', self.kind, self.D, _p(Xs), m_pad, Xs.stride(0), _p(mv_dev), _p(self.X),
             self.n_pad, self.X.stride(0), _p(self.nv), _p(self.theta), self.theta.stride(0),
             _p(self.alpha), self.alpha.stride(0), _p(out), out.stride(0), B, _stream())
        _count(1)
        return out

    def mean_from_v(self, V, mv_dev):
        """mean = mu0 + V^T w with V = L^{-1} K(X, Xs) from ``solve_cross`` (same value as
        ``mean``; HBM-bound single pass over V instead of m n kernel evaluations)."""
        B, _, m_pad = V.shape
        out = torch.empty((B, m_pad), dtype=F64, device=V.device)
        call('ocbo_post_mean_var_from_v', _p(V), self.n_pad, V.stride(1), V.stride(0), _p(self.nv),
             _p(self.w), self.w.stride(0), _p(self.theta), self.theta.stride(0), m_pad, _p(mv_dev),
             _p(out), out.stride(0), None, 0, B, _stream())
        _count(1)
        return out

    def solve_cross(self, Xs, mv_dev):
        """V = L^{-1} K(X, Xs)  [B, n_pad, m_pad]."""
        B, m_pad, _ = Xs.shape
        V = torch.empty((B, self.n_pad, m_pad), dtype=F64, device=Xs.device)
        gram(self.kind, self.X, self.nv, Xs, mv_dev, self.theta, V, 0)
        trsm_lower(self.L, self.invd, V, self.info)
        return V

    def cov(self, Xs, mv_dev, V, lower_only, out=None, b0=0, b1=None):
        b1 = self.B if b1 is None else b1
        Bn, m_pad = b1 - b0, Xs.shape[1]
        if out is None:
            out = torch.empty((self.B, m_pad, m_pad), dtype=F64, device=Xs.device)
        o, Xv, Vv, mvv, th = out[b0:b1], Xs[b0:b1], V[b0:b1], mv_dev[b0:b1], self.theta[b0:b1]
        call('ocbo_post_cov', self.kind, self.D, _p(Xv), m_pad, Xv.stride(0), _p(mvv), _p(Vv),
             self.n_pad, Vv.stride(1), Vv.stride(0), _p(th), th.stride(0), _p(o), o.stride(1),
             o.stride(0), 1 if lower_only else 0, Bn, _stream())
        _count(1)
        return out

    def cross_cov(self, Xa, mav, Va, Xb, mbv, Vb):
        B, ma, mb = self.B, Xa.shape[1], Xb.shape[1]
        out = torch.empty((B, ma, mb), dtype=F64, device=Xa.device)
        call('ocbo_post_cross_cov', self.kind, self.D, _p(Xa), ma, Xa.stride(0), _p(mav), _p(Xb), mb,
             Xb.stride(0), _p(mbv), _p(Va), Va.stride(1), Va.stride(0), _p(Vb), Vb.stride(1),
             Vb.stride(0), self.n_pad, _p(self.theta), self.theta.stride(0), _p(out), out.stride(1),
             out.stride(0), B, _stream())
        _count(1)
        return out


GRAPH_SPEC_RUNGS = 2    # rungs captured into a StepPlan's graph (a parked rung replays in ~1 us per kernel)
NO_RUNG = -100          # rung[b] while no jitter has been needed
LADDER = range(-11, 5)  # stable_cholesky's exponents


def ladder_rungs(A, A0, invd, info, rung, nv, exponents, ws=None):
    """Enqueue ladder rungs for a whole batch (include/ocbo_b200.h, ocbo_ladder_*): problems
    whose last attempt failed are restored from ``A0`` with jitter 10^p max(diag) and
    refactored; the others are parked.  No synchronisation."""
    B, n, _ = A.shape
    st = _stream()
    for p in exponents:
        call('ocbo_ladder_restore', _p(A), _p(A0), n, A.stride(1), A.stride(0), A0.stride(1),
             A0.stride(0), _p(nv), _p(info), float(10.0 ** p), B, st)
        call('ocbo_ladder_advance', _p(info), _p(rung), int(p), 0, B, st)
        _count(2)
        potrf(A, invd, info, nv, ws)
    call('ocbo_ladder_advance', _p(info), _p(rung), 0, 1, B, st)
    _count(1)


_SIDE = {}
_CHUNK_STREAMS = {}


def chunk_stream(k):
    """Persistent stream number k (per device) for the chunks of a split pass."""
    key = (torch.cuda.current_device(), k)
    s = _CHUNK_STREAMS.get(key)
    if s is None:
        s = _CHUNK_STREAMS[key] = torch.cuda.Stream()
    return s


def _side_stream():
    dev = torch.cuda.current_device()
    s = _SIDE.get(dev)
    if s is None:
        s = _SIDE[dev] = torch.cuda.Stream(device=dev)
    return s


class PendingCholesky(object):
    """A batched stable_cholesky in flight: the plain attempt and the first ``spec`` rungs
    are enqueued without a host round trip (small problems hit the ladder routinely:
    MTS samples at the queried points, so Sigma is numerically singular); ``status`` is the
    one synchronisation, ``resolve`` walks the remaining rungs for whatever still fails."""

    def __init__(self, A, nv_dev, spec=None, concurrent_first=True):
        B, m_pad, _ = A.shape
        nt = m_pad // TILE
        self.A, self.nv = A, nv_dev
        self.invd = torch.empty((B, nt, TILE, TILE), dtype=F64, device=A.device)
        self.info = torch.zeros((B,), dtype=I32, device=A.device)
        self.rung = torch.full((B,), NO_RUNG, dtype=I32, device=A.device)
        self.A0 = A.clone()
        _count(3)
        if spec is None:
            # a parked rung still costs its (empty) launches: speculate only when a rung is cheap
            spec = 2 if nt <= 4 else 0
        self.next = LADDER[0] + spec
        if spec and concurrent_first:
            self._first_rung_concurrently()
            if spec > 1:
                ladder_rungs(A, self.A0, self.invd, self.info, self.rung, nv_dev, LADDER[1:spec])
            return
        potrf(A, self.invd, self.info, nv_dev)
        if spec:
            ladder_rungs(A, self.A0, self.invd, self.info, self.rung, nv_dev, LADDER[:spec])

    def _first_rung_concurrently(self):
        """The plain attempt on ``A`` and rung -11 on a copy, side by side on two streams
        (include/ocbo_b200.h, ocbo_ladder_adopt): in MTS the first almost always fails and
        the second almost always succeeds, so the factorisation's latency is max, not sum.
        Same kernels on the same data as the sequential walk: same bits."""
        A, A0, nv = self.A, self.A0, self.nv
        B, n, _ = A.shape
        p = LADDER[0]
        cur, side = torch.cuda.current_stream(), _side_stream()
        A1 = torch.empty_like(A)
        invd1 = torch.empty_like(self.invd)
        info1 = torch.zeros_like(self.info)
        every = torch.ones_like(self.info)  # restore gate: all problems
        if not torch.cuda.is_current_stream_capturing():  # (a graph's pool is never recycled)
            for t in (A1, invd1, info1, every, A0) + ((nv,) if nv is not None else ()):
                t.record_stream(side)
        fork = torch.cuda.Event()
        fork.record(cur)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            call('ocbo_ladder_restore', _p(A1), _p(A0), n, A1.stride(1), A1.stride(0), A0.stride(1),
                 A0.stride(0), _p(nv), _p(every), float(10.0 ** p), B, _stream())
            _count(1)
            potrf(A1, invd1, info1, nv)
            join = torch.cuda.Event()
            join.record(side)
        potrf(A, self.invd, self.info, nv)
        cur.wait_event(join)
        call('ocbo_ladder_adopt', _p(A), _p(A1), n, A.stride(1), A.stride(0), _p(self.invd), _p(invd1),
             _p(self.info), _p(info1), _p(self.rung), int(p), B, _stream())
        _count(2)
        self._spec_buffers = (A1, invd1, info1, every)  # alive until the adopt has been enqueued

    def status(self):
        return self.info.cpu().numpy(), self.rung.cpu().numpy()

    def resolve(self, status=None):
        """Finish the ladder (one synchronisation per further rung).  Returns the jitter
        exponent per problem (None = none needed); ValueError like stable_cholesky."""
        info_h, rung_h = status if status is not None else self.status()
        while (info_h > 0).any():
            if self.next > LADDER[-1]:
                raise ValueError('stable_cholesky: matrix could not be made PD.')
            ladder_rungs(self.A, self.A0, self.invd, self.info, self.rung, self.nv, [self.next])
            self.next += 1
            info_h, rung_h = self.status()
        return [None if r == NO_RUNG else int(r) for r in rung_h]


def factor_covariance(post_or_none, Sigma, mv_dev, mv_host=None, reassemble=None):
    """stable_cholesky of a batch of (padded) covariance matrices, in place (synchronous
    form of PendingCholesky; ``reassemble`` is not needed any more: retries restore from a
    device copy)."""
    pend = PendingCholesky(Sigma, mv_dev)
    jit = pend.resolve()
    return pend.invd, pend.info, jit


def sample_from_factor(mean, Lsig, Z, info=None):
    """F = mean + L Z  -> [B, m_pad, S_pad] (mean [B, m_pad], Z [B, m_pad, S_pad])."""
    B, m_pad, _ = Lsig.shape
    S = Z.shape[2]
    Fm = torch.empty((B, m_pad, S), dtype=F64, device=Lsig.device)
    call('ocbo_sample', _p(mean), mean.stride(0), _p(Lsig), m_pad, Lsig.stride(1), Lsig.stride(0),
         _p(Z), S, Z.stride(1), Z.stride(0), _p(Fm), Fm.stride(1), Fm.stride(0), _p(info), B, _stream())
    _count(1)
    return Fm


def philox_normals(B, m_pad, S, seed, trial_ids, dev):
    Z = torch.empty((B, m_pad, S), dtype=F64, device=dev)
    tid = _h2d(np.asarray(trial_ids, dtype=np.int32), I32)
    call('ocbo_philox_normal', _p(Z), m_pad, S, Z.stride(1), Z.stride(0),
         ctypes.c_uint64(int(seed)), _p(tid), B, _stream())
    _count(1)
    return Z


def segment_argmax(Fm, S, seg_ptr_host):
    """seg_ptr_host: int array [B, nseg + 1] -> (max [B, nseg, S], arg [B, nseg, S])."""
    if torch.is_tensor(seg_ptr_host):
        seg_d = seg_ptr_host
        B, nseg = seg_d.shape[0], seg_d.shape[1] - 1
    else:
        seg = np.asarray(seg_ptr_host, dtype=np.int32)
        B, nseg = seg.shape[0], seg.shape[1] - 1
        seg_d = _h2d(seg, I32)
    mx = torch.empty((B, nseg, S), dtype=F64, device=Fm.device)
    ag = torch.empty((B, nseg, S), dtype=I32, device=Fm.device)
    call('ocbo_segment_argmax', _p(Fm), S, Fm.stride(1), Fm.stride(0), _p(seg_d), nseg, _p(mx),
         _p(ag), B, _stream())
    _count(1)
    return mx, ag


def joint_pick(mx, ag, n_old_seg, T):
    B, nseg, S = mx.shape
    task = torch.empty((B, S), dtype=I32, device=mx.device)
    act = torch.empty((B, S), dtype=I32, device=mx.device)
    gain = torch.empty((B, S), dtype=F64, device=mx.device)
    call('ocbo_joint_pick', _p(mx), _p(ag), nseg, n_old_seg, T, S, _p(task), _p(act), _p(gain), B,
         _stream())
    _count(1)
    return task, act, gain


def knowledge_gradients(judge_means, ld_means, inter, cand_var, noise, C, G, E):
    """kg[c, g] of REVI (include/ocbo_b200.h, ocbo_kg): ``judge_means`` holds G groups of E
    values (group stride ``ld_means``), ``inter`` [C_pad, >= G*E] the posterior covariances
    candidate x judgement point, ``cand_var`` [>= C] the candidates' posterior variances."""
    out = torch.empty((C, G), dtype=F64, device=inter.device)
    ld_inter = inter.stride(0) if inter.shape[0] > 1 else inter.shape[1]  # (a 1-row stride is arbitrary)
    call('ocbo_kg', _p(judge_means), int(ld_means), _p(inter), int(ld_inter), _p(cand_var),
         float(noise), int(C), int(G), int(E), _p(out), _stream())
    _count(1)
    return out


def pad_normals(Z_list, mv, m_pad, S):
    """Host normals [(m_b, S)] -> device [B, m_pad, S_pad] (zero padded)."""
    S_pad = S if S <= 8 else pad_to(S, 64)
    if S_pad == S:
        return pack_rows(Z_list, S, m_pad, stage=True)[0]
    Zp = np.zeros((len(Z_list), m_pad, S_pad))
    Zp[:, :, :S] = pack_rows(Z_list, S, m_pad)[0]
    return _h2d(Zp)
