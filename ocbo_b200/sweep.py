"""Joint-MTS posterior-sampling step for a batch of independent BO trials
(BASELINE.json configs[4], SURVEY.md 8d "sweep").

One step, per trial, is the whole hot path of ``JointMTS._find_best_query``
(/root/reference/src/strategies/joint_mts.py:24-57) at fixed hyper-parameters:

  K = k(X,X) + noise I -> L_K -> alpha            (build_posterior, a6)
  mu = mu0 + k(Xs,X) alpha                        (eval mean, a4)
  V = L_K^{-1} k(X,Xs); Sigma = k(Xs,Xs) - V^T V  (eval covar, a4)
  L_S = chol(Sigma)                               (stable_cholesky, a5)
  F = mu + L_S Z, Z ~ Philox(seed, trial)         (draw_samples, a5; S draws)
  per draw: task = argmax_t max_a F[t,a], action  (the pick, a2)

Everything is resident in HBM; a step only moves the trial inputs in and the
picks out.  Trials are independent, so a step processes ``batch`` of them per
kernel launch (grid.z) and several ``SweepStep`` objects may run on different
CUDA streams to fill the latency-bound panel phases of one trial with the
trailing updates of another.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib, engine
from ._lib import TILE, call
from .engine import F64, I32, _p

from .workload import (SEED, SWEEP_BANDWIDTHS, hartmann6, make_trial_inputs,  # noqa: F401
                       step_flops)


class SweepStep(object):
    """Pre-allocated device state + the kernel sequence for ``batch`` trials."""

    def __init__(self, n=1024, T=64, A=128, S=256, dim=6, batch=1, kernel_type='se',
                 scale=1.0, bandwidths=SWEEP_BANDWIDTHS, noise_var=1e-2, mu0=0.0, seed=SEED,
                 stream=None, fused_tail=None, ladder_rungs=0, ozaki=None):
        engine.require_cuda()
        if n % TILE or (T * A) % TILE:
            raise ValueError('n and T*A must be multiples of %d' % TILE)
        if S > 8 and S % 64:
            raise ValueError('S must be <= 8 or a multiple of 64')
        self.n, self.T, self.A, self.S, self.D, self.B = n, T, A, S, dim, batch
        self.m = T * A
        self.kind = engine.kernel_kind(kernel_type)
        self.seed = seed
        self.stream = stream or torch.cuda.current_stream()
        dev = engine.device()
        B, m, D = batch, self.m, dim
        e = lambda *shape, dt=F64: torch.empty(shape, dtype=dt, device=dev)
        self.X, self.y, self.Xs = e(B, n, D), e(B, n), e(B, m, D)
        bws = np.broadcast_to(np.asarray(bandwidths, dtype=np.float64), (D,))
        th = np.concatenate([[scale, noise_var, mu0], bws])
        self.theta = torch.from_numpy(np.tile(th, (B, 1))).to(dev)
        self.LK, self.invdK = e(B, n, n), e(B, n // TILE, TILE, TILE)
        self.alpha, self.lml, self.mean = e(B, n), e(B), e(B, m)
        self.V = e(B, n, m)
        self.Sigma, self.invdS = e(B, m, m), e(B, m // TILE, TILE, TILE)
        # Fused sampling tail (ocbo_sample_tile_argmax): with task segments of exactly one 128-row
        # tile the per-task max / arg-max come out of the L Z kernel's epilogue and the m x S
        # sample matrix F is never written; other layouts materialise F and reduce it.
        can_fuse = A == TILE and S % 32 == 0
        if fused_tail and not can_fuse:
            raise ValueError('the fused sampling tail needs A == %d and S %% 32 == 0' % TILE)
        self.fused_tail = can_fuse if fused_tail is None else bool(fused_tail)
        self.Z = e(B, m, S)
        # V^T V of the posterior covariance and the trailing updates of chol(Sigma) on the int8 tensor cores
        # (ocbo_post_cov_ws / ocbo_potrf_ws, csrc/ozaki.cu; shapes too small for it keep the fp64 DMMA
        # kernels).  The library never allocates: the step owns the slice workspace.  Default on;
        # OCBO_OZAKI=0 or ozaki=False keeps every product on the fp64 tensor path.
        if ozaki is None:
            ozaki = os.environ.get('OCBO_OZAKI', '1') != '0'
        self.ozaki = bool(ozaki)
        self.ws = None
        if self.ozaki:
            nb = max(int(_lib.LIB.ocbo_potrf_ws_bytes(m, B)), int(_lib.LIB.ocbo_post_cov_ws_bytes(m, n, B)),
                     int(_lib.LIB.ocbo_sample_ws_bytes(m, S, B)) if (can_fuse and S % 64 == 0 and
                                                                   os.environ.get('OCBO_OZAKI_SAMPLE', '1') != '0') else 0)
            if nb > 0:
                self.ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
        # stable_cholesky's jitter ladder for Sigma walked on the device (engine.ladder_rungs):
        # the plain attempt and the first ``ladder_rungs`` rungs (10^-11, 10^-10, ... x max diag)
        # are enqueued without a host round trip; a trial that fails them all is redone from the
        # host (_redo_with_ladder).  0 = plain attempt only (the headline configuration never fails).
        self.ladder_rungs = int(ladder_rungs)
        if self.ladder_rungs:
            self.Sigma0 = e(B, m, m)
            self.every = torch.ones((B,), dtype=I32, device=dev)
        self.rung = torch.full((B,), engine.NO_RUNG, dtype=I32, device=dev)
        self.hrung = torch.empty((B,), dtype=I32).pin_memory()
        self._F = None if self.fused_tail else e(B, m, S)
        self.info = torch.zeros((2, B), dtype=I32, device=dev)
        seg = np.tile((A * np.arange(T + 1)).astype(np.int32), (B, 1))
        self.seg = torch.from_numpy(seg).to(dev)
        self.seg_max, self.seg_arg = e(B, T, S), e(B, T, S, dt=I32)
        self.task, self.action, self.gain = e(B, S, dt=I32), e(B, S, dt=I32), e(B, S)
        self.trial_ids = torch.zeros((B,), dtype=I32, device=dev)
        # pinned staging for the host-buffer API
        pin = lambda *shape, dt=F64: torch.empty(shape, dtype=dt).pin_memory()
        self.hX, self.hy, self.hXs = pin(B, n, D), pin(B, n), pin(B, m, D)
        self.htrial = pin(B, dt=I32)
        self.htask, self.haction = pin(B, S, dt=I32), pin(B, S, dt=I32)
        self.hgain = pin(B, S)
        self.hinfo = pin(2, B, dt=I32)
        # kernels per run(): gram, potrf(n), solve, mean, gram, trsm, cov, potrf(m), philox,
        # sample, segment, pick
        self.launches_per_run = 1 + (3 * (n // TILE) - 2) + 1 + 1 + 1 + (2 * (n // TILE) - 1) + 1 \
            + (3 * (self.m // TILE) - 2) + 1 + 1 + 1 + (0 if self.fused_tail else 1)
        self.h2d_bytes = B * (n * D + n + m * D) * 8 + B * 4
        self.d2h_bytes = B * S * (8 + 8) + 2 * B * 4

    # -- device-resident step -------------------------------------------------
    def run(self):
        """Launch the whole step on ``self.stream`` (asynchronous)."""
        st = ctypes.c_void_p(self.stream.cuda_stream)
        B, n, m, D, S, T = self.B, self.n, self.m, self.D, self.S, self.T
        X, Xs, th = self.X, self.Xs, self.theta
        infoK, infoS = self.info[0], self.info[1]
        with torch.cuda.stream(self.stream):
            self.info.zero_()
        call('ocbo_gram', self.kind, D, _p(X), n, X.stride(0), None, _p(X), n, X.stride(0), None,
             _p(th), th.stride(0), _p(self.LK), n, self.LK.stride(0), B,
             _lib.GRAM_SYM | _lib.GRAM_LOWER | _lib.GRAM_ADD_NOISE, st)
        call('ocbo_potrf', _p(self.LK), n, n, self.LK.stride(0), _p(self.invdK), _p(infoK), B, st)
        # w = L_K^{-1}(y - mu0) (+ LML); the mean is mu0 + V^T w, one HBM pass over V
        # instead of a second sweep over the kernel and a back-substitution
        call('ocbo_solve_forward_lml', _p(self.LK), n, n, self.LK.stride(0), _p(self.invdK),
             _p(self.y), self.y.stride(0), None, _p(th), th.stride(0), _p(self.alpha),
             self.alpha.stride(0), _p(self.lml), _p(infoK), B, st)
        call('ocbo_gram', self.kind, D, _p(X), n, X.stride(0), None, _p(Xs), m, Xs.stride(0), None,
             _p(th), th.stride(0), _p(self.V), m, self.V.stride(0), B, 0, st)
        call('ocbo_trsm_lower', _p(self.LK), n, n, self.LK.stride(0), _p(self.invdK), _p(self.V), m,
             m, self.V.stride(0), _p(infoK), B, st)
        call('ocbo_post_mean_var_from_v', _p(self.V), n, m, self.V.stride(0), None, _p(self.alpha),
             self.alpha.stride(0), _p(th), th.stride(0), m, None, _p(self.mean), self.mean.stride(0),
             None, 0, B, st)
        if self.ws is not None:
            call('ocbo_post_cov_ws', self.kind, D, _p(Xs), m, Xs.stride(0), None, _p(self.V), n, m,
                 self.V.stride(0), _p(th), th.stride(0), _p(self.Sigma), m, self.Sigma.stride(0), 1, B,
                 _p(self.ws), self.ws.numel(), st)
        else:
            call('ocbo_post_cov', self.kind, D, _p(Xs), m, Xs.stride(0), None, _p(self.V), n, m,
                 self.V.stride(0), _p(th), th.stride(0), _p(self.Sigma), m, self.Sigma.stride(0), 1, B, st)
        if self.ladder_rungs:
            with torch.cuda.stream(self.stream):
                self.rung.fill_(engine.NO_RUNG)
                # keep Sigma's lower blocks: restore with an all-pass gate and zero jitter, Sigma0 <- Sigma
                call('ocbo_ladder_restore', _p(self.Sigma0), _p(self.Sigma), m, m, self.Sigma0.stride(0),
                     m, self.Sigma.stride(0), None, _p(self.every), 0.0, B, st)
        self._potrf_sigma(st)
        if self.ladder_rungs:
            with torch.cuda.stream(self.stream):
                engine.ladder_rungs(self.Sigma, self.Sigma0, self.invdS, infoS, self.rung, None,
                                    engine.LADDER[:self.ladder_rungs], ws=self.ws)
        call('ocbo_philox_normal', _p(self.Z), m, S, S, self.Z.stride(0),
             ctypes.c_uint64(self.seed), _p(self.trial_ids), B, st)
        self._sample_and_reduce(slice(0, B), st, l_in_ws=True)
        call('ocbo_joint_pick', _p(self.seg_max), _p(self.seg_arg), T, 0, T, S, _p(self.task),
             _p(self.action), _p(self.gain), B, st)
        engine._count(self.launches_per_run)

    def _potrf_sigma(self, st):
        m = self.m
        if self.ws is not None:
            call('ocbo_potrf_ws', _p(self.Sigma), m, m, self.Sigma.stride(0), _p(self.invdS), _p(self.info[1]),
                 self.B, _p(self.ws), self.ws.numel(), st)
        else:
            call('ocbo_potrf', _p(self.Sigma), m, m, self.Sigma.stride(0), _p(self.invdS), _p(self.info[1]),
                 self.B, st)

    @property
    def F(self):
        """[B, m, S] joint samples mu + L Z (diagnostics / parity tests).  On the fused path they
        are not part of a step: computed here on demand from the step's factor and normals."""
        if self._F is None:
            self._F = torch.empty((self.B, self.m, self.S), dtype=F64, device=self.Z.device)
        if self.fused_tail:
            m, S = self.m, self.S
            with torch.cuda.stream(self.stream):
                call('ocbo_sample', _p(self.mean), self.mean.stride(0), _p(self.Sigma), m, m,
                     self.Sigma.stride(0), _p(self.Z), S, S, self.Z.stride(0), _p(self._F), S,
                     self._F.stride(0), _p(self.info[1]), self.B, ctypes.c_void_p(self.stream.cuda_stream))
            self.stream.synchronize()
        return self._F

    def _sample_and_reduce(self, s, st, l_in_ws=False):
        """mu + L Z -> per-task max / first arg-max of every draw, for the trial slots ``s``.
        ``l_in_ws``: Sigma was factored by ocbo_potrf_ws with this step's workspace just before (run()): the
        factor's int8 slices are still there."""
        m, S, T = self.m, self.S, self.T
        nb = s.stop - s.start
        sd = lambda t: 0 if nb == 1 else t.stride(0)
        mean, Sig, Z, info = self.mean[s], self.Sigma[s], self.Z[s], self.info[1, s]
        if self.fused_tail:
            if self.ws is not None and os.environ.get('OCBO_OZAKI_SAMPLE', '1') != '0':
                # (too small a workspace or shape: the entry point is ocbo_sample_tile_argmax)
                call('ocbo_sample_tile_argmax_ws', _p(mean), sd(mean), _p(Sig), m, m, sd(Sig), _p(Z), S, S,
                     sd(Z), _p(self.seg_max[s]), _p(self.seg_arg[s]), _p(info), nb, _p(self.ws), self.ws.numel(),
                     1 if (l_in_ws and nb == self.B) else 0, st)
            else:
                call('ocbo_sample_tile_argmax', _p(mean), sd(mean), _p(Sig), m, m, sd(Sig), _p(Z), S, S,
                     sd(Z), _p(self.seg_max[s]), _p(self.seg_arg[s]), _p(info), nb, st)
            return
        F = self._F[s]
        call('ocbo_sample', _p(mean), sd(mean), _p(Sig), m, m, sd(Sig), _p(Z), S, S, sd(Z), _p(F), S,
             sd(F), _p(info), nb, st)
        call('ocbo_segment_argmax', _p(F), S, S, sd(F), _p(self.seg[s]), T, _p(self.seg_max[s]),
             _p(self.seg_arg[s]), nb, st)

    def set_inputs_device(self, trial_ids, inputs):
        """Load trial inputs once (device-resident timing)."""
        for b, (X, y, Xs) in enumerate(inputs):
            self.hX[b].copy_(torch.from_numpy(X))
            self.hy[b].copy_(torch.from_numpy(y))
            self.hXs[b].copy_(torch.from_numpy(Xs))
            self.htrial[b] = int(trial_ids[b])
        self._upload()

    def _upload(self):
        with torch.cuda.stream(self.stream):
            self.X.copy_(self.hX, non_blocking=True)
            self.y.copy_(self.hy, non_blocking=True)
            self.Xs.copy_(self.hXs, non_blocking=True)
            self.trial_ids.copy_(self.htrial, non_blocking=True)

    def _download(self):
        with torch.cuda.stream(self.stream):
            self.htask.copy_(self.task, non_blocking=True)
            self.haction.copy_(self.action, non_blocking=True)
            self.hgain.copy_(self.gain, non_blocking=True)
            self.hinfo.copy_(self.info, non_blocking=True)
            self.hrung.copy_(self.rung, non_blocking=True)

    # -- host-buffer API (what bench.py's e2e leg times) ------------------------
    def submit_host(self, trial_ids, inputs):
        """Stage host inputs (pinned), H2D, run, D2H of the picks; asynchronous."""
        for b, (X, y, Xs) in enumerate(inputs):
            self.hX[b].copy_(torch.from_numpy(X))
            self.hy[b].copy_(torch.from_numpy(y))
            self.hXs[b].copy_(torch.from_numpy(Xs))
            self.htrial[b] = int(trial_ids[b])
        self._upload()
        self.run()
        self._download()

    def result_host(self):
        """Synchronise the stream and return (task[B,S], action[B,S], gain[B,S])."""
        self.stream.synchronize()
        info = self.hinfo.numpy()
        self.jitter = [(None, None if r == engine.NO_RUNG else int(r)) for r in self.hrung.numpy()]
        if info.any():
            # the fast path factors without the ladder; trials whose K or Sigma is not
            # numerically PD are redone alone through Dragonfly's stable_cholesky ladder
            for b in np.nonzero(info.any(axis=0))[0]:
                self._redo_with_ladder(int(b))
            self._download()
            self.stream.synchronize()
        return self.htask.numpy().copy(), self.haction.numpy().copy(), self.hgain.numpy().copy()

    def _redo_with_ladder(self, b):
        """Slow path for one trial: same kernels, stable_cholesky semantics
         for both factorisations; the normals Z[b] are reused."""
        s = slice(b, b + 1)
        n, m, D, S, T = self.n, self.m, self.D, self.S, self.T
        X, Xs, th = self.X[s], self.Xs[s], self.theta[s]
        with torch.cuda.stream(self.stream):
            def asm_K(b0, b1):
                engine.gram(self.kind, X, None, X, None, th, self.LK[s],
                            _lib.GRAM_SYM | _lib.GRAM_LOWER | _lib.GRAM_ADD_NOISE)
            asm_K(0, 1)
            jK = engine.chol_with_ladder(asm_K, self.LK[s], self.invdK[s], self.info[0, s], None, [n])
            st = ctypes.c_void_p(self.stream.cuda_stream)
            # same kernels and formulas as run(): w = L^-1 (y - mu0), mean = mu0 + V^T w
            call('ocbo_solve_forward_lml', _p(self.LK[s]), n, n, 0, _p(self.invdK[s]), _p(self.y[s]), 0,
                 None, _p(th), 0, _p(self.alpha[s]), 0, _p(self.lml[s]), _p(self.info[0, s]), 1, st)
            engine.gram(self.kind, X, None, Xs, None, th, self.V[s], 0)
            engine.trsm_lower(self.LK[s], self.invdK[s], self.V[s], self.info[0, s])
            call('ocbo_post_mean_var_from_v', _p(self.V[s]), n, m, 0, None, _p(self.alpha[s]), 0,
                 _p(th), 0, m, None, _p(self.mean[s]), 0, None, 0, 1, st)

            def asm_S(b0, b1):
                call('ocbo_post_cov', self.kind, D, _p(Xs), m, 0, None, _p(self.V[s]), n, m, 0,
                     _p(th), 0, _p(self.Sigma[s]), m, 0, 1, 1, st)
            asm_S(0, 1)
            jS = engine.chol_with_ladder(asm_S, self.Sigma[s], self.invdS[s], self.info[1, s], None, [m])
            self._sample_and_reduce(s, st)
            call('ocbo_joint_pick', _p(self.seg_max[s]), _p(self.seg_arg[s]), T, 0, T, S,
                 _p(self.task[s]), _p(self.action[s]), _p(self.gain[s]), 1, st)
        self.jitter[b] = (jK[0], jS[0])

    def samples_host(self, b):
        """The (S, m) joint samples of trial slot ``b`` of the last step (diagnostics / parity
        tests; the picks never need them on the host)."""
        self.stream.synchronize()
        return self.F[b].t().contiguous().cpu().numpy()

    def step_host(self, trial_ids, inputs):
        self.submit_host(trial_ids, inputs)
        return self.result_host()
