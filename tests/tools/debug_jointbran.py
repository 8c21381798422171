"""GPU-box diagnostic: run the jointbran scenario with the B200 engine and, at
every step, evaluate the numpy oracle on the SAME inputs and normals; print both
picks and the oracle's top-two gain gap."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import gp_numpy as o
from tests import scenarios as sc
from ocbo_b200 import rng, synth
from ocbo_b200.gp.gp_interface import B200GPFitter
from ocbo_b200.strategies.joint_mts import JointMTS
from ocbo_b200.util import sample_grid


class Dbg(JointMTS):
    def _find_best_query(self):
        A, T = self.options.max_opt_evals, len(self.fcns)
        state = np.random.get_state()
        grid_pts = sample_grid(self.f_locs, self.domains[0], A)
        past, num_qs = [], []
        for f_idx, f_name in enumerate(self.f_names):
            for h in self.query_history[f_name]:
                past.append(np.append(self.f_locs[f_idx], h.pt))
            num_qs.append(len(self.query_history[f_name]))
        to_samp = np.vstack([np.asarray(past), grid_pts])
        z = np.random.normal(size=(len(to_samp), 1))
        core = self.gps[0].gp_core
        og = o.make_gp(np.asarray(core.X), core.Y, 'se', core.kernel.hyperparams['scale'],
                       core.kernel.hyperparams['dim_bandwidths'], core.noise_var, core._mu0())
        mu, cov = og.eval(to_samp, include_covar=True)
        L, p = o.stable_cholesky(cov, return_jitter=True)
        s = (L @ z).ravel() + mu
        tq = sum(num_qs)
        divs = np.concatenate([[0], np.cumsum(num_qs)])
        old = np.asarray([s[divs[i]:divs[i + 1]].max() for i in range(T)])[:, None]
        gains = s[tq:].reshape(T, A) - old
        flat = np.sort(gains.ravel())[::-1]
        ref_pick = np.unravel_index(np.argmax(gains), gains.shape)
        np.random.set_state(state)
        rng.NORMAL_SOURCE = None
        out = JointMTS._find_best_query(self)
        mb, cb = self.gps[0].eval(to_samp, include_covar=True)
        rng.NORMAL_SOURCE = lambda size: z.reshape(size)
        sg = self.gps[0].draw_sample(samp_pts=to_samp).ravel()
        pg = core.last_sample_jitter
        rng.NORMAL_SOURCE = None
        md = np.max(np.diag(cov))
        def chol_at(M, pp):
            try:
                return np.linalg.cholesky(M + (10.0 ** pp) * np.max(np.diag(M)) * np.eye(len(M)))
            except np.linalg.LinAlgError:
                return None
        alt = {}
        for nm, M in (('oracleSigma', cov), ('gpuSigma', cb)):
            for pp in (-11, -10, -9):
                Lx = chol_at(M, pp)
                alt[(nm, pp)] = None if Lx is None else float(np.max(np.abs((Lx @ z).ravel() + mu - sg)))
        print('   |s_gpu - s_oracle|max %.3e  sqrt(j9*md)*zmax %.3e  gpu_p %s  alt(vs gpu sample) %s'
              % (np.max(np.abs(sg - s)), np.sqrt(1e-9 * md) * np.max(np.abs(z)), pg, alt))
        print('step', self.t, 'oracle pick', ref_pick, 'gain %.6e gap %.3e jitter_p %s | b200 task %d jitter %s'
              ' | mean err %.2e cov err %.2e hps %s'
              % (flat[0], flat[0] - flat[1], p, out[0], core.last_sample_jitter,
                 np.max(np.abs(mb - mu)), np.max(np.abs(cb - cov)),
                 np.round(core.theta(), 4).tolist()))
        return out


cfg = sc.DISCRETE['jointbran']
np.random.seed(sc.SEED)
B200GPFitter._fit_counter = 0
f_infos = synth.assemble_chopped_1d_functions(cfg['cts_f'], cfg['num_chops'])
got = sc.run_discrete(cfg, Dbg, f_infos, sc.discrete_options(cfg, 'b200'))
print(got['choice_timeline'])
