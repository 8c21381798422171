"""Host-side strategy logic without a GPU: ocbo_b200's strategy classes with the
device calls replaced by the numpy oracle must reproduce the golden traces that
the REFERENCE's strategy classes produced (tests/golden/make_golden.py).  This
pins the loop order, the bookkeeping and the consumption of the global numpy
stream; the GPU test (test_gpu_strategies.py) then swaps the real engine in."""
import json
import os

import numpy as np
import pytest

from oracle import gp_numpy as o
from oracle import picks
from tests import scenarios as sc

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'mts_traces.json')


class _Core(o.EuclideanGP):
    """numpy core with the two extra entry points the B200 core offers."""

    def sample_and_pick(self, to_samp, seg_ptr, num_tasks):
        from ocbo_b200 import rng
        z = rng.normals((len(to_samp), 1))
        mu, cov = self.eval(to_samp, 'covar')
        s = (o.stable_cholesky(cov).dot(z)).ravel() + mu
        num_qs = list(np.diff(np.asarray(seg_ptr)[:num_tasks + 1]))
        return picks.joint_mts_pick(s, num_qs, num_tasks)


class _Fitter(o.OracleGPFitter):
    def get_next_gp(self):
        theta = self.gpf_core.hp_list[self.gpf_core._next % len(self.gpf_core.hp_list)]
        self.gpf_core._next += 1
        D = self.gpf_core.dim
        kern = o.make_kernel(self.options.kernel_type, D, theta[0], theta[3:3 + D])
        core = _Core(self.gpf_core.X, self.gpf_core.Y, kern, o.ConstMean(theta[2]), theta[1],
                     build_posterior=False)
        return o.OracleGP(core, self.options)


def _sample(gp, pts, z):
    mu, cov = gp.gp_core.eval(pts, 'covar')
    return (o.stable_cholesky(cov).dot(np.asarray(z).reshape(-1, 1))).ravel() + mu, mu


def _fake_mts_step(gps, samp_pts_list, n_old_list, Z_list):
    out = []
    for gp, pts, n_old, z in zip(gps, samp_pts_list, n_old_list, Z_list):
        s, _ = _sample(gp, pts, z)
        out.append((float(np.max(s[:n_old])), int(np.argmax(s[n_old:])), float(np.max(s[n_old:]))))
    return out


def _fake_profile_step(gp, act_sets, Z_list):
    res = []
    for pts, z in zip(act_sets, Z_list):
        s, mu = _sample(gp, pts, z)
        res.append((np.max(s), int(np.argmax(s)), np.max(mu), int(np.argmax(mu)), s[int(np.argmax(mu))]))
    return [np.asarray(c) for c in zip(*res)]


def _ei(gp, pts, best):
    from scipy.stats import norm
    mu, cov = gp.gp_core.eval(pts, 'covar')
    sd = np.sqrt(np.diag(cov))
    z = (mu - best) / sd
    ei = sd * (z * norm.cdf(z) + norm.pdf(z))
    return float(np.max(ei)), int(np.argmax(ei)), float(np.max(mu))


def _fake_mei_step(gps, pts_list, best_list):
    return [_ei(g, p, b)[:2] for g, p, b in zip(gps, pts_list, best_list)]


def _fake_shared_ei_step(gp, pts_list, best_list=None, cap=None):
    out = []
    for i, p in enumerate(pts_list):
        if cap is not None:
            best = min(float(np.max(gp.gp_core.eval(p)[0])), float(cap))
        else:
            best = best_list[i]
        out.append(_ei(gp, p, best)[:2])
    return np.asarray([o[0] for o in out]), np.asarray([o[1] for o in out])


def _fake_revi_improvements(gp, cand_pts, judge_pts, num_groups, group_size):
    from oracle import kg
    cand, judge = np.asarray(cand_pts), np.asarray(judge_pts)
    means = gp.eval(judge)[0].reshape(num_groups, group_size)
    inter = gp.get_pt_relations(cand, judge)
    cand_vars = gp.eval(cand, include_covar=True)[1].diagonal()
    return kg.revi_improvements(means, inter, cand_vars, gp.get_estimated_noise())


@pytest.fixture
def oracle_engine(monkeypatch):
    from ocbo_b200 import batched
    from ocbo_b200.cstrats import cts_opt
    from ocbo_b200.strategies import joint_opt, multi_opt
    fitter = lambda engine, x, y, opts: _Fitter(x, y, opts)
    for mod in (multi_opt, joint_opt, cts_opt):
        monkeypatch.setattr(mod, 'get_gp_fitter', fitter)
    monkeypatch.setattr(batched, 'mts_step', _fake_mts_step)
    monkeypatch.setattr(batched, 'profile_step', _fake_profile_step)
    monkeypatch.setattr(batched, 'mei_step', _fake_mei_step)
    monkeypatch.setattr(batched, 'shared_ei_step', _fake_shared_ei_step)
    monkeypatch.setattr(batched, 'revi_improvements', _fake_revi_improvements)
    o.EuclideanGPFitter._fit_counter = 0
    with open(GOLDEN) as fh:
        return json.load(fh)


def _functions(cfg):
    from ocbo_b200 import synth
    if 'functions' in cfg:
        return synth.assemble_functions(cfg['functions'])
    if 'cts_f' in cfg:
        return synth.assemble_chopped_1d_functions(cfg['cts_f'], cfg['num_chops'])
    easy, med, hard = cfg['masses']
    return synth.random_mass_functions(easy, med, hard, cfg['massf_dim'])


@pytest.mark.parametrize('name', sorted(sc.DISCRETE))
def test_discrete_host_logic_matches_reference(oracle_engine, name):
    from ocbo_b200.strategies import strategies
    cfg = sc.DISCRETE[name]
    np.random.seed(sc.SEED)
    f_infos = _functions(cfg)
    cls = [s.impl for s in strategies if s.name == cfg['kind']][0]
    got = sc.run_discrete(cfg, cls, f_infos, sc.discrete_options(cfg, 'b200'))
    ref = oracle_engine[name]
    assert got['choice_timeline'] == ref['choice_timeline']
    assert np.allclose(got['points'], ref['points'], rtol=0, atol=1e-12)
    assert np.allclose(got['values'], ref['values'], rtol=0, atol=1e-12)
    assert np.allclose(got['total_best'], ref['total_best'], rtol=0, atol=1e-12)


@pytest.mark.parametrize('name', sorted(sc.DISCRETE_BASELINES))
def test_discrete_baselines_match_reference(oracle_engine, name):
    """agn-thomp / agn-ei / Random / ja-thomp / ja-ei / joint-rand: same picks, points and values
    as the reference's own classes on the same engine and seed."""
    from ocbo_b200.strategies import strategies
    cfg = sc.DISCRETE_BASELINES[name]
    np.random.seed(sc.SEED)
    f_infos = _functions(cfg)
    cls = [s.impl for s in strategies if s.name == cfg['kind']][0]
    got = sc.run_discrete(cfg, cls, f_infos, sc.discrete_options(cfg, 'b200'))
    ref = oracle_engine[name]
    assert got['choice_timeline'] == ref['choice_timeline']
    assert np.allclose(got['points'], ref['points'], rtol=0, atol=1e-12)
    assert np.allclose(got['values'], ref['values'], rtol=0, atol=1e-12)
    assert np.allclose(got['total_best'], ref['total_best'], rtol=0, atol=1e-12)


@pytest.mark.parametrize('name', sorted(sc.CONTINUOUS_BASELINES))
def test_continuous_baselines_match_reference(oracle_engine, name):
    from ocbo_b200 import synth
    from ocbo_b200.cstrats import cstrats
    cfg = sc.CONTINUOUS_BASELINES[name]
    np.random.seed(sc.SEED)
    cls = [s.impl for s in cstrats if s.name == cfg['kind']][0]
    got = sc.run_continuous(cfg, cls, synth.hartmann6, [[0, 1] for _ in range(6)],
                            sc.continuous_options(cfg, 'b200'))
    ref = oracle_engine[name]
    assert np.allclose(got['points'], ref['points'], rtol=0, atol=1e-12)
    assert np.allclose(got['regret'], ref['regret'], rtol=0, atol=1e-12)


@pytest.mark.parametrize('name', sorted(sc.CONTINUOUS))
def test_continuous_host_logic_matches_reference(oracle_engine, name):
    from ocbo_b200 import synth
    from ocbo_b200.cstrats import cstrats
    cfg = sc.CONTINUOUS[name]
    np.random.seed(sc.SEED)
    cls = [s.impl for s in cstrats if s.name == cfg['kind']][0]
    domain = [[0, 1] for _ in range(6)]
    got = sc.run_continuous(cfg, cls, synth.hartmann6, domain, sc.continuous_options(cfg, 'b200'))
    ref = oracle_engine[name]
    assert np.allclose(got['points'], ref['points'], rtol=0, atol=1e-12)
    assert np.allclose(got['regret'], ref['regret'], rtol=0, atol=1e-12)


def test_registry_and_names():
    from ocbo_b200.cstrats import cstrats
    from ocbo_b200.strategies import strategies
    # the reference's registries (src/strategies/__init__.py:13-27, src/cstrats/__init__.py:12-15)
    assert {s.name for s in strategies} == {'mts', 'joint-mts', 'mei', 'joint-mei', 'agn-ei', 'agn-thomp',
                                            'ja-thomp', 'ja-ei', 'Random', 'joint-rand', 'revi'}
    assert {s.name for s in cstrats} == {'cmts', 'cmts-pm', 'pei', 'rand', 'ei', 'ts', 'revi'}
    for s in strategies:
        assert s.impl.get_opt_method_name() == s.name
    for s in cstrats:
        assert s.impl.get_strat_name() == s.name


def test_unknown_engine_raises_value_error():
    from ocbo_b200.gp.gp_util import get_gp_fitter
    with pytest.raises(ValueError):
        get_gp_fitter('gpytorch', [[0.0]], [0.0], None)


LONG_GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'mts_long_traces.json')


@pytest.mark.parametrize('name,steps', [('set2d-full', 100), ('rand6d-full', 40), ('jointbran-full', 12)])
def test_teacher_forced_long_traces_on_the_oracle_engine(oracle_engine, name, steps):
    """The option-file-length traces (tests/scenarios.py LONG_*) with ocbo_b200's strategy classes on
    the numpy engine, teacher forced: same engine as the golden run, so every decision, every
    hyper-parameter digest and the bookkeeping must agree exactly -- no tie is needed."""
    from ocbo_b200.strategies import strategies
    with open(LONG_GOLDEN) as fh:
        ref = json.load(fh)[name]
    cfg = dict(sc.LONG_DISCRETE[name], capital=steps)
    ref = {k: (v[:steps] if isinstance(v, list) and k != 'total_best' else v) for k, v in ref.items()}
    np.random.seed(sc.SEED)
    cls = [s.impl for s in strategies if s.name == cfg['kind']][0]
    got = sc.run_discrete_long(cfg, cls, _functions(cfg), sc.discrete_options(cfg, 'b200'), ref=ref)
    stats = sc.compare_long(got, ref, atol=1e-12, hp_atol=1e-12)
    assert stats['steps_compared'] == steps
    assert stats['steps_with_a_different_pick'] == 0, stats
    assert stats['steps_with_different_hyper_parameters'] == 0, stats


def test_teacher_forced_continuous_trace_on_the_oracle_engine(oracle_engine):
    from ocbo_b200 import synth
    from ocbo_b200.cstrats import cstrats
    steps = 25
    with open(LONG_GOLDEN) as fh:
        ref = json.load(fh)['conth42-full-cmts']
    cfg = dict(sc.LONG_CONTINUOUS['conth42-full-cmts'], capital=steps)
    ref = {k: (v[:steps] if isinstance(v, list) and k != 'y_init' else v) for k, v in ref.items()}
    np.random.seed(sc.SEED)
    cls = [s.impl for s in cstrats if s.name == 'cmts'][0]
    got = sc.run_continuous_long(cfg, cls, synth.hartmann6, [[0, 1] for _ in range(6)],
                                 sc.continuous_options(cfg, 'b200'), ref=ref)
    stats = sc.compare_long(got, ref, atol=1e-12, hp_atol=1e-12)
    assert stats['steps_with_a_different_pick'] == 0, stats
    assert stats['steps_with_different_hyper_parameters'] == 0, stats
    assert np.allclose(got['regret'], ref['regret'][:steps], rtol=0, atol=1e-12)
