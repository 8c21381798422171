"""End-to-end MTS selections: ocbo_b200's strategies on the B200 engine against
golden traces produced by the REFERENCE's own strategy classes on the numpy
oracle (tests/golden/make_golden.py), same global seed, same consumption of the
numpy stream.  Picks must match exactly apart from DOCUMENTED TIES (north_star);
the tie rule and its tolerance are defined in tests/scenarios.py and the margins
are stored with the golden traces."""
import json
import os

import numpy as np
import pytest

from tests import scenarios as sc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'mts_traces.json')


@pytest.fixture(scope='module')
def golden():
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
def test_discrete_mts_reproduces_reference_selections(golden, name):
    from ocbo_b200.gp.gp_interface import B200GPFitter
    from ocbo_b200.strategies import strategies
    cfg = sc.DISCRETE[name]
    np.random.seed(sc.SEED)
    B200GPFitter._fit_counter = 0
    f_infos = _functions(cfg)
    cls = [s.impl for s in strategies if s.name == cfg['kind']][0]
    got = sc.run_discrete(cfg, cls, f_infos, sc.discrete_options(cfg, 'b200'))
    ref = golden[name]
    k = sc.assert_same_up_to_ties(got['points'], ref)
    upto = len(ref['points']) if k is None else k
    assert got['choice_timeline'][:upto] == ref['choice_timeline'][:upto]
    assert np.allclose(got['values'][:upto], ref['values'][:upto], rtol=0, atol=1e-9)
    assert np.allclose(got['total_best'][:upto], ref['total_best'][:upto], rtol=0, atol=1e-9)
    if name == 'rand6d':
        assert k is None  # well-conditioned fixture (rung sensitivity ~1e-4): no excuse


@pytest.mark.parametrize('name', sorted(sc.CONTINUOUS))
def test_continuous_mts_reproduces_reference_selections(golden, name):
    from ocbo_b200 import synth
    from ocbo_b200.cstrats import cstrats
    from ocbo_b200.gp.gp_interface import B200GPFitter
    cfg = sc.CONTINUOUS[name]
    np.random.seed(sc.SEED)
    B200GPFitter._fit_counter = 0
    cls = [s.impl for s in cstrats if s.name == cfg['kind']][0]
    domain = [[0, 1] for _ in range(6)]
    got = sc.run_continuous(cfg, cls, synth.hartmann6, domain, sc.continuous_options(cfg, 'b200'))
    ref = golden[name]
    k = sc.assert_same_up_to_ties(got['points'], ref)
    upto = len(ref['points']) if k is None else k
    assert np.allclose(got['rewards'][:upto], ref['rewards'][:upto], rtol=0, atol=1e-9)
    assert np.allclose(got['regret'][:upto], ref['regret'][:upto], rtol=0, atol=1e-9)


def test_reference_strategy_classes_drive_the_b200_engine():
    """Drop-in at the GP seam: a strategy written against the reference's
    GPWrapper API (per-task eval / draw_sample loop, as src/strategies/mts.py does)
    gets the same answers from B200GP as from the oracle wrapper."""
    from oracle import gp_numpy as o
    from ocbo_b200 import rng
    from ocbo_b200.gp.gp_interface import make_gp
    from tests.util import problem
    X, y, hp = problem(3, 25, 2, bw=0.05)
    og = o.make_gp(X, y, 'se', 1.0, hp['bandwidths'], 1e-2, 0.0)
    bg = make_gp(X, y, 'se', 1.0, hp['bandwidths'], 1e-2, 0.0)
    rs = np.random.RandomState(0)
    pts = np.vstack([X, rs.uniform(size=(100, 2))])
    z = rs.normal(size=(len(pts), 1))
    o.NORMAL_SOURCE = lambda size: z.reshape(size)
    rng.NORMAL_SOURCE = lambda size: z.reshape(size)
    try:
        for gp in (og, bg):
            gp.build_posterior()
        so, sb = og.draw_sample(samp_pts=pts).ravel(), bg.draw_sample(samp_pts=pts).ravel()
    finally:
        o.NORMAL_SOURCE = None
        rng.NORMAL_SOURCE = None
    assert int(np.argmax(sb[25:])) == int(np.argmax(so[25:]))
    assert abs((sb[25:].max() - sb[:25].max()) - (so[25:].max() - so[:25].max())) <= 1e-6
    bg.add_data_single(pts[30], 0.3)
    og.add_data_single(pts[30], 0.3)
    assert bg.gp_core.alpha is None  # rebuilt lazily
    bg.build_posterior()
    assert np.max(np.abs(bg.gp_core.alpha - og.gp_core.alpha)) <= 1e-8 * np.max(np.abs(og.gp_core.alpha))
    assert set(bg.get_kernel_hps()) == set(og.get_kernel_hps()) == {'scale', 'lengthscale'}
    assert bg.get_estimated_noise() == og.get_estimated_noise() == 1e-2


EXACT_DISCRETE = ['jointbran-revi', 'set2d-agn-ei', 'jointbran-ja-ei', 'set2d-random', 'jointbran-rand']
EXACT_CONTINUOUS = ['conth42-revi', 'conth42-rand', 'conth42-agn-ei']


@pytest.mark.parametrize('name', EXACT_DISCRETE)
def test_sampling_free_methods_reproduce_reference_selections(golden, name):
    """REVI (knowledge gradient on the device) and the sampling-free baselines draw no normals:
    their selections on the B200 engine must equal the reference's own classes' step for step."""
    from ocbo_b200.gp.gp_interface import B200GPFitter
    from ocbo_b200.strategies import strategies
    cfg = sc.DISCRETE_BASELINES[name]
    np.random.seed(sc.SEED)
    B200GPFitter._fit_counter = 0
    cls = [s.impl for s in strategies if s.name == cfg['kind']][0]
    got = sc.run_discrete(cfg, cls, _functions(cfg), sc.discrete_options(cfg, 'b200'))
    ref = golden[name]
    assert got['choice_timeline'] == ref['choice_timeline']
    assert np.allclose(got['points'], ref['points'], rtol=0, atol=1e-12)
    assert np.allclose(got['total_best'], ref['total_best'], rtol=0, atol=1e-9)


@pytest.mark.parametrize('name', EXACT_CONTINUOUS)
def test_sampling_free_continuous_methods_reproduce_reference_selections(golden, name):
    from ocbo_b200 import synth
    from ocbo_b200.cstrats import cstrats
    from ocbo_b200.gp.gp_interface import B200GPFitter
    cfg = sc.CONTINUOUS_BASELINES[name]
    np.random.seed(sc.SEED)
    B200GPFitter._fit_counter = 0
    cls = [s.impl for s in cstrats if s.name == cfg['kind']][0]
    got = sc.run_continuous(cfg, cls, synth.hartmann6, [[0, 1] for _ in range(6)],
                            sc.continuous_options(cfg, 'b200'))
    ref = golden[name]
    assert np.allclose(got['points'], ref['points'], rtol=0, atol=1e-12)
    assert np.allclose(got['regret'], ref['regret'], rtol=0, atol=1e-9)


# ---------------------------------------------------------------------------
# Option-file-length traces, teacher forced (tests/scenarios.py "LONG_*")
# ---------------------------------------------------------------------------
LONG_GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'mts_long_traces.json')
STATS_OUT = os.path.join(os.path.dirname(os.path.dirname(__file__)), 'gpurun_out', 'long_trace_stats.json')


@pytest.fixture(scope='module')
def long_golden():
    with open(LONG_GOLDEN) as fh:
        return json.load(fh)


def _save_stats(name, stats):
    try:
        os.makedirs(os.path.dirname(STATS_OUT), exist_ok=True)
        allstats = {}
        if os.path.exists(STATS_OUT):
            with open(STATS_OUT) as fh:
                allstats = json.load(fh)
        allstats[name] = stats
        with open(STATS_OUT, 'w') as fh:
            json.dump(allstats, fh, indent=1)
    except OSError:
        pass
    print('\n[long trace] %s: %s' % (name, json.dumps(stats)))


def _mts_rungs(method):
    return [g.gp_core.last_sample_jitter for g in method.gps]


def _joint_rungs(method):
    return [method.gps[0].gp_core.last_sample_jitter]


def _profile_rungs(method):
    j = method.gp.gp_core.last_sample_jitter
    return list(j) if isinstance(j, (list, tuple)) else [j]


@pytest.mark.parametrize('name', sorted(sc.LONG_DISCRETE))
def test_option_file_length_discrete_traces(long_golden, name):
    """set2d (T=5, 100 steps), jointbran (T=10, 110 steps), rand6d (T=30, 300 steps): every
    decision of the B200 strategies against the reference's own classes, continuing past
    accepted ties; a differing pick must be a documented tie.  The statistics are written to
    gpurun_out/long_trace_stats.json (DESIGN.md section 2 quotes them)."""
    from ocbo_b200.gp.gp_interface import B200GPFitter
    from ocbo_b200.strategies import strategies
    cfg = sc.LONG_DISCRETE[name]
    np.random.seed(sc.SEED)
    B200GPFitter._fit_counter = 0
    f_infos = _functions(cfg)
    cls = [s.impl for s in strategies if s.name == cfg['kind']][0]
    ref = long_golden[name]
    got = sc.run_discrete_long(cfg, cls, f_infos, sc.discrete_options(cfg, 'b200'), ref=ref,
                               rungs=_joint_rungs if cfg['kind'] == 'joint-mts' else _mts_rungs)
    stats = sc.compare_long(got, ref)
    _save_stats(name, stats)
    assert not stats['undocumented'], stats
    assert stats['steps_with_different_hyper_parameters'] == 0, stats
    # teacher forced: both runs evaluated the same queries, so the bookkeeping agrees throughout
    assert np.allclose(got['total_best'], ref['total_best'], rtol=0, atol=1e-9)


@pytest.mark.parametrize('name', sorted(sc.LONG_CONTINUOUS))
def test_option_file_length_continuous_traces(long_golden, name):
    """conth42 (50 contexts x 100 actions per step, 200 steps) for cmts and cmts-pm."""
    from ocbo_b200 import synth
    from ocbo_b200.cstrats import cstrats
    from ocbo_b200.gp.gp_interface import B200GPFitter
    cfg = sc.LONG_CONTINUOUS[name]
    np.random.seed(sc.SEED)
    B200GPFitter._fit_counter = 0
    cls = [s.impl for s in cstrats if s.name == cfg['kind']][0]
    ref = long_golden[name]
    got = sc.run_continuous_long(cfg, cls, synth.hartmann6, [[0, 1] for _ in range(6)],
                                 sc.continuous_options(cfg, 'b200'), ref=ref, rungs=_profile_rungs)
    stats = sc.compare_long(got, ref)
    _save_stats(name, stats)
    assert not stats['undocumented'], stats
    assert stats['steps_with_different_hyper_parameters'] == 0, stats
    assert np.allclose(got['regret'], ref['regret'], rtol=0, atol=1e-8)
