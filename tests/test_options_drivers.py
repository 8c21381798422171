"""Option handling and drivers (host logic, no GPU unless marked)."""
import os

import numpy as np
import pytest

REF_OPTS = '/root/reference/src/options'


def test_option_specs_types_and_file(tmp_path):
    from ocbo_b200.options import get_option_specs, load_options
    specs = [get_option_specs('trials', False, 10, ''), get_option_specs('risk_neutral', False, False, ''),
             get_option_specs('description', False, None, ''), get_option_specs('scale', False, 0.5, ''),
             get_option_specs('methods', False, None, '')]
    f = tmp_path / 'o.txt'
    f.write_text('--trials 3\n--risk_neutral 1\n--description "two words"\n--methods a,b\n')
    o = load_options(specs, cmd_line=True, argv=['--options', str(f), '--scale', '2'])
    assert o.trials == 3 and o.risk_neutral is True and o.description == 'two words'
    assert o.scale == 2.0 and o.methods == 'a,b'
    d = load_options(specs, cmd_line=False)
    assert d.trials == 10 and d.description is None


@pytest.mark.refsrc
@pytest.mark.parametrize('name', ['set2d.txt', 'jointbran.txt', 'rand6d.txt'])
def test_reference_discrete_option_files_parse_unchanged(name):
    from ocbo_b200 import ocbo
    from ocbo_b200.options import load_options
    from ocbo_b200.strategies.multi_opt import multi_opt_args
    o = load_options(ocbo.docbo_args + multi_opt_args + ocbo.rand_fs_args, cmd_line=True,
                     argv=['--options', os.path.join(REF_OPTS, name)])
    assert o.run_id == name[:-4] and o.trials == 10 and o.tune_every == 1
    assert 'mts' in o.methods
    np.random.seed(0)
    f_infos = ocbo.assemble(o)
    assert len(f_infos) == {'set2d.txt': 5, 'jointbran.txt': 10, 'rand6d.txt': 30}[name]


@pytest.mark.refsrc
def test_reference_continuous_option_file_parses_unchanged():
    from ocbo_b200 import cts_ocbo
    from ocbo_b200.cstrats.cts_opt import cts_opt_args
    from ocbo_b200.cstrats.profile_cts import prof_args
    from ocbo_b200.options import load_options
    from ocbo_b200.cstrats.postmax_cts import pm_args
    o = load_options(cts_ocbo.ocbo_args + cts_opt_args + prof_args + pm_args, cmd_line=True,
                     argv=['--options', os.path.join(REF_OPTS, 'conth42.txt')])
    assert o.run_id == 'conth42' and o.max_capital == 750 and o.act_dim == 2
    assert cts_ocbo.find_function(o.function).name == 'hartmann6'


@pytest.mark.refsrc
def test_every_reference_option_file_parses_and_assembles():
    """All 15 option files of the reference (src/options/*.txt) load through the drivers'
    option specs and assemble their task functions (f4: callers either side of the path)."""
    from ocbo_b200 import cts_ocbo, ocbo
    from ocbo_b200.cstrats.cts_opt import cts_opt_args
    from ocbo_b200.cstrats.profile_cts import prof_args
    from ocbo_b200.options import load_options
    from ocbo_b200.strategies.multi_opt import multi_opt_args
    expect = {'jointbran.txt': 10, 'jointh22.txt': 9, 'jointh31.txt': 8, 'jointh42.txt': 16,
              'rand4d.txt': 30, 'rand6d.txt': 30, 'set2d.txt': 5}
    names = sorted(os.listdir(REF_OPTS))
    assert len(names) == 15
    for name in names:
        path = os.path.join(REF_OPTS, name)
        if name.startswith('cont'):
            from ocbo_b200.cstrats.postmax_cts import pm_args
            o = load_options(cts_ocbo.ocbo_args + cts_opt_args + prof_args + pm_args, cmd_line=True,
                             argv=['--options', path])
            f = cts_ocbo.find_function(o.function)
            assert len(f.domain) - o.act_dim >= 1 and 'cmts' in o.methods
        else:
            o = load_options(ocbo.docbo_args + multi_opt_args + ocbo.rand_fs_args, cmd_line=True,
                             argv=['--options', path])
            np.random.seed(0)
            assert len(ocbo.assemble(o)) == expect[name]


@pytest.mark.refsrc
def test_synthetic_functions_equal_the_reference_bit_for_bit():
    """ocbo_b200.synth against the reference's own src/synth (imported where it lies)."""
    from oracle import refshim
    refshim.install()
    from synth import continuous_2d as c2, sixd, twod
    from ocbo_b200 import synth
    rs = np.random.RandomState(0)
    for ref_f, our_f, d in ((twod.hartmann4, synth.hartmann4, 4), (twod.hartmann3, synth.hartmann3, 3),
                            (sixd.hartmann6, synth.hartmann6, 6), (twod.branin, synth.branin, 2)):
        assert all(ref_f(x) == our_f(x) for x in rs.uniform(size=(500, d)))
    for ref_set, our_set in ((c2.get_chopped_h42, synth.get_chopped_h42),
                             (c2.get_chopped_h22, synth.get_chopped_h22),
                             (c2.get_chopped_h31, synth.get_chopped_h31)):
        r, o = ref_set(), our_set()
        assert len(r) == len(o)
        for a, b in zip(r, o):
            assert a.name == b.name and [list(x) for x in a.domain] == [list(x) for x in b.domain]
            assert np.array_equal(np.asarray(a.f_loc, dtype=float), np.asarray(b.f_loc, dtype=float))
            assert all(a.function(x) == b.function(x) for x in rs.uniform(size=(10, len(a.domain))))


def test_strict_methods_raise_like_the_reference():
    from argparse import Namespace
    from ocbo_b200 import ocbo
    opts = Namespace(methods='kg-ucb', tunings=None, strict_methods=True)
    with pytest.raises(ValueError, match='Invalid method'):
        ocbo.assemble_methods([], opts)


@pytest.mark.gpu
def test_discrete_driver_runs_an_option_file(tmp_path):
    from ocbo_b200 import ocbo
    f = tmp_path / 'mini.txt'
    f.write_text('--run_id mini\n--trials 2\n--max_capital 3\n--tune_every 1\n--tuning_methods ml\n'
                 '--methods agn-thomp,mts,random\n--functions branin,parabaloid\n'
                 '--write_dir %s\n' % tmp_path)
    res = ocbo.run(argv=['--options', str(f)])
    assert len(res) == 2 and set(res[0]) == {'agn-thomp', 'mts', 'Random'}
    assert len(res[0]['mts'].choice_timeline) == 3
    assert os.path.exists(tmp_path / 'mini' / 'trial_1.pkl')


@pytest.mark.gpu
def test_continuous_driver_runs_an_option_file(tmp_path):
    from ocbo_b200 import cts_ocbo
    f = tmp_path / 'mini.txt'
    f.write_text('--run_id cmini\n--trials 1\n--max_capital 2\n--init_capital 40\n--methods cmts,cmts-pm,revi\n'
                 '--function hartmann6\n--act_dim 2\n--num_profiles 5\n--eval_ctx_fidel 4\n'
                 '--eval_act_fidel 4\n--judge_ctx_size 5\n--judge_act_size 10\n--cand_size 20\n'
                 '--write_dir %s\n' % tmp_path)
    res = cts_ocbo.run(argv=['--options', str(f)])
    assert set(res[0]) == {'cmts', 'cmts-pm', 'revi'}
    # the reference's file set (src/cts_ocbo.py:63-78)
    assert sorted(os.listdir(tmp_path / 'cmini')) == ['eval_set.pkl', 'function_info.pkl', 'rand_state.pkl',
                                                      'run_info.pkl', 'trial_0.pkl']
    assert len(res[0]['cmts'].query_history) == 2 and len(res[0]['cmts'].score_history) == 2


@pytest.mark.gpu
def test_every_method_of_the_option_files_runs(tmp_path):
    """set2d / jointbran / conth42 method lists: the MTS family, the EI family, the agnostic and
    random baselines and REVI all run through the same drivers."""
    from ocbo_b200 import cts_ocbo, ocbo
    f = tmp_path / 'a.txt'
    f.write_text('--run_id all_d\n--trials 1\n--max_capital 4\n--tune_every 1\n--tuning_methods ml\n'
                 '--methods agn-thomp,agn-ei,mts,mei,random\n--functions branin,parabaloid\n')
    res = ocbo.run(argv=['--options', str(f)])
    assert set(res[0]) == {'agn-thomp', 'agn-ei', 'mts', 'mei', 'Random'}
    f.write_text('--run_id all_j\n--trials 1\n--max_capital 4\n--init_capital 3\n--tune_every 1\n'
                 '--tuning_methods ml\n--methods joint-mts,joint-mei,ja-thomp,ja-ei,joint-rand,revi\n'
                 '--cts_f branin\n--risk_neutral 1\n--num_chops 4\n--run_ei_after 3\n--num_eval_pts 40\n')
    res = ocbo.run(argv=['--options', str(f)])
    assert set(res[0]) == {'joint-mts', 'joint-mei', 'ja-thomp', 'ja-ei', 'joint-rand', 'revi'}
    assert all(len(h.choice_timeline) == 4 for h in res[0].values())
    f.write_text('--run_id all_c\n--trials 1\n--max_capital 3\n--methods cmts,cmts-pm,pei,revi,rand\n'
                 '--function branin\n--act_dim 1\n--num_profiles 6\n--eval_ctx_fidel 5\n--eval_act_fidel 5\n'
                 '--judge_act_size 20\n--judge_ctx_size 10\n')
    res = cts_ocbo.run(argv=['--options', str(f)])
    assert set(res[0]) == {'cmts', 'cmts-pm', 'pei', 'rand', 'revi'}
    assert all(len(h.query_history) == 3 for h in res[0].values())


@pytest.mark.gpu
@pytest.mark.parametrize('chop,T', [('chop_h2_2', 9), ('chop_h3_1', 8), ('chop_h4_2', 16)])
def test_joint_driver_runs_the_chopped_hartmann_task_sets(tmp_path, chop, T):
    """jointh22 / jointh31 / jointh42.txt: joint task x action GP over the chopped Hartmann sets."""
    from ocbo_b200 import ocbo
    f = tmp_path / 'mini.txt'
    f.write_text('--run_id j_%s\n--trials 1\n--max_capital 3\n--init_capital 6\n--tune_every 1\n'
                 '--tuning_methods ml\n--methods joint-mts,joint-mei,revi\n--num_eval_pts 250\n'
                 '--risk_neutral 1\n--%s 1\n' % (chop, chop))
    res = ocbo.run(argv=['--options', str(f)])
    assert set(res[0]) == {'joint-mts', 'joint-mei', 'revi'}
    tl = res[0]['joint-mts'].choice_timeline
    assert len(tl) == 3 and all(0 <= c < T for c in tl)


@pytest.mark.gpu
@pytest.mark.parametrize('extra', ['', '--hp_tune_samps 30\n'])
def test_continuous_driver_runs_hartmann4(tmp_path, extra):
    """conth22.txt / conth22_sethps.txt shapes (2 contexts + 2 actions), with and without
    pre-tuned hyper-parameters."""
    from ocbo_b200 import cts_ocbo
    f = tmp_path / 'mini.txt'
    f.write_text('--run_id c_h4\n--trials 1\n--max_capital 3\n--methods cmts,cmts-pm,pei\n'
                 '--function hartmann4\n--act_dim 2\n--num_profiles 6\n--eval_ctx_fidel 5\n'
                 '--eval_act_fidel 5\n' + extra)
    res = cts_ocbo.run(argv=['--options', str(f)])
    assert set(res[0]) == {'cmts', 'cmts-pm', 'pei'}
    assert len(res[0]['cmts'].query_history) == 3


def test_vectorised_objective_is_bitwise_equal():
    """synth.hartmann6.batch (used by the continuous drivers' policy score) equals the per-point
    function bit for bit."""
    from ocbo_b200 import synth
    X = np.random.RandomState(3).uniform(size=(5000, 6))
    assert np.array_equal(synth.hartmann6.batch(X), np.asarray([synth.hartmann6(x) for x in X]))


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['pretune_for_joint', 'tune_every_0'])
def test_discrete_driver_pre_tuned_modes(tmp_path, mode):
    """src/ocbo.py:146-159: ``--pretune_for_joint`` (one shared empty-data joint GP with tuned
    hyper-parameters, tune_every forced to 0) and ``--tune_every 0`` (a tuned empty prior per
    task): the GPs start without data and grow a point at a time, nothing is re-fitted."""
    from ocbo_b200 import ocbo
    from ocbo_b200.gp import hp_search
    f = tmp_path / 'pre.txt'
    if mode == 'pretune_for_joint':
        f.write_text('--run_id pre\n--trials 2\n--max_capital 3\n--init_capital 3\n--pretune_for_joint 1\n'
                     '--methods joint-mts,joint-mei\n--cts_f branin\n--num_chops 3\n--risk_neutral 1\n')
        names = {'joint-mts', 'joint-mei'}
    else:
        f.write_text('--run_id pre\n--trials 2\n--max_capital 3\n--init_capital 3\n--tune_every 0\n'
                     '--methods mts,mei\n--functions branin,parabaloid\n')
        names = {'mts', 'mei'}
    fits = []
    real = hp_search.batched_lml

    def counting(*a, **k):
        fits.append(1)
        return real(*a, **k)
    hp_search.batched_lml = counting
    try:
        res = ocbo.run(argv=['--options', str(f)])
        n_fits = len(fits)
        res_fl = ocbo.run(argv=['--options', str(f), '--trials_in_flight', '2'])
    finally:
        hp_search.batched_lml = real
    assert n_fits == (1 if mode == 'pretune_for_joint' else 2)  # the prior fits only, none per step
    assert len(res) == 2 and set(res[0]) == names and set(res_fl[1]) == names
    for r in res + res_fl:
        for h in r.values():
            assert len(h.choice_timeline) == 3 and np.all(np.isfinite(h.total_best))


@pytest.mark.refsrc
def test_results_feed_the_reference_analysis_functions(tmp_path, monkeypatch):
    """the per-trial results keep the reference's Namespace schema -- its own
    analysis helpers (src/util/post_util.py:32-90, imported where they lie) assemble them.  The
    run uses the numpy oracle engine behind the driver (host logic only, no GPU)."""
    import importlib.util
    import pickle
    from ocbo_b200 import batched, ocbo
    from ocbo_b200.strategies import joint_opt, multi_opt
    from oracle import gp_numpy as o
    from tests.test_host_strategies import _Fitter, _fake_mts_step
    spec = importlib.util.spec_from_file_location('ref_post_util', '/root/reference/src/util/post_util.py')
    post_util = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(post_util)
    fitter = lambda engine, x, y, opts: _Fitter(x, y, opts)
    monkeypatch.setattr(multi_opt, 'get_gp_fitter', fitter)
    monkeypatch.setattr(joint_opt, 'get_gp_fitter', fitter)
    monkeypatch.setattr(batched, '_mts_flat', _fake_mts_step)
    o.EuclideanGPFitter._fit_counter = 0
    f = tmp_path / 'sch.txt'
    f.write_text('--run_id sch\n--trials 3\n--max_capital 4\n--tune_every 1\n--tuning_methods ml\n'
                 '--methods mts,random\n--functions branin,parabaloid\n--write_dir %s\n' % tmp_path)
    ocbo.run(argv=['--options', str(f)])
    data = []
    for t in range(3):  # (post_util.load_results opens the pickles in text mode: a python-2 idiom)
        with open(tmp_path / 'sch' / ('trial_%d.pkl' % t), 'rb') as fh:
            data.append(pickle.load(fh))
    with open(tmp_path / 'sch' / 'functions.pkl', 'rb') as fh:
        f_infos = pickle.load(fh)
    totals = post_util.get_total_best_mats(data)
    assert set(totals) == {'mts', 'Random'} and totals['mts'].shape == (3, 4)
    assert post_util.get_choice_timelines(data)['mts'].shape == (3, 4)
    assert f_infos[0].function is None and f_infos[0].max_val is not None  # (src/ocbo.py:229-235)
    name = f_infos[0].name
    assert len(post_util.get_best_for_function(name, data)['mts']) == 3
