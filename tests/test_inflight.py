"""Trials in flight (ocbo_b200/inflight.py): the lockstep scheduler on the CPU.  A trial run
in a pool must see exactly what it would see alone -- its own numpy stream, its own
criterion counter, its own results -- while the requests of all trials reach the merged
pass together."""
import os

import numpy as np
import pytest

from ocbo_b200 import inflight
from tests import scenarios as sc


class _Counter(object):
    value = 0


def _toy_trial(steps, log):
    def run_many(payloads):
        log.append(len(payloads))
        return [float(np.sum(p)) for p in payloads]

    def trial():
        acc = []
        for _ in range(steps):
            x = np.random.uniform(size=3)
            _Counter.value += 1
            acc.append(inflight.merge('sum', x, run_many) + _Counter.value)
            acc.append(float(np.random.normal()))
        return acc
    return trial


def test_pool_isolates_streams_and_merges_requests():
    seeds = [11, 12, 13, 14]
    states = [np.random.RandomState(s).get_state() for s in seeds]
    alone = []
    for st in states:
        np.random.set_state(st)
        _Counter.value = 0
        alone.append(_toy_trial(5, [])())
    np.random.seed(99)
    before = np.random.get_state()
    _Counter.value = 0
    log = []
    pool = inflight.TrialPool([(_Counter, 'value')])
    got = pool.run([_toy_trial(5, log) for _ in seeds], states)
    assert got == alone
    assert log == [4] * 5 and pool.stats == {'sum': (5, 20)}
    # the caller's own stream and globals are untouched
    after = np.random.get_state()
    assert before[0] == after[0] and np.array_equal(before[1], after[1]) and before[2:] == after[2:]
    assert _Counter.value == 0


def test_pool_handles_trials_of_different_lengths_and_kinds():
    log = []

    def run_a(payloads):
        log.append(('a', len(payloads)))
        return [p + 1 for p in payloads]

    def run_b(payloads):
        log.append(('b', len(payloads)))
        return [p * 2 for p in payloads]

    def trial(k):
        def f():
            v = k
            for i in range(k + 1):
                v = inflight.merge('a', v, run_a)
                if i % 2:
                    v = inflight.merge('b', v, run_b)
            return v
        return f
    states = [np.random.RandomState(i).get_state() for i in range(3)]
    got = inflight.TrialPool().run([trial(k) for k in range(3)], states)
    want = [inflight.merge('a', 0, run_a), None, None]  # outside a pool: direct call
    for k in (1, 2):
        want[k] = trial(k)()
    assert got == want
    assert ('a', 3) in log  # the first request of all three trials ran as one pass


def test_errors_propagate():
    def bad_pass(payloads):
        raise ValueError('no device')

    def t_bad():
        return inflight.merge('x', 1, bad_pass)

    def t_raises():
        raise KeyError('trial bug')
    st = [np.random.RandomState(0).get_state()] * 2
    with pytest.raises(ValueError, match='no device'):
        inflight.TrialPool().run([t_bad, t_bad], st)
    with pytest.raises(KeyError):
        inflight.TrialPool().run([t_raises, lambda: 3], st)
    assert not inflight.in_flight()


def test_trial_states_and_waves():
    np.random.seed(5)
    st = inflight.trial_rng_states(3, 100)
    now = np.random.get_state()
    assert np.array_equal(st[0][1], now[1]) and st[0][2] == now[2]
    assert np.array_equal(st[2][1], np.random.RandomState(102).get_state()[1])
    res, stats = inflight.run_waves(5, 2, 7, lambda t: (lambda: (t, float(np.random.uniform()))))
    assert [r[0] for r in res] == [0, 1, 2, 3, 4]
    assert res[3][1] == float(np.random.RandomState(10).uniform())


@pytest.mark.parametrize('name', ['set2d', 'rand6d'])
def test_mts_trials_in_flight_equal_the_trials_alone(monkeypatch, name):
    """The MTS host mirror on the numpy oracle engine (as in test_host_strategies): three trials
    in flight -- their decisions merged into one ``_mts_flat`` call per step -- reproduce the three
    trials run one by one from the same stream states."""
    from ocbo_b200 import batched, synth
    from ocbo_b200.strategies import multi_opt, strategies
    from oracle import gp_numpy as o
    from tests.test_host_strategies import _Fitter, _fake_mts_step
    monkeypatch.setattr(multi_opt, 'get_gp_fitter', lambda engine, x, y, opts: _Fitter(x, y, opts))
    sizes = []

    def flat(gps, pts, n_old, Z):
        sizes.append(len(gps))
        return _fake_mts_step(gps, pts, n_old, Z)
    monkeypatch.setattr(batched, '_mts_flat', flat)
    cfg = dict(sc.DISCRETE[name], capital=4)
    np.random.seed(sc.SEED)
    if 'functions' in cfg:
        f_infos = synth.assemble_functions(cfg['functions'])
    else:
        f_infos = synth.random_mass_functions(*cfg['masses'], cfg['massf_dim'])
    cls = [s.impl for s in strategies if s.name == 'mts'][0]
    opts = sc.discrete_options(cfg, 'b200')
    states = [np.random.RandomState(40 + t).get_state() for t in range(3)]
    alone = []
    for st in states:
        np.random.set_state(st)
        o.EuclideanGPFitter._fit_counter = 0
        alone.append(sc.run_discrete(cfg, cls, f_infos, opts))
    T = len(f_infos)
    assert sizes == [T] * (3 * cfg['capital'])
    del sizes[:]
    o.EuclideanGPFitter._fit_counter = 0
    pool = inflight.TrialPool([(o.EuclideanGPFitter, '_fit_counter')])
    got = pool.run([lambda: sc.run_discrete(cfg, cls, f_infos, opts)] * 3, states)
    assert got == alone
    assert sizes == [3 * T] * cfg['capital']


def test_pack_rows_pads_ragged_inputs():
    """Staging of many ragged problems into one zero-padded block."""
    from ocbo_b200 import engine
    rs = np.random.RandomState(0)
    arrs = [rs.uniform(size=(n, 3)) for n in (5, 0, 130, 1, 77)]
    out, lens, n_pad = engine.pack_rows(arrs, 3)
    assert n_pad == 256 and list(lens) == [5, 0, 130, 1, 77] and out.shape == (5, 256, 3)
    for b, a in enumerate(arrs):
        assert np.array_equal(out[b, :len(a)], a) and not out[b, len(a):].any()
    out1, _, n1 = engine.pack_rows([np.zeros((0, 2))], 2)
    assert out1.shape == (1, 128, 2) and n1 == 128 and not out1.any()
    out2 = engine.pack_rows([a[:, :1] for a in arrs], 1, 384)[0]     # (non-contiguous inputs)
    assert out2.shape == (5, 384, 1) and np.array_equal(out2[2, :130, 0], arrs[2][:, 0])


def test_pack_rows_c_helper_equals_the_numpy_loop(monkeypatch):
    """csrc/hostpack.c (built by build()) and the fallback loop give the same block; inputs the
    helper does not take (strided, integer) go through the loop."""
    from ocbo_b200 import build, engine
    build.build_hostpack()
    helper = engine._load_hostpack()
    assert helper is not None
    rs = np.random.RandomState(1)
    cases = [[rs.uniform(size=(n, 4)) for n in (3, 0, 129, 40)],
             [rs.uniform(size=(7, 1)), rs.uniform(size=(2, 1))],
             [rs.uniform(size=(6, 8))[:, ::2], np.arange(8).reshape(2, 4)]]
    for arrs in cases:
        width = arrs[0].shape[1]
        monkeypatch.setattr(engine, '_HOSTPACK', helper)
        fast = engine.pack_rows(arrs, width)
        monkeypatch.setattr(engine, '_HOSTPACK', None)
        slow = engine.pack_rows(arrs, width)
        assert np.array_equal(fast[0], slow[0]) and fast[1:] == slow[1:]
    out = np.empty((2, 128, 4))
    assert helper.pack_rows(cases[0][:2], 4, 128, out) == [3, 0] and not out[0, 3:].any()
    with pytest.raises(TypeError):
        helper.pack_rows([np.arange(4)], 4, 128, np.empty((1, 128, 4)))
    with pytest.raises(ValueError):
        helper.pack_rows([np.zeros((200, 4))], 4, 128, np.empty((1, 128, 4)))


def _timeline(hist, names):
    pts = {n: [[float(v) for v in np.ravel(q.pt)] for q in hist.query_history[n]] for n in names}
    return [int(c) for c in hist.choice_timeline], pts, [float(v) for v in hist.total_best]


@pytest.mark.gpu
@pytest.mark.parametrize('methods,tuning', [('mts', 'ml'), ('mts,mei', 'ml-post_sampling')])
def test_discrete_driver_in_flight_equals_the_trials_alone(tmp_path, methods, tuning):
    """``--trials_in_flight 3`` through the option-file driver on the B200 engine: every trial's
    choices, points and values are those of the same trial run alone from the same stream state
    (merged MTS / MEI / LML passes give each problem the bits it gets in its own batch)."""
    from ocbo_b200 import ocbo
    from ocbo_b200.gp.gp_interface import B200GPFitter
    f = tmp_path / 'fl.txt'
    f.write_text('--run_id fl\n--trials 3\n--max_capital 6\n--tune_every 1\n--tuning_methods %s\n'
                 '--hp_samples 3\n--methods %s\n--functions branin,parabaloid,parabaloid\n'
                 '--write_dir %s\n' % (tuning, methods, tmp_path))
    argv = ['--options', str(f)]
    B200GPFitter._fit_counter = 0
    got = ocbo.run(argv=argv + ['--trials_in_flight', '3'])
    stats = ocbo.LAST_MERGE_STATS[0]
    assert os.path.exists(tmp_path / 'fl' / 'trial_2.pkl')
    n_methods = len(methods.split(','))
    assert stats['mts'] == (6, 18)                      # one merged decision per step for the 3 trials
    assert stats['lml'][1] == 3 * n_methods * (3 + 6)   # 3 initial fits + one refit per step, per trial
    assert stats['lml'][0] == n_methods * (3 + 6)
    # the same trials one by one, from the same states
    from ocbo_b200.options import load_options
    from ocbo_b200.strategies.multi_opt import multi_opt_args
    options = load_options(ocbo.docbo_args + multi_opt_args + ocbo.rand_fs_args, cmd_line=True, argv=argv)
    np.random.seed(ocbo.SEED)
    f_infos = ocbo.assemble(options)
    meths = ocbo.assemble_methods(f_infos, options)
    names = [fi.name for fi in f_infos]
    for t, st in enumerate(inflight.trial_rng_states(3, ocbo.SEED)):
        np.random.set_state(st)
        B200GPFitter._fit_counter = 0
        alone = ocbo.run_single(meths, f_infos, options)
        for key in alone:
            assert _timeline(got[t][key], names) == _timeline(alone[key], names), (t, key)


@pytest.mark.gpu
def test_continuous_driver_in_flight(tmp_path):
    """cts_ocbo with two trials in flight: the hyper-parameter fits and the policy-score evaluations
    merge; the profile steps (ONE posterior shared by 50 contexts, per trial) are enqueued together
    on their own streams; results equal the trials alone."""
    from ocbo_b200 import cts_ocbo
    from ocbo_b200.gp.gp_interface import B200GPFitter
    f = tmp_path / 'cfl.txt'
    f.write_text('--run_id cfl\n--trials 2\n--max_capital 3\n--init_capital 30\n--methods cmts\n'
                 '--function hartmann6\n--act_dim 2\n--num_profiles 5\n--eval_ctx_fidel 4\n'
                 '--eval_act_fidel 4\n')
    argv = ['--options', str(f)]
    B200GPFitter._fit_counter = 0
    got = cts_ocbo.run(argv=argv + ['--trials_in_flight', '2'])
    assert cts_ocbo.LAST_MERGE_STATS[0]['lml'][1] == 2 * cts_ocbo.LAST_MERGE_STATS[0]['lml'][0]
    assert cts_ocbo.LAST_MERGE_STATS[0]['cts'] == (3, 6)  # the trials' profile steps are enqueued together
    B200GPFitter._fit_counter = 0
    alone0 = cts_ocbo.run(argv=argv + ['--trials', '1'])
    a = [np.ravel(q.pt).tolist() + [q.reward] for q in got[0]['cmts'].query_history]
    b = [np.ravel(q.pt).tolist() + [q.reward] for q in alone0[0]['cmts'].query_history]
    assert a == b
    assert len(got[1]['cmts'].query_history) == 3


@pytest.mark.gpu
def test_joint_driver_in_flight_equals_the_trials_alone(tmp_path):
    """joint-MTS (jointbran.txt's method) with three trials in flight: the trials' joint posterior
    samples share one batched pass (batched.joint_many) and each trial reproduces itself alone."""
    from ocbo_b200 import ocbo
    from ocbo_b200.gp.gp_interface import B200GPFitter
    from ocbo_b200.options import load_options
    from ocbo_b200.strategies.multi_opt import multi_opt_args
    f = tmp_path / 'jfl.txt'
    f.write_text('--run_id jfl\n--trials 3\n--max_capital 5\n--init_capital 5\n--tune_every 1\n'
                 '--tuning_methods ml\n--methods joint-mts\n--cts_f branin\n--risk_neutral 1\n--num_chops 4\n')
    argv = ['--options', str(f)]
    B200GPFitter._fit_counter = 0
    got = ocbo.run(argv=argv + ['--trials_in_flight', '3'])
    stats = ocbo.LAST_MERGE_STATS[0]
    assert stats['joint'] == (5, 15)
    options = load_options(ocbo.docbo_args + multi_opt_args + ocbo.rand_fs_args, cmd_line=True, argv=argv)
    np.random.seed(ocbo.SEED)
    f_infos = ocbo.assemble(options)
    meths = ocbo.assemble_methods(f_infos, options)
    names = [fi.name for fi in f_infos]
    for t, st in enumerate(inflight.trial_rng_states(3, ocbo.SEED)):
        np.random.set_state(st)
        B200GPFitter._fit_counter = 0
        alone = ocbo.run_single(meths, f_infos, options)
        assert _timeline(got[t]['joint-mts'], names) == _timeline(alone['joint-mts'], names), t


@pytest.mark.gpu
def test_pipelined_chunks_equal_the_single_pass(monkeypatch):
    """A merged pass of >= 128 problems is split into chunks whose staging overlaps the device
    (batched._chunks); every problem gets the bits of the unsplit pass."""
    from ocbo_b200 import batched
    from ocbo_b200.gp.gp_interface import make_gp
    rs = np.random.RandomState(2)
    gps, pts, n_old, Z = [], [], [], []
    for k in range(150):
        n = 4 + k % 9
        X = rs.uniform(size=(n, 2))
        y = np.sin(3 * X.sum(1))
        gps.append(make_gp(X, y, 'se', 1.0, [0.3, 0.4], 1e-2, 0.1, build=False))
        pts.append(np.vstack([X, rs.uniform(size=(20 + k % 5, 2))]))
        n_old.append(n)
        Z.append(rs.normal(size=(len(pts[-1]), 1)))
    assert [len(c) for c in batched._chunks(list(range(150)))] == [75, 75]
    got = batched.mts_step(gps, pts, n_old, Z)
    again = batched.mts_step(gps, pts, n_old, Z)       # replayed plans
    monkeypatch.setattr(batched, 'PIPELINE_MIN', 10 ** 9)
    ref = batched.mts_step(gps, pts, n_old, Z)
    assert got == ref and again == ref


def _merged(kind_fn, payloads):
    """Run one merged pass the way the scheduler does (several requests at once)."""
    return kind_fn(payloads)


@pytest.mark.gpu
def test_merged_joint_and_mean_passes_fall_back_per_request():
    """joint_many / mean_many with several requests: a well-conditioned group equals the requests
    run alone; a group holding a problem that needs the host-continued ladder (duplicated
    candidates, vanishing noise) hands every request back to its own tuned path."""
    from ocbo_b200 import batched
    from ocbo_b200.gp.gp_interface import make_gp
    rs = np.random.RandomState(6)
    T, A = 3, 20

    def request(noise, duplicate):
        n = 12
        X = np.hstack([np.repeat(np.linspace(0, 1, T), n // T)[:, None], rs.uniform(size=(n, 1))])
        y = np.sin(4 * X.sum(1))
        gp = make_gp(X, y, 'se', 1.0, [0.5, 0.5], noise, 0.0, build=False)
        grid = np.hstack([np.repeat(np.linspace(0, 1, T), A)[:, None], rs.uniform(size=(T * A, 1))])
        if duplicate:
            grid[A:2 * A] = grid[:A]
        rows = np.vstack([X, grid])
        seg = np.concatenate([np.arange(0, n + 1, n // T), n + A * np.arange(1, T + 1)]).astype(np.int32)
        return gp.gp_core, rows, seg, T, rs.normal(size=(len(rows), 1))

    for dup in (False, True):
        reqs = [request(1e-2, False), request(1e-12 if dup else 1e-2, dup), request(1e-2, False)]
        alone = [c._sample_and_pick_alone(r, s, t, z) for c, r, s, t, z in reqs]
        jit = [c.last_sample_jitter for c, _, _, _, _ in reqs]
        for c, _, _, _, _ in reqs:
            c._post = None
        assert _merged(batched.joint_many, reqs) == alone
        assert [c.last_sample_jitter for c, _, _, _, _ in reqs] == jit
    cores = [r[0] for r in reqs]
    pts = [rs.uniform(size=(30 + 5 * k, 2)) for k in range(3)]
    means_alone = [c._mean_alone(p) for c, p in zip(cores, pts)]
    merged = _merged(batched.mean_many, list(zip(cores, pts)))
    assert all(np.array_equal(a, b) for a, b in zip(means_alone, merged))


@pytest.mark.gpu
def test_merged_lml_group_equals_single_passes(monkeypatch):
    """_lml_many over three data sets of different sizes == three single passes, including a
    candidate that needs the jitter ladder."""
    from ocbo_b200 import engine
    from ocbo_b200.gp import hp_search
    rs = np.random.RandomState(8)
    jobs = []
    for n in (20, 57, 130):
        X = rs.uniform(size=(n, 3))
        if n == 57:
            X[10:20] = X[:10]
        y = np.sin(4 * X.sum(1))
        th = hp_search.hp_candidates(X, y, 8, uniforms=rs.uniform(size=(7, 6)))
        if n == 57:
            th[3, 1] = 1e-18
        jobs.append((X, y, 0, th))
    single = [hp_search._lml_single(*j)[0] for j in jobs]
    for graphs in (True, False):
        monkeypatch.setattr(engine, 'graphs_enabled', lambda g=graphs: g)
        merged = hp_search._lml_many(jobs)
        for a, (b, _) in zip(single, merged):
            assert np.array_equal(a, b)


@pytest.mark.gpu
def test_driver_processes_share_one_gpu(tmp_path):
    """``torchrun --nproc-per-node 2`` on ONE GPU: both ranks use the device (time-sliced), the
    trials are sharded over them and gathered through gloo; the pickles equal the single-process
    run's (a trial's stream is keyed by its global id)."""
    import pickle
    import subprocess
    import sys
    import torch
    from ocbo_b200 import ocbo
    from ocbo_b200.gp.gp_interface import B200GPFitter
    if torch.cuda.device_count() != 1:
        pytest.skip('needs exactly one visible GPU')
    f = tmp_path / 'pp.txt'
    f.write_text('--run_id pp\n--trials 4\n--max_capital 4\n--tune_every 1\n--tuning_methods ml\n'
                 '--methods mts\n--functions branin,parabaloid\n--trials_in_flight 2\n--write_dir %s\n' % tmp_path)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root)
    subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                    '--master-addr', '127.0.0.1', '--master-port', '29533', '-m', 'ocbo_b200.ocbo',
                    '--options', str(f)], check=True, env=env, cwd=root, timeout=240)
    B200GPFitter._fit_counter = 0
    ref = ocbo.run(argv=['--options', str(f), '--write_dir', str(tmp_path / 'one')])
    for t in range(4):
        with open(tmp_path / 'pp' / ('trial_%d.pkl' % t), 'rb') as fh:
            got = pickle.load(fh)['mts']
        assert got.choice_timeline == ref[t]['mts'].choice_timeline
        assert got.total_best == ref[t]['mts'].total_best
