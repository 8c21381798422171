"""World-size-2 gloo test of the trial sharding + single gather (no GPU): the
gathered picks must be indexed by global trial id and identical to the
world-size-1 result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_picks(trial, S):
    rs = np.random.RandomState(trial)
    return rs.randint(0, 64, S), rs.randint(0, 128, S), rs.normal(size=S)


def _worker(rank, world, port, num_trials, S, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ocbo_b200.dist import gather_picks, shard_trials
    mine = shard_trials(num_trials)
    assert mine == list(range(rank, num_trials, world))
    t, a, g = zip(*[_fake_picks(tr, S) for tr in mine]) if mine else ((), (), ())
    res = gather_picks(mine, np.asarray(t).reshape(-1, S), np.asarray(a).reshape(-1, S),
                       np.asarray(g).reshape(-1, S), num_trials, device=torch.device('cpu'))
    if rank == 0:
        np.savez(out_path, task=res[0], action=res[1], gain=res[2])
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gather_matches_single_rank(tmp_path):
    num_trials, S = 7, 5
    out = str(tmp_path / 'w2.npz')
    mp.spawn(_worker, args=(2, _free_port(), num_trials, S, out), nprocs=2, join=True)
    got = np.load(out)
    ref = [_fake_picks(tr, S) for tr in range(num_trials)]
    assert np.array_equal(got['task'], np.asarray([r[0] for r in ref]))
    assert np.array_equal(got['action'], np.asarray([r[1] for r in ref]))
    assert np.array_equal(got['gain'], np.asarray([r[2] for r in ref]))


def test_single_process_gather_and_sharding():
    from ocbo_b200.dist import gather_picks, shard_trials
    assert shard_trials(10, 1, 4) == [1, 5, 9]
    assert sum(len(shard_trials(512, r, 8)) for r in range(8)) == 512
    t, a, g = gather_picks([2, 0, 1], np.arange(6).reshape(3, 2), np.arange(6).reshape(3, 2) + 10,
                           np.arange(6.0).reshape(3, 2), 3, device=torch.device('cpu'))
    assert t.tolist() == [[2, 3], [4, 5], [0, 1]] and a[0].tolist() == [12, 13]


# ---------------------------------------------------------------------------
# option-file driver, trials sharded over two ranks (gloo), numpy oracle engine
# ---------------------------------------------------------------------------
def _oracle_engine():
    """What tests/test_host_strategies.py's fixture does, without pytest (spawned workers)."""
    from ocbo_b200 import batched
    from ocbo_b200.strategies import multi_opt
    from oracle import gp_numpy as o
    from tests.test_host_strategies import _Fitter, _fake_mts_step
    multi_opt.get_gp_fitter = lambda engine, x, y, opts: _Fitter(x, y, opts)
    batched._mts_flat = _fake_mts_step
    o.EuclideanGPFitter._fit_counter = 0  # 'ml' only below: the criterion never alternates


_DRIVER_ARGV = ['--run_id', 'shard', '--trials', '5', '--max_capital', '3', '--tune_every', '1',
                '--tuning_methods', 'ml', '--methods', 'mts', '--functions', 'branin,parabaloid',
                '--trials_in_flight', '2']


def _summary(results):
    return [[int(c) for c in r['mts'].choice_timeline] + [float(v) for v in r['mts'].total_best]
            for r in results]


def _driver_worker(rank, world, port, out_path, write_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    _oracle_engine()
    from ocbo_b200 import ocbo
    res = ocbo.run(argv=_DRIVER_ARGV + ['--write_dir', write_dir])
    if rank == 0:
        import json
        with open(out_path, 'w') as fh:
            json.dump(_summary(res), fh)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_driver_shards_trials_over_two_ranks(tmp_path, monkeypatch):
    """``torchrun -m ocbo_b200.ocbo`` semantics on CPU: 5 trials, 2 in flight per rank, world size 2;
    rank 0 gathers every trial and writes trial_%d.pkl; the results equal the single-process run
    (a trial's stream is keyed by its global id)."""
    import json
    out = str(tmp_path / 'w2.json')
    mp.spawn(_driver_worker, args=(2, _free_port(), out, str(tmp_path / 'w2')), nprocs=2, join=True)
    with open(out) as fh:
        got = json.load(fh)
    assert sorted(os.listdir(tmp_path / 'w2' / 'shard'))[-5:] == ['trial_%d.pkl' % t for t in range(5)]
    from ocbo_b200 import batched, ocbo
    from ocbo_b200.strategies import multi_opt
    from oracle import gp_numpy as o
    from tests.test_host_strategies import _Fitter, _fake_mts_step
    monkeypatch.setattr(multi_opt, 'get_gp_fitter', lambda engine, x, y, opts: _Fitter(x, y, opts))
    monkeypatch.setattr(batched, '_mts_flat', _fake_mts_step)
    monkeypatch.delenv('WORLD_SIZE', raising=False)
    o.EuclideanGPFitter._fit_counter = 0
    ref = _summary(ocbo.run(argv=_DRIVER_ARGV))
    assert got == ref


def _stream_worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ocbo_b200.dist import share_numpy_stream
    np.random.seed(1000 + rank)          # an unseeded run: every rank starts from its own entropy
    share_numpy_stream()                 # ... until rank 0's stream is shared
    draws = np.random.uniform(size=5)
    np.save(os.path.join(out_dir, 'draws_%d.npy' % rank), draws)
    dist.barrier()
    dist.destroy_process_group()


def test_unseeded_ranks_share_rank_zeros_numpy_stream(tmp_path):
    """--set_seed 0 under a launcher: the drivers assemble their task functions, risk-neutral grids and
    evaluation sets from the global numpy stream on EVERY rank; they must all see rank 0's stream
    (ocbo_b200/ocbo.py, cts_ocbo.py -> dist.share_numpy_stream)."""
    mp.spawn(_stream_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    d0, d1 = np.load(str(tmp_path / 'draws_0.npy')), np.load(str(tmp_path / 'draws_1.npy'))
    assert np.array_equal(d0, d1)
    np.random.seed(1000)
    assert np.array_equal(d0, np.random.uniform(size=5))
