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
