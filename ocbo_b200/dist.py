"""Multi-GPU plumbing: one process per GPU, independent trials sharded with no
collective inside a step, ONE gather of the per-trial picks at the end
(SURVEY.md 8e).  torch.distributed (NCCL on the box, gloo in the CPU tests) is
plumbing only."""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env():
    """Join the process group a launcher (torchrun) described in the environment; a no-op for a
    plain ``python`` run.  Returns (rank, world size).

    One process per GPU is the rule for the sweep.  The option-file drivers may be launched
    with MORE processes than GPUs (``--nproc-per-node 8`` on one B200): their loops are bound by
    the per-trial Python of the strategy code, which only more processes run in parallel, and
    their short device passes time-slice the GPU (measured: rand6d 0.70 k -> 4.3 k iterations/s
    with 8 processes on one B200).  Local rank r then uses GPU r mod #GPUs, and the one gather of
    results goes through gloo (NCCL does not accept two ranks on one device)."""
    import os
    if dist.is_available() and not dist.is_initialized() and int(os.environ.get('WORLD_SIZE', '1')) > 1:
        backend = 'gloo'
        if torch.cuda.is_available():
            ndev = torch.cuda.device_count()
            local = int(os.environ.get('LOCAL_RANK', '0'))
            per_node = int(os.environ.get('LOCAL_WORLD_SIZE', os.environ['WORLD_SIZE']))
            torch.cuda.set_device(local % ndev)
            if per_node <= ndev:
                backend = 'nccl'
        dist.init_process_group(backend)
    return world()


def share_numpy_stream():
    """Give every rank rank 0's global numpy stream.  The drivers build their task functions,
    risk-neutral grids and evaluation sets from that stream before the trials are sharded; with
    ``--set_seed 0`` the ranks' streams start from different entropy, and each rank would optimise
    different random functions while rank 0 writes its own functions.pkl next to everybody's
    trials.  One broadcast of the generator state makes the assembly identical everywhere."""
    import numpy as _np
    rank, W = world()
    if W == 1:
        return
    box = [_np.random.get_state() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    if rank != 0:
        _np.random.set_state(box[0])


def gather_trials(local, trial_ids, num_trials):
    """Per-trial result objects of every rank -> list indexed by global trial id on rank 0
    (None elsewhere).  The one collective of a sharded driver run."""
    rank, W = world()
    if W == 1:
        parts = [list(zip(trial_ids, local))]
    else:
        parts = [None] * W if rank == 0 else None
        dist.gather_object(list(zip(trial_ids, local)), parts, dst=0)
        if rank != 0:
            return None
    out = [None] * int(num_trials)
    for part in parts:
        for t, res in part:
            out[t] = res
    return out


def shard_trials(num_trials, rank=None, world_size=None):
    """Trial r lives on rank r mod W; results are keyed by trial, so a sweep's
    output does not depend on the world size (trial-indexed Philox)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return list(range(rank, num_trials, world_size))


def gather_picks(trial_ids, task, action, gain, num_trials, device=None):
    """Gather (task[S], action[S], gain[S]) of every local trial to rank 0.
    Returns on rank 0 arrays indexed by GLOBAL trial id: task [N, S] int32,
    action [N, S] int32, gain [N, S] float64; None elsewhere.  One collective."""
    rank, W = world()
    task = np.asarray(task, dtype=np.int32)
    S = task.shape[1] if task.ndim == 2 else 1
    per = (num_trials + W - 1) // W
    dev = device or (torch.device('cuda', torch.cuda.current_device())
                     if dist.is_initialized() and dist.get_backend() == 'nccl' else torch.device('cpu'))
    buf = torch.zeros((per, 1 + 3 * S), dtype=torch.float64, device=dev)
    buf[:, 0] = -1
    n = len(trial_ids)
    if n:
        rows = np.concatenate([np.asarray(trial_ids, dtype=np.float64).reshape(n, 1),
                               task.reshape(n, S), np.asarray(action).reshape(n, S),
                               np.asarray(gain, dtype=np.float64).reshape(n, S)], axis=1)
        buf[:n] = torch.from_numpy(rows).to(dev)
    if W == 1:
        parts = [buf]
    else:
        parts = [torch.empty_like(buf) for _ in range(W)] if rank == 0 else None
        dist.gather(buf, parts, dst=0)
        if rank != 0:
            return None
    allrows = torch.cat(parts).cpu().numpy()
    allrows = allrows[allrows[:, 0] >= 0]
    out_t = np.full((num_trials, S), -1, dtype=np.int32)
    out_a = np.full((num_trials, S), -1, dtype=np.int32)
    out_g = np.full((num_trials, S), np.nan)
    ids = allrows[:, 0].astype(np.int64)
    out_t[ids] = allrows[:, 1:1 + S].astype(np.int32)
    out_a[ids] = allrows[:, 1 + S:1 + 2 * S].astype(np.int32)
    out_g[ids] = allrows[:, 1 + 2 * S:]
    return out_t, out_a, out_g


def run_sweep(num_trials, step_factory, make_inputs, streams=4, batch=1, groups=None):
    """Run ``num_trials`` independent trial-steps sharded over the ranks (the reference's
    trial loop, src/ocbo.py:104-107, with trial r on rank r mod W).
    ``step_factory(stream)`` -> SweepStep, or ``groups`` = SweepStep objects already built (one
    per stream, equal batch); ``make_inputs(trial)`` -> (X, y, Xs).
    Returns gather_picks(...) (rank 0) / None."""
    mine = shard_trials(num_trials)
    if groups is None:
        groups = [step_factory(torch.cuda.Stream()) for _ in range(streams)]
    streams, batch = len(groups), groups[0].B
    ids, tasks, acts, gains = [], [], [], []
    pending = [None] * streams
    chunks = [mine[i:i + batch] for i in range(0, len(mine), batch)]
    for c, chunk in enumerate(chunks):
        g = c % streams
        if pending[g] is not None:
            t, a, gn = groups[g].result_host()
            k = len(pending[g])
            ids += pending[g]; tasks.append(t[:k]); acts.append(a[:k]); gains.append(gn[:k])
        padded = chunk + [chunk[-1]] * (batch - len(chunk))
        groups[g].submit_host(padded, [make_inputs(t) for t in padded])
        pending[g] = chunk
    for g in range(streams):
        if pending[g] is not None:
            t, a, gn = groups[g].result_host()
            k = len(pending[g])
            ids += pending[g]; tasks.append(t[:k]); acts.append(a[:k]); gains.append(gn[:k])
    S = groups[0].S
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros((0, S), dtype=dt)
    return gather_picks(ids, cat(tasks, np.int32), cat(acts, np.int32), cat(gains, np.float64),
                        num_trials)
