'''
Multi-GPU sharding of independent problem instances (multi-start guesses, vehicle-parameter and track sweeps).

The reference solves its problems one after the other in one process (scripts/fig_8.py:21-62,
scripts/race.py:33-49); nothing is distributed.  Instances are independent, so the B200 layout is one process
per GPU, a contiguous slice of the instance index per rank, every rank running the whole batched pipeline on
its slice, and NO collective on the data path -- only a final gather of per-instance results (status, lap
time, iterations; optionally the trajectories) on rank 0.  `torch.distributed` is plumbing: NCCL on the GPU
box, gloo in the CPU tests.
'''
import numpy as np


def shard_range(total, rank, world):
    ''' contiguous slice [lo, hi) of `total` instances owned by `rank`; sizes differ by at most one '''
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of {world}')
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(local: dict, total, dist=None, dst=0):
    '''
    gather per-instance result arrays (same keys, leading dimension = local slice) on rank `dst`.
    local: {'lap_time': (n_local,), 'status': (n_local,), ...} numpy arrays.  Returns the dict of full
    arrays on `dst` (instances in global order) and None elsewhere.  Without an initialised process group it
    returns `local` unchanged (world of one).
    '''
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return {k: np.asarray(v) for k, v in local.items()}
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(total, rank, world)
    for v in local.values():
        if len(v) != hi - lo:
            raise ValueError('local arrays must cover exactly this rank\'s slice')
    gathered = [None] * world if rank == dst else None
    dist.gather_object({k: np.asarray(v) for k, v in local.items()}, gathered, dst=dst)
    if rank != dst:
        return None
    out = {}
    for k in local:
        out[k] = np.concatenate([g[k] for g in gathered], axis=0)
        assert out[k].shape[0] == total
    return out
