'''
Multi-process sharding of independent instances (N > 1 path), exercised with world_size 2 on CPU (gloo):
every instance is owned by exactly one rank, the per-rank seeds of bench.py's inputs follow the global
instance index, and the final gather returns results in global order.
'''
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aircraft_trajectory_optimization_b200.sharding import shard_range, gather_results

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything_once():
    for total in (0, 1, 7, 512, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(total, dtype=int)
            sizes = []
            for r in range(world):
                lo, hi = shard_range(total, r, world)
                seen[lo:hi] += 1
                sizes.append(hi - lo)
            assert (seen == 1).all() and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = shard_range(total, rank, world)
        # a stand-in for the per-instance pipeline: results depend on the GLOBAL instance index only
        idx = np.arange(lo, hi)
        local = dict(lap_time=5.0 + 1e-3 * idx, status=(idx % 3 == 0).astype(np.int64), index=idx)
        full = gather_results(local, total, dist, dst=0)
        t = torch.tensor([float(hi - lo)])
        dist.all_reduce(t)                      # the only collective a bench run uses: a reduction of scalars
        if rank == 0:
            q.put((full, float(t)))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_in_global_order():
    total, world = 37, 2
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    full, n = q.get()
    assert n == total
    assert np.array_equal(full['index'], np.arange(total))
    assert np.allclose(full['lap_time'], 5.0 + 1e-3 * np.arange(total))
    assert np.array_equal(full['status'], (np.arange(total) % 3 == 0).astype(np.int64))


def test_bench_inputs_follow_the_global_instance_index():
    ''' rank r of a 2-rank run sees exactly the instances [r*B, (r+1)*B) of the 1-rank workload '''
    sys.path.insert(0, ROOT)
    import bench
    from types import SimpleNamespace as NS
    st = NS(nw=12, ng=5, w0=np.linspace(0, 1, 12), lbw=-np.ones(12), ubw=2 * np.ones(12))
    vp0 = np.arange(1.0, 14.0)
    X, L, VP = bench.make_inputs(st, vp0, 8, seed0=0)
    X1, L1, VP1 = bench.make_inputs(st, vp0, 4, seed0=4)
    assert np.array_equal(X[4:], X1) and np.array_equal(L[4:], L1) and np.array_equal(VP[4:], VP1)
