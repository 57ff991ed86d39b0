"""CPU, world_size 2 over gloo: the N>1 path of the projection workload -- contiguous sharding of
the renderer batch, no data-path collective, results identical to the unsharded run."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dpc_b200.distributed import shard_range


def test_shard_range_partitions_exactly():
    for total in (1, 2, 7, 32, 33, 128):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from dpc_b200 import distributed as D
    from dpc_b200.util.config import default_config
    from oracle import dpc_oracle as O
    torch.set_num_threads(1)
    D.init(backend="gloo")
    cfg = default_config(vox_size=8, pc_gauss_kernel_size=3)
    g = torch.Generator().manual_seed(5)
    pc = torch.tanh(0.5 * torch.randn(5, 40, 3, generator=g)) / 2
    q = torch.randn(5, 4, generator=g)
    lo, hi = D.shard_range(5, rank, world)
    out = O.pointcloud_project_fast(cfg, pc[lo:hi], q[lo:hi], None, None, O.smoothing_kernel(cfg, 0.7))
    full = O.pointcloud_project_fast(cfg, pc, q, None, None, O.smoothing_kernel(cfg, 0.7))
    ok = torch.equal(out["proj"], full["proj"][lo:hi])
    # bookkeeping collectives of bench.py: max of times, sum of processed units
    tmax = D.reduce_scalar(float(rank + 1), "max")
    units = D.reduce_scalar(float(hi - lo), "sum")
    D.barrier()
    ret[rank] = (ok, tmax, units)
    dist.destroy_process_group()


def test_two_ranks_gloo_shard_equals_unsharded():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        ok, tmax, units = ret[r]
        assert ok
        assert tmax == 2.0
        assert units == 5.0
