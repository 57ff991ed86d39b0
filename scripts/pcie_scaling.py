#!/usr/bin/env python
"""Host <-> device copy bandwidth with 1, 2, 4, ... GPUs copying AT THE SAME TIME (one process per GPU, pinned buffers,
the e2e step's blob sizes): is bench.py's e2e 1 -> 8 curve bounded by the host fabric?  Prints one JSON line per GPU
count: per-rank and aggregate GB/s for H2D alone, D2H alone and both directions at once.
usage: python scripts/pcie_scaling.py [max_gpus]"""
import json
import os
import socket
import sys
import time

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = {"affinity": sorted(os.sched_getaffinity(0))[:2] + [len(os.sched_getaffinity(0))]}
    for mb in (3.6, 64.0):
        n = int(mb * 1e6 / 4)
        h_in = torch.empty(n, dtype=torch.float32).pin_memory()
        h_out = torch.empty(n, dtype=torch.float32).pin_memory()
        d_in = torch.empty(n, dtype=torch.float32, device=dev)
        d_out = torch.empty(n, dtype=torch.float32, device=dev)
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        reps = 200 if mb < 10 else 20

        def h2d():
            with torch.cuda.stream(s1):
                for _ in range(reps):
                    d_in.copy_(h_in, non_blocking=True)

        def d2h():
            with torch.cuda.stream(s2):
                for _ in range(reps):
                    h_out.copy_(d_out, non_blocking=True)

        def both():
            h2d(); d2h()

        for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
            fn(); torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            dist.barrier()
            out["%s_%gMB" % (name, mb)] = mb * 1e-3 * reps / dt          # GB/s per direction on this rank
    ret[rank] = out
    dist.destroy_process_group()


def main():
    max_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    n = 1
    while n <= min(max_gpus, torch.cuda.device_count()):
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(worker, args=(n, _free_port(), ret), nprocs=n, join=True)
        keys = [k for k in ret[0] if k != "affinity"]
        row = {"gpus": n, "affinity_rank0": ret[0]["affinity"]}
        for k in keys:
            vals = [ret[r][k] for r in range(n)]
            row[k] = {"per_rank_min": round(min(vals), 1), "per_rank_max": round(max(vals), 1), "aggregate": round(sum(vals), 1)}
        print(json.dumps(row), flush=True)
        n *= 2


if __name__ == "__main__":
    main()
