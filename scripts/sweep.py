#!/usr/bin/env python
"""BASELINE config 5: N in {2k, 8k, 16k, 32k} x V in {32, 64, 128}, K = 21, B = 32 per GPU, forward + backward of the
fused path, timed exactly like bench.py's headline (the C-ABI step captured in a CUDA graph with the timing events as its
first / last nodes, L2 evicted between steps), at 1 / 2 / 4 / 8 GPUs (weak scaling, no data-path collective; launch with
torchrun for N > 1), with the reference's CPU path (oracle port, a bounded B = 4 sample) beside every cell (--cpu).

    python scripts/sweep.py --cpu                                                    # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/sweep.py --gpus 8
Prints one JSON object (rank 0); `frac` = algorithmic bytes of the full path / step time / HBM peak.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dpc_b200 import distributed as D  # noqa: E402


def run_cell(dev, n, v, steps, flush, rank):
    bench.N, bench.V = n, v
    bench.G_BYTES = v * v * v * 4
    pipe = bench.Pipeline(dev, seed_shift=rank)
    for _ in range(3):
        pipe.step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True, external=True)
    e1 = torch.cuda.Event(enable_timing=True, external=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        e0.record()
        pipe.step()
        e1.record()
    total = 0.0
    D.barrier()
    for it in range(steps + 2):
        flush.fill_(it & 0xff)
        g.replay()
        e1.synchronize()
        if it >= 2:
            total += e0.elapsed_time(e1)
    torch.cuda.synchronize()
    del g, pipe
    torch.cuda.empty_cache()
    return total / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--cpu", action="store_true", help="time the CPU port beside every cell (rank 0, B = 4 sample)")
    args = ap.parse_args()
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    bench._quiet_stdout()
    rank, local_rank, world = D.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    flush = bench.L2Flush(dev, "write+read")
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    threads = bench.pick_threads() if (args.cpu and rank == 0) else None
    cells = []
    for v in (32, 64, 128):
        for n in (2000, 8000, 16000, 32000):
            ms = run_cell(dev, n, v, args.steps, flush, rank)
            ms = D.reduce_scalar(ms, "max", dev)
            alg = 6 * 4 * v ** 3 + 4 * 12 * n + 64 * n + 2 * 4 * v * v
            cell = {"N": n, "V": v, "B_per_gpu": bench.B, "ms_per_step": ms, "proj_per_s": world * bench.B / (ms * 1e-3),
                    "alg_bytes_per_proj": alg, "alg_GBps_per_gpu": alg * bench.B / (ms * 1e-3) / 1e9,
                    "frac": alg * bench.B / (ms * 1e-3) / 1e9 / peak,
                    "smoothing_kernels": "tcgen05 pipelines" if v == 64 else "FFMA2 (CUDA cores)"}
            if threads is not None:
                total, k = bench.time_oracle(4, 2, 1, threads)
                cell["cpu_port_proj_per_s"] = 4 * k / total
                cell["cpu_cores"] = threads
            cells.append(cell)
            if rank == 0:
                print(cell, file=sys.stderr, flush=True)
    if rank == 0:
        bench.emit({"workload": "config 5 sweep: fused fwd+bwd, B=32 per GPU, K=21, sigma_rel=3, DRC, graph replay, L2 evicted between steps",
                    "n_gpus": world, "scaling": "weak", "hbm_peak_gbs": peak, "cells": cells})
    D.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
