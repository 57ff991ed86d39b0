#!/usr/bin/env python
"""The splat backward's gathers alone (lab build, csrc/dpc_gather_bench.cuh): B x N threads each read the 2 x 2 rows of a
uniformly random cell of a [B,V,V,V] grid.  Times 20 back-to-back launches per configuration with CUDA events (grid
resident in L2, as behind the x/y pass of the backward) and, once per table, behind an L2 eviction."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DPC_LAB"] = "1"
from dpc_b200 import _capi  # noqa: E402

L = _capi.lib()
B, N, V = 32, 8000, 64
dev = torch.device("cuda", 0)
grid = torch.rand(B * V * V * V + 1024, device=dev)
out = torch.zeros(B * N, device=dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
NAMES = {0: "4 x 16 B + 4 x 4 B on straddling lanes (product pattern)", 1: "4 x 16 B", 2: "8 x 4 B", 3: "4 x 16 B, dependent chain",
         4: "4 x 16 B + 4 x 4 B all lanes", 5: "2 x 16 B", 6: "1 x 16 B", 7: "product pattern through cp.async -> smem"}


def run(variant, threads=128, ppt=1, share=0, pad=0, cg=0, reps=20, cold=False):
    def go():
        _capi.check(L.dpc_debug_gather_bench(grid.data_ptr(), out.data_ptr(), B, N, V, variant, threads, ppt, share, pad, cg, st))
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    if cold:
        tot = 0.0
        for r in range(5):
            flush.fill_(r)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); go(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / 5 * 1e3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print("B=%d N=%d V=%d; us per launch (back-to-back launches, grid in L2); empty-kernel floor = variant 6 with share 5" % (B, N, V))
print("\n-- variants, 128-thread CTAs, 1 point per thread")
for v in range(8):
    print("  var %d %-58s %6.2f us   (.cg %6.2f, cold L2 %6.2f)" % (v, NAMES[v], run(v), run(v, cg=1), run(v, cold=True)), flush=True)
print("\n-- product pattern (var 0): CTA size x points per thread")
for threads in (64, 128, 256, 512):
    print("  %4d threads: " % threads + "  ".join("ppt %d %6.2f" % (p, run(0, threads, p)) for p in (1, 2, 4, 8)), flush=True)
print("\n-- product pattern: resident CTAs per SM limited by dynamic shared memory (128-thread CTAs)")
for pad_kb, label in ((0, "unlimited"), (14, "16/SM"), (28, "8/SM"), (37, "6/SM"), (56, "4/SM"), (75, "3/SM"), (112, "2/SM"), (200, "1/SM")):
    try:        # variant 7 holds 5 KB of static shared memory itself
        v7 = "%6.2f" % run(7, pad=max(0, pad_kb - 6) * 1024)
    except Exception as exc:
        v7 = "n/a (%s)" % (str(exc)[:40],)
    print("  %-10s %6.2f us   var 1: %6.2f   var 3 (dependent): %6.2f   var 7 (cp.async): %s" % (label, run(0, pad=pad_kb * 1024), run(1, pad=pad_kb * 1024), run(3, pad=pad_kb * 1024), v7), flush=True)
print("\n-- lanes sharing a cell in groups of 2^s (distinct lines per load instruction / 2^s; sectors from L2 / 2^s)")
for s in range(6):
    print("  s=%d: var 0 %6.2f   var 1 %6.2f   var 2 %6.2f   var 6 %6.2f" % (s, run(0, share=s), run(1, share=s), run(2, share=s), run(6, share=s)), flush=True)
