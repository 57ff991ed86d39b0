#!/usr/bin/env python
"""Per-kernel timeline of ONE forward+backward step as bench.py runs it (C-ABI step, L2 evicted first):
for every kernel of the fused path, when its first CTA started, when CTAs got past the grid dependency,
and when the last CTA exited (%globaltimer, dpc_debug_set(12, 1)).  Shows the gaps between kernels."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dpc_b200 import _capi
L = _capi.lib()
dev = torch.device("cuda:0")
pipe = bench.Pipeline(dev, 0)
flush = bench.L2Flush(dev, "write")
NAMES = ["zero", "splat_fwd", "xy_fwd", "z_fwd", "zero4", "z_bwd", "xy_bwd", "splat_bwd"]


def one(label, graph=None):
    for rep in range(3):
        flush.fill_(rep)
    L.dpc_debug_set(12, 1)
    if graph is None:
        pipe.step()
    else:
        graph.replay()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 64)()
    _capi.check(L.dpc_debug_ktrace_read(ctypes.cast(buf, ctypes.c_void_p)))
    L.dpc_debug_set(12, 0)
    rows = [(NAMES[k], [buf[4 * k + j] for j in range(4)]) for k in range(8)]
    rows = [(n, v) for n, v in rows if v[3] != 0 and v[0] != 2 ** 64 - 1]
    t0 = min(v[0] for _, v in rows)
    print("%s" % label)
    prev_exit = None
    for n, v in rows:
        e, w0, w1, x = [(t - t0) / 1e3 for t in v]
        gap = "" if prev_exit is None else "  gap after prev exit %+6.2f" % (w0 - prev_exit)
        print("  %-10s entry %7.2f  dep passed %7.2f..%7.2f  exit %7.2f  busy %6.2f us%s" % (n, e, w0, w1, x, x - w0, gap))
        prev_exit = x
    print("  total %.2f us" % ((max(v[3] for _, v in rows) - t0) / 1e3))
    if buf[4 * 9 + 3] != 0:     # fused x/y + gather kernel: its roles
        rel = lambda t: (t - t0) / 1e3  # noqa: E731
        print("  fused xy_bwd roles: pipelines done (last CTA) %.2f | gather warps first start %.2f, last end %.2f | "
              "mean wait per gather warp %.2f us" % (rel(buf[4 * 8 + 3]), rel(buf[4 * 9 + 0]), rel(buf[4 * 9 + 3]),
                                                    buf[4 * 10 + 2] / 1e3 / (148 * (10 if '17=1' in os.environ.get('DPC_KNOBS', '') else 6))))


for _ in range(3):
    pipe.step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    pipe.step()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        pipe.step()
knobs = os.environ.get("DPC_KNOBS", "defaults")
for rep in range(3):
    one("graph replay, knobs %s (#%d)" % (knobs, rep), g)
one("eager, knobs %s" % knobs)
