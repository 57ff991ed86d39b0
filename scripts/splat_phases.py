#!/usr/bin/env python
"""Where the two splat kernels spend their time INSIDE a real step (graph replay, L2 evicted first): thread 0 of every
CTA stamps %globaltimer at its phase boundaries (dpc_debug_set(12, 1), dpc_debug_phase_read)."""
import ctypes, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dpc_b200 import _capi
os.environ["DPC_LAB"] = "1"      # lab build: experiment knobs / lab-only diagnostics
L = _capi.lib()
dev = torch.device("cuda:0")
pipe = bench.Pipeline(dev, 0)
flush = bench.L2Flush(dev, "write")
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
PH = [["entry", "staged", "transformed", "tr_pc stored", "past dependency", "reductions issued"],
      ["entry", "staged", "transformed", "past dependency", "gathers issued", "chain rule done", "d_pc stored"]]
for rep in range(2):
    for r in range(3):
        flush.fill_(r)
    L.dpc_debug_set(12, 1)
    g.replay()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (2 * 512 * 8))()
    _capi.check(L.dpc_debug_phase_read(ctypes.cast(buf, ctypes.c_void_p)))
    L.dpc_debug_set(12, 0)
    for which, name in ((0, "splat_fwd"), (1, "splat_bwd")):
        n = len(PH[which])
        ctas = [[buf[(which * 512 + c) * 8 + k] for k in range(n)] for c in range(256)]
        ctas = [c for c in ctas if c[0] != 0]
        t0 = min(c[0] for c in ctas)
        print("%s (#%d), %d CTAs, us since the first CTA's entry: median [min .. max]" % (name, rep, len(ctas)))
        for k in range(n):
            v = [(c[k] - t0) / 1e3 for c in ctas]
            d = [(c[k] - c[k - 1]) / 1e3 for c in ctas] if k else v
            print("   %-18s at %6.2f [%6.2f .. %6.2f]   phase length %6.2f [%6.2f .. %6.2f]" % (
                PH[which][k], statistics.median(v), min(v), max(v), statistics.median(d), min(d), max(d)))
