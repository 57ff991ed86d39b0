#!/usr/bin/env python
"""Timeline of CTA 0 of the persistent depth-pass forward kernel (dpc_debug_set(9,1))."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpc_b200 import _capi
os.environ["DPC_LAB"] = "1"      # lab build: experiment knobs / lab-only diagnostics
L = _capi.lib()
dev = torch.device("cuda:0")
P = _capi.ptr
st = torch.cuda.current_stream().cuda_stream
B = 32
v = torch.rand(B, 64, 64, 64, device=dev)
x = torch.arange(21, dtype=torch.float64) - 10
t = torch.exp(-x * x / 18); t = (t / t.sum()).float().to(dev)
sc = torch.rand(B, device=dev) + 0.5
out = torch.empty_like(v); proj = torch.empty(B, 64, 64, device=dev); mask2 = torch.zeros(B * 64 * 64 * 2, dtype=torch.int32, device=dev)
L.dpc_debug_set(8, 2)
def run():
    _capi.check(L.dpc_conv_z_fwd(P(v), P(t), 21, 10, P(sc), 0, 1e-5, 2.0, 10.0, 0, B, 64, 64, P(out), P(mask2), P(proj), None, None, st))
for _ in range(3): run()
torch.cuda.synchronize()
L.dpc_debug_set(9, 1)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print("kernel %.1f us" % (e0.elapsed_time(e1) * 1e3))
buf = (ctypes.c_longlong * 736)()
_capi.check(L.dpc_debug_trace_read(ctypes.cast(buf, ctypes.c_void_p)))
tr = [[buf[e * 16 + i] for i in range(16)] for e in range(16)]
t0 = min(x for row in tr[:10] for x in row if x > 0)
names = ["P:top", "P:sfull", "P:Afree", "P:published", "I:opfull", "I:issued", "C:done", "C:epi end"]
print("tile " + " ".join("%12s" % n for n in names))
for i in range(8):
    print("%4d " % i + " ".join("%12d" % (tr[e][i] - t0 if tr[e][i] > 0 else -1) for e in range(8)))

import statistics
ct = [[buf[256 + 3 * c + j] for j in range(3)] for c in range(148)]
t00 = min(c[0] for c in ct)
print("CTA entry   : min %d max %d ns after first entry" % (min(c[0] for c in ct) - t00, max(c[0] for c in ct) - t00))
print("setup done  : min %d med %d max %d" % (min(c[1] for c in ct) - t00, statistics.median(c[1] for c in ct) - t00, max(c[1] for c in ct) - t00))
print("CTA exit    : min %d med %d max %d" % (min(c[2] for c in ct) - t00, statistics.median(c[2] for c in ct) - t00, max(c[2] for c in ct) - t00))
print("loop length : min %d med %d max %d ns" % (min(c[2] - c[1] for c in ct), statistics.median(c[2] - c[1] for c in ct), max(c[2] - c[1] for c in ct)))
print("CTA 0: entry %d setup %d exit %d" % tuple(x - t00 for x in ct[0]))

# back-to-back launches: steady-state cost per kernel including the launch gap
L.dpc_debug_set(9, 0)
def xy():
    _capi.check(L.dpc_conv_xy(P(v), P(out), P(t), 21, 10, P(t), 21, 10, B, 64, 64, 1, None, None, st))
for name, fn in (("conv_z_fwd", run), ("conv_xy", xy)):
    for lv in (2, 0):
        L.dpc_debug_set(8, lv)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        print("%s level %d: %.2f us per launch (20 back to back)" % (name, lv, e0.elapsed_time(e1) * 1e3 / 20))
# alternating z / xy (different kernels back to back)
for lv in (2, 0):
    L.dpc_debug_set(8, lv)
    for _ in range(3): run(); xy()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10): run(); xy()
    e1.record(); torch.cuda.synchronize()
    print("alternating z/xy level %d: %.2f us per pair" % (lv, e0.elapsed_time(e1) * 1e3 / 10))

# warm / produced-by-previous-kernel / cold inputs, timed per launch
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
v2 = v.clone()
def timed(fn, prep, n=10):
    tot = 0.0
    for it in range(n + 2):
        prep(it)
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b2.record(); torch.cuda.synchronize()
        if it >= 2: tot += a.elapsed_time(b2)
    return tot * 1e3 / n
for lv in (2, 0):
    L.dpc_debug_set(8, lv)
    for name, fn in (("conv_z_fwd", run), ("conv_xy", xy)):
        w = timed(fn, lambda it: None)
        p = timed(fn, lambda it: (flush.fill_(it & 255), v.copy_(v2)))
        c = timed(fn, lambda it: flush.fill_(it & 255))
        print("%s level %d: warm %.1f us, input just written by a kernel (after flush) %.1f us, cold %.1f us" % (name, lv, w, p, c))
