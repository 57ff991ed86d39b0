#!/usr/bin/env python
"""Per-CTA globaltimer trace of the LAST persistent tensor-core kernel of a fused forward (= the depth pass) and of
a full forward+backward step (= the backward x/y pass), run exactly as bench.py runs them."""
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

def fwd_only():
    p = pipe.p
    _capi.check(L.dpc_project_fast_fwd(ctypes.byref(p), pipe.pc.data_ptr(), pipe.q.data_ptr(), None, None, pipe.sc1.data_ptr(),
                                       pipe.taps.data_ptr(), pipe.taps.data_ptr(), pipe.tr_pc.data_ptr(), pipe.vox.data_ptr(),
                                       pipe.proj.data_ptr(), None, None, pipe.scratch.data_ptr(), pipe.scratch_bytes,
                                       pipe.saved.data_ptr(), pipe.saved_bytes, pipe.stream))

def report(name):
    buf = (ctypes.c_longlong * 736)()
    _capi.check(L.dpc_debug_trace_read(ctypes.cast(buf, ctypes.c_void_p)))
    ct = [[buf[256 + 3 * c + j] for j in range(3)] for c in range(148)]
    t00 = min(c[0] for c in ct)
    print("%s: entry spread %d ns | setup done med %d max %d | exit min %d med %d max %d | loop med %d max %d" % (
        name, max(c[0] for c in ct) - t00, statistics.median(c[1] for c in ct) - t00, max(c[1] for c in ct) - t00,
        min(c[2] for c in ct) - t00, statistics.median(c[2] for c in ct) - t00, max(c[2] for c in ct) - t00,
        statistics.median(c[2] - c[1] for c in ct), max(c[2] - c[1] for c in ct)))

for _ in range(3):
    pipe.step()
torch.cuda.synchronize()
L.dpc_debug_set(9, 1)
for rep in range(3):
    flush.fill_(rep); flush.fill_(rep + 1)
    fwd_only(); torch.cuda.synchronize()
    report("fwd only  -> conv_z_fwd ")
for rep in range(3):
    flush.fill_(rep); flush.fill_(rep + 1)
    pipe.step(); torch.cuda.synchronize()
    report("full step -> conv_xy_bwd")
for rep in range(3):
    pipe.step(); torch.cuda.synchronize()
    report("full step, no flush     ")
