#!/usr/bin/env python
"""tcgen05.mma micro-benchmark: cycles for a chain of M x N x 8 tf32 SS MMAs + commit + wait (one issuing thread)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpc_b200 import _capi
os.environ["DPC_LAB"] = "1"      # lab build: experiment knobs / lab-only diagnostics
L = _capi.lib()
out = torch.zeros(148 * 3, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for M, N, ts in ((128, 64, 0), (128, 128, 0), (128, 64, 3), (128, 128, 3), (128, 256, 3), (64, 64, 3)):
    for nmma in (8, 24):
        out.zero_()
        _capi.check(L.dpc_debug_mma_bench(out.data_ptr(), 148, 128, 50, nmma, ts, M, N, st))
        torch.cuda.synchronize()
        o = out.view(-1, 3)[:148].float().mean(0).tolist()
        print("ts %d M %3d N %3d nmma %2d: %7.0f cyc/rep (issue %6.0f, wait %6.0f) -> %5.1f cyc/mma" % (ts, M, N, nmma, o[0], o[1], o[2], o[0] / nmma), flush=True)
