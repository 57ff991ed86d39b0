#!/usr/bin/env python
"""BASELINE config 5: N in {2k,8k,16k,32k} x V in {32,64,128}, K=21, B=32, forward+backward
through the C-ABI with resident inputs (CUDA events).  Writes gpurun_out/sweep.json."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpc_b200 import _capi  # noqa: E402
from dpc_b200.util import gauss_kernel as gk  # noqa: E402
from dpc_b200.util.config import default_config  # noqa: E402

B, K = 32, 21


def run(n, v, steps=20):
    dev = torch.device("cuda", 0)
    L = _capi.lib()
    cfg = default_config(vox_size=v, pc_gauss_kernel_size=K)
    g = torch.Generator().manual_seed(1234)
    pc = (torch.tanh(0.5 * torch.randn(B, n, 3, generator=g)) / 2).to(dev)
    q = torch.randn(B, 4, generator=g).to(dev)
    sc = torch.sigmoid(torch.randn(B, generator=g)).to(dev)
    gt = (torch.rand(B, v, v, generator=g) > 0.5).float().to(dev)
    taps = gk.smoothing_kernel(cfg, torch.tensor(3.0, device=dev)).taps_xy
    p = _capi.ProjectParams(B=B, N=n, Vz=v, V=v, pose_kind=0, mode=0, K=K, Kz=K, focal_const=1.875, cam_dist=2.0,
                            clip_eps=1e-5, max_depth=10.0, flags=0)
    sb, vb = L.dpc_project_fast_scratch_bytes(ctypes.byref(p)), L.dpc_project_fast_saved_bytes(ctypes.byref(p))
    scratch = torch.zeros(sb, dtype=torch.uint8, device=dev)
    saved = torch.empty(vb, dtype=torch.uint8, device=dev)
    f = lambda *s: torch.empty(*s, device=dev)  # noqa: E731
    tr, vox, proj, gp, dpc, dq, dsc = f(B, n, 3), f(B, v, v, v), f(B, v, v), f(B, v, v), f(B, n, 3), f(B, 4), f(B)
    st = torch.cuda.current_stream().cuda_stream

    def step():
        _capi.check(L.dpc_project_fast_fwd(ctypes.byref(p), pc.data_ptr(), q.data_ptr(), None, None, sc.data_ptr(),
                                           taps.data_ptr(), taps.data_ptr(), tr.data_ptr(), vox.data_ptr(), proj.data_ptr(),
                                           None, None, scratch.data_ptr(), sb, saved.data_ptr(), vb, st))
        torch.sub(proj, gt, out=gp)
        gp.mul_(1.0 / B)
        _capi.check(L.dpc_project_fast_bwd(ctypes.byref(p), pc.data_ptr(), q.data_ptr(), None, None, sc.data_ptr(),
                                           taps.data_ptr(), taps.data_ptr(), vox.data_ptr(), gp.data_ptr(), None, None, None,
                                           None, dpc.data_ptr(), dq.data_ptr(), None, None, dsc.data_ptr(),
                                           scratch.data_ptr(), sb, saved.data_ptr(), vb, st))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    gbytes = 4 * v ** 3
    alg = 6 * gbytes + 4 * 12 * n + 64 * n + 2 * 4 * v * v
    return {"N": n, "V": v, "ms_per_step": ms, "proj_per_s": B / (ms * 1e-3), "alg_bytes_per_proj": alg,
            "alg_GBps": alg * B / (ms * 1e-3) / 1e9}


def main():
    out = []
    for v in (32, 64, 128):
        for n in (2000, 8000, 16000, 32000):
            r = run(n, v)
            out.append(r)
            print(r, flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
