#!/usr/bin/env python
"""Experiment sweep on the GPU box: per-stage device times for each tuning-knob setting."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    pipe = bench.Pipeline(dev, 0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {}
    for clustered in (False,):
        if clustered:
            pc = bench.make_inputs(bench.B, 0, clustered=True)[0]
            pipe.pc.copy_(pc.to(dev))
        for zmin, zcp in ((1, 0), (0, 0), (0, 1)):
            pipe.L.dpc_debug_set(5, zmin)        # 1 = ignore the host taps (vector-register kernels)
            pipe.L.dpc_debug_set(6, zcp)         # 1 = cp.async tile load in the depth kernels
            for _ in range(3):
                pipe.step()
            st = pipe.stage_times(20, flush)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                pipe.step()
            e1.record()
            torch.cuda.synchronize()
            key = "%s_devtaps%d_zcp%d" % ("clustered" if clustered else "spread", zmin, zcp)
            out[key] = {k: round(v * 1000, 2) for k, v in st.items()}
            out[key]["step_us"] = round(e0.elapsed_time(e1) / 20 * 1000, 2)
            print(key, out[key], flush=True)
    pipe.L.dpc_debug_set(5, 0)
    pipe.L.dpc_debug_set(6, 0)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
