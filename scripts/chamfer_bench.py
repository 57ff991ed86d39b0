#!/usr/bin/env python
"""f-4 measurement: point_cloud_distance (both directions = one chamfer pair) at the evaluation's size, fp64 like
run/eval_chamfer.py and fp32, CUDA events; the numpy oracle on a bounded sample of the sources beside it."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dpc_b200.util.point_cloud_distance import point_cloud_distance
from oracle import dpc_oracle as O

NS, NT = 8000, 100000
g = torch.Generator().manual_seed(0)
out = {"workload": "chamfer pair: %d predicted vs %d ground-truth points, both directions" % (NS, NT)}
for dtype, name in ((torch.float64, "f64"), (torch.float32, "f32")):
    pred = (torch.tanh(0.5 * torch.randn(NS, 3, generator=g, dtype=torch.float64)) / 2).to(dtype).cuda()
    gt = (torch.tanh(0.5 * torch.randn(NT, 3, generator=g, dtype=torch.float64)) / 2).to(dtype).cuda()
    for _ in range(3):
        point_cloud_distance(pred, gt); point_cloud_distance(gt, pred)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        point_cloud_distance(pred, gt); point_cloud_distance(gt, pred)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pairs = 2.0 * NS * NT
    t0 = time.perf_counter()
    O.point_cloud_distance(pred[:400].cpu(), gt.cpu())
    cpu_s = (time.perf_counter() - t0) * (NS / 400.0) * 2.0      # both directions have the same pair count
    # roofline: 8 separately rounded operations per pair (3 sub, 3 mul, 2 add: the reference's op order forbids FMA
    # contraction), one FP64 (or FP32) instruction each.  Pipe peak = SMs x lanes/clk x SM clock, nominal lane counts
    # (64 FP64, 128 FP32 per SM and clock); the SM clock is read from NVML under load.
    try:
        import pynvml
        pynvml.nvmlInit()
        mhz = pynvml.nvmlDeviceGetClockInfo(pynvml.nvmlDeviceGetHandleByIndex(0), pynvml.NVML_CLOCK_SM)
    except Exception:
        mhz = 1965
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    lanes = 64 if name == "f64" else 128
    peak_ginstr = sms * lanes * mhz * 1e6 / 1e9
    achieved_ginstr = 8.0 * pairs / (ms * 1e-3) / 1e9
    out[name] = {"gpu_ms_per_pair_of_directions": ms, "gpu_gpairs_per_s": pairs / ms / 1e6,
                 "oracle_numpy_s_extrapolated_from_400_sources": cpu_s, "speedup": cpu_s * 1e3 / ms,
                 "roofline": {"bound": "%s pipe (8 non-fused ops per pair)" % ("FP64" if name == "f64" else "FP32"),
                              "achieved": achieved_ginstr, "peak": peak_ginstr, "unit": "G instr/s", "frac": achieved_ginstr / peak_ginstr,
                              "peak_source": "%d SMs x %d lanes/clk x %d MHz (nominal lanes, NVML clock)" % (sms, lanes, mhz),
                              "memory_traffic_note": "%.1f MB of points per direction pair against %.1f G pair evaluations: "
                                                     "not memory-bound" % ((NS + NT) * 3 * (8 if name == "f64" else 4) * 2 / 1e6, pairs / 1e9)}}
print(json.dumps(out))
