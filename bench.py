#!/usr/bin/env python
"""Benchmark of the projection hot path (BASELINE.json metric):

    point-cloud projections/sec (forward + backward) at B=32 N=8000 V=64, K=21 taps, sigma_rel=3.0,
    quaternion pose, occupancy scaling, DRC projection, loss = sum((gt-proj)^2)/2/B.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

One "step" = one forward+backward pass of `pointcloud_project_fast` over one batch of B=32
synthetic clouds per GPU (weak scaling: every rank renders its own 32 samples, no data-path
collective).  `value` is timed on the device with CUDA events, inputs resident in HBM, calling
the C-ABI directly; `e2e` is the same metric through the reference-shaped Python API with pinned
HOST buffers, H2D/D2H copies inside the timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B, N, V, K, SIGMA = 32, 8000, 64, 21, 3.0
METRIC = "point-cloud projections/sec (fwd+bwd) at B=32 N=8000 V=64"
UNIT = "projections/s"
G_BYTES = V * V * V * 4
# algorithmic bytes per projection, SURVEY.md 8(d): full path 6g + 4n + 64N + 2s
FULL_PATH_BYTES = 6 * G_BYTES + 4 * 12 * N + 64 * N + 2 * 4 * V * V
# per-launch algorithmic bytes of each stage, per projection (DESIGN.md "kernels and rooflines")
# dram__bytes_read.sum + dram__bytes_write.sum per launch at B=32, from the committed `ncu --set full` capture of this
# very command (profiles/r02_w_final_ncu_summary.md; cold caches: ncu flushes L2 before every kernel, so grids the
# previous kernel left in L2 are re-read from HBM, and the 32 MiB of output stays in L2 until a later kernel evicts it)
NCU_TRAFFIC_B32 = {
    "splat_fwd": 22.458880e6 + 0.077056e6,      # + 33.55 MB zero fill by cudaMemsetAsync (not a kernel of ours)
    "conv_xy_fwd": 33.595904e6 + 0.047104e6,
    "conv_z_fwd": 33.594112e6 + 0.878336e6,
    "conv_z_bwd": 35.167232e6 + 0.461568e6,
    "conv_xy_bwd": 34.641920e6 + 0.022016e6,
    "splat_bwd": 29.796864e6 + 0.000768e6,      # the software-pipelined form also reads the forward's tr_pc (3 MB)
}
NCU_TRAFFIC_SOURCE = "ncu --set full, dram read+write per launch, profiles/r02_w_final_ncu_summary.md"
STAGE_BYTES = {
    "splat_fwd": G_BYTES + 2 * 12 * N + 32 * N,          # zero grid + read pc + write tr_pc + 8 corner RMW
    "conv_xy_fwd": 2 * G_BYTES + G_BYTES // 32,          # read raw, write xy-smoothed, clip-mask bits
    "conv_z_fwd": 2 * G_BYTES + G_BYTES // 32 + 4 * V * V,  # read, write voxels, mask bits, silhouette
    "conv_z_bwd": 2 * G_BYTES + G_BYTES // 32 + 4 * V * V,
    "conv_xy_bwd": 2 * G_BYTES + G_BYTES // 32,
    "splat_bwd": G_BYTES + 2 * 12 * N + 32 * N,          # read d_raw (gathers) + read pc + write d_pc
}


def make_inputs(batch, seed_shift=0, clustered=False):
    g = lambda s: torch.Generator().manual_seed(s + 1000 * seed_shift)  # noqa: E731
    spread, seed = (0.025, 1235) if clustered else (0.5, 1234)
    pc = torch.tanh(spread * torch.randn(batch, N, 3, generator=g(seed))) / 2
    q = torch.randn(batch, 4, generator=g(1236))
    sc = torch.sigmoid(torch.randn(batch, 1, generator=g(1237)))
    gt = (torch.rand(batch, V, V, 1, generator=g(1238)) > 0.5).float()
    return pc, q, sc, gt


WORKLOAD = ("configs[1]: pointcloud_project_fast fwd+bwd, B=32 per GPU, N=8000, V=64, K=21, sigma_rel=3.0, "
            "DRC projection, quaternion pose, occupancy scaling")


def workload_config(args):
    """The `config` object of the JSON line: built from the command line only, so both arms print the same object
    (the driver compares them); what is specific to an arm goes to `config_detail`."""
    world = args.gpus
    strong = args.scaling == "strong"
    if args.l2_flush:
        l2 = ("GPU arm: 256 MiB written, then 256 MiB read, between steps (untimed): L2 evicted and left clean"
              if args.l2_flush_mode == "write+read" else "GPU arm: 256 MiB written between steps (untimed) to evict L2")
    else:
        l2 = "GPU arm: no flush; a step touches ~200 MB of grids > 126 MB L2"
    return {"workload": WORKLOAD.replace("B=32 per GPU", "B=32 in total, split over the ranks") if strong else WORKLOAD,
            "global_batch": 32 if strong else 32 * world,
            "parallelism": "independent samples sharded over ranks, no collective", "l2": l2}


def bench_cfg():
    from dpc_b200.util.config import default_config
    return default_config(vox_size=V, pc_gauss_kernel_size=K, pc_relative_sigma=SIGMA)


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def time_oracle(batch, steps, warmup, threads, keep=None):
    """The reference's CPU implementation of the path (oracle port: torch-CPU op-for-op restatement
    built like the reference -- 8 scatter grids + add_n, three conv3d, log-space DRC, autograd).
    keep: a dict that receives the last step's outputs and gradients (the checker of the `parity` block)."""
    from oracle import dpc_oracle as O
    torch.set_num_threads(threads)
    cfg = bench_cfg()
    pc, q, sc, gt = make_inputs(batch)
    ker = O.smoothing_kernel(cfg, torch.tensor(SIGMA))
    times = []
    for it in range(warmup + steps):
        a = [t.clone().requires_grad_(True) for t in (pc, q, sc)]
        t0 = time.perf_counter()
        out = O.pointcloud_project_fast(cfg, a[0], a[1], None, None, ker, a[2])
        loss = ((gt - out["proj"]) ** 2).sum() / 2 / batch
        loss.backward()
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0)
    if keep is not None:
        keep.update(proj=out["proj"].detach(), voxels=out["voxels"].detach(), tr_pc=out["tr_pc"].detach(),
                    d_pc=a[0].grad, d_q=a[1].grad, d_scale=a[2].grad, loss=float(loss))
    return sum(times), len(times)


def pick_threads():
    """The reference's CPU path 'with all the host threads it can use': on a 128-core host the
    single-channel conv3d of the restatement gets SLOWER beyond a few dozen threads (measured:
    1.25 proj/s at 128 threads), so calibrate on a tiny sample and keep the fastest count."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_t = cands[0], None
    for c in cands:
        total, n = time_oracle(2, 1, 1, c)
        if best_t is None or total < best_t:
            best, best_t = c, total
    return best


def run_reference(args, rank):
    if rank != 0:
        return
    threads = pick_threads()
    batch = args.ref_batch
    total, n = time_oracle(batch, args.steps, args.warmup, threads)
    value = batch * n / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": n, "warmup": args.warmup, "ms_per_step": 1000.0 * total / n,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "config_detail": {"sample_batch": batch, "note": "reference TF1 path restated on torch-CPU (TensorFlow unavailable); "
                          "every step is one B=%d batch of the workload on the host cores (rank 0 only)" % batch},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d steps of a B=%d batch of the same workload" % (n, batch)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for nm in dir(nv):
            if nm.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, nm), int):
                names[getattr(nv, nm)] = nm.replace("nvmlClocksThrottleReason", "")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if bit and (bits & bit) == bit and nm not in ("None", "All"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


class L2Flush:
    """Evicts L2 between timed steps (untimed): writes a 256 MiB buffer (> 126 MB L2).  Mode "write+read" then also
    READS a second 256 MiB buffer, so that the lines left in L2 are clean: otherwise the first ~126 MB of lines the
    timed step allocates each have to write a dirty line of the FLUSH buffer back to HBM first, and the step is
    charged for the flush's own write-back traffic."""

    def __init__(self, dev, mode):
        self.mode = mode
        self.w = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        self.r = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev) if mode == "write+read" else None
        self.sink = torch.zeros(1, dtype=torch.float32, device=dev)

    def fill_(self, v):
        self.w.fill_(v)
        if self.r is not None:
            torch.sum(self.r, dim=0, keepdim=True, out=self.sink)

    def describe(self):
        if self.mode == "write+read":
            return ("256 MiB written, then 256 MiB read, between steps (untimed): L2 evicted and left clean, so the "
                    "timed step is not charged for writing the flush buffer's dirty lines back")
        return "256 MiB written between steps (untimed) to evict L2"


# ----------------------------------------------------------------------------- our arm
class Pipeline:
    """Device buffers + C-ABI calls for one rank (B samples)."""

    def __init__(self, dev, seed_shift, clustered=False, sigma=None, max_projection=False):
        from dpc_b200 import _capi
        from dpc_b200.util import gauss_kernel as gk
        self.capi, self.L, self.dev = _capi, _capi.lib(), dev
        self.cfg = bench_cfg()
        sigma = SIGMA if sigma is None else sigma
        host = make_inputs(B, seed_shift, clustered)
        self.host = [t.pin_memory() for t in host]
        self.pc, self.q, self.sc, self.gt = [t.to(dev) for t in host]
        self.sc1 = self.sc.reshape(-1).contiguous()
        self.gt3 = self.gt.reshape(B, V, V).contiguous()
        self.neg_gt_over_b = (-self.gt3 / B).contiguous()
        # sigma is a host value (the reference's schedule is a function of the step count, model_pc.py:35-40):
        # the kernel object holds the taps on the host and on the device
        self.kernel = gk.smoothing_kernel(self.cfg, sigma)
        self.taps = self.kernel.device_taps(dev)[0]
        self.p = _capi.ProjectParams(B=B, N=N, Vz=V, V=V, pose_kind=_capi.POSE_QUAT,
                                     mode=_capi.PROJ_MAX if max_projection else _capi.PROJ_DRC, K=K, Kz=K,
                                     focal_const=float(self.cfg.focal_length), cam_dist=float(self.cfg.camera_distance),
                                     clip_eps=float(self.cfg.drc_logsum_clip_val), max_depth=float(self.cfg.max_depth))
        self.p.flags = 0
        self.p.taps_xy_host = self.kernel.host_taps_xy.data_ptr()
        self.p.taps_z_host = self.kernel.host_taps_z.data_ptr()
        self.scratch_bytes = self.L.dpc_project_fast_scratch_bytes(ctypes.byref(self.p))
        self.saved_bytes = self.L.dpc_project_fast_saved_bytes(ctypes.byref(self.p))
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)  # noqa: E731
        self.scratch = torch.zeros(self.scratch_bytes, dtype=torch.uint8, device=dev)
        self.saved = torch.empty(self.saved_bytes, dtype=torch.uint8, device=dev)
        self.tr_pc, self.vox, self.proj, self.g_proj = f(B, N, 3), f(B, V, V, V), f(B, V, V), f(B, V, V)
        self.d_pc, self.d_q, self.d_sc = f(B, N, 3), f(B, 4), f(B)
        self.p.tr_pc = self.tr_pc.data_ptr()       # backward: the forward's own cells (read only by dpc_project_fast_bwd)
        self.loss = f(1)
        self.loss_ws_bytes = int(self.L.dpc_proj_l2_loss_workspace_bytes())
        self.loss_ws = torch.zeros((self.loss_ws_bytes + 3) // 4, dtype=torch.int32, device=dev)

    @property
    def stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream      # the capture stream while a graph is recorded

    def step(self):
        L, p, c = self.L, self.p, self.capi.check
        c(L.dpc_project_fast_fwd(ctypes.byref(p), self.pc.data_ptr(), self.q.data_ptr(), None, None, self.sc1.data_ptr(),
                                 self.taps.data_ptr(), self.taps.data_ptr(), self.tr_pc.data_ptr(), self.vox.data_ptr(),
                                 self.proj.data_ptr(), None, None, self.scratch.data_ptr(), self.scratch_bytes,
                                 self.saved.data_ptr(), self.saved_bytes, self.stream))
        # dL/dproj = (proj - gt)/B of loss = sum((gt-proj)^2)/2/B  (model_pc.py:414-415): one launch of the library's loss
        # kernel in its gradient-only form (the device-only step never read the loss VALUE; the e2e step, which returns
        # it to the host, runs the full form through util/losses.proj_l2_loss)
        c(L.dpc_proj_l2_loss(self.proj.data_ptr(), self.gt3.data_ptr(), B * V * V, 1.0 / B, None,
                             self.g_proj.data_ptr(), None, 0, self.stream))
        c(L.dpc_project_fast_bwd(ctypes.byref(p), self.pc.data_ptr(), self.q.data_ptr(), None, None, self.sc1.data_ptr(),
                                 self.taps.data_ptr(), self.taps.data_ptr(), self.vox.data_ptr(), self.g_proj.data_ptr(),
                                 None, None, None, None, self.d_pc.data_ptr(), self.d_q.data_ptr(), None, None,
                                 self.d_sc.data_ptr(), self.scratch.data_ptr(), self.scratch_bytes,
                                 self.saved.data_ptr(), self.saved_bytes, self.stream))

    LAUNCHES_PER_STEP = 7  # splat_fwd, conv_xy, conv_z_fwd | proj_l2_loss | conv_z_bwd, conv_xy, splat_bwd (+ the driver's memset node)

    KT_NAMES = ("zero", "splat_fwd", "conv_xy_fwd", "conv_z_fwd", "zero4", "conv_z_bwd", "conv_xy_bwd", "splat_bwd")

    def kernel_timeline(self, graph, flush, reps=5):
        """Per-kernel busy time inside the step exactly as it is timed (graph replay, PDL chaining intact):
        %globaltimer stamps taken by the kernels themselves (dpc_debug_set(12, 1)); busy = first CTA past its grid
        dependency -> last CTA exit.  The stage events above serialise the kernels, this does not."""
        L = self.L
        acc, total = {}, 0.0
        buf = (ctypes.c_ulonglong * 64)()
        for rep in range(reps):
            if flush is not None:
                flush.fill_(rep & 0xff)
            L.dpc_debug_set(12, 1)
            if graph is not None:
                graph.replay()
            else:
                self.step()
            torch.cuda.synchronize()
            self.capi.check(L.dpc_debug_ktrace_read(ctypes.cast(buf, ctypes.c_void_p)))
            L.dpc_debug_set(12, 0)
            rows = {self.KT_NAMES[k]: [buf[4 * k + j] for j in range(4)] for k in range(8)}
            rows = {n: v for n, v in rows.items() if v[3] != 0 and v[0] != 2 ** 64 - 1}
            if not rows:
                return None
            for n, v in rows.items():
                acc[n] = acc.get(n, 0.0) + (v[3] - v[1]) / 1e3
            total += (max(v[3] for v in rows.values()) - min(v[0] for v in rows.values())) / 1e3
        out = {n: round(t / reps, 2) for n, t in acc.items()}
        out["first_entry_to_last_exit"] = round(total / reps, 2)
        return out

    STAGES = ("splat_fwd", "conv_xy_fwd", "conv_z_fwd", "conv_z_bwd", "conv_xy_bwd", "splat_bwd")

    def stage_times(self, steps, flush):
        """Per-stage device time of the SAME fused entry points step() calls: the library records
        CUDA events on the launch stream around every stage (dpc_debug_set(3, 1))."""
        L = self.L
        tot = [0.0] * 6
        buf = (ctypes.c_float * 6)()
        L.dpc_debug_set(3, 1)
        try:
            for it in range(steps + 2):
                if flush is not None:
                    # three 256 MiB fills (~0.25 ms of GPU work): evicts L2 AND lets the host finish enqueueing the
                    # whole step before the GPU starts it, so the stage events time kernels, not host launch gaps
                    for rep in range(3):
                        flush.fill_((it + rep) & 0xff)
                self.step()
                self.capi.check(L.dpc_debug_stage_ms(ctypes.cast(buf, ctypes.c_void_p)))
                if it >= 2:
                    for i in range(6):
                        tot[i] += buf[i]
        finally:
            L.dpc_debug_set(3, 0)
        return {k: tot[i] / steps for i, k in enumerate(self.STAGES)}

    def _e2e_setup(self):
        """One pinned host blob per direction: inputs (pc|q|scale|gt) and results (loss|proj|d_pc|d_q|d_scale)."""
        hp, hq, hs, hg = self.host
        self.in_sizes = [t.numel() for t in (hp, hq, hs, hg)]
        self.in_shapes = [t.shape for t in (hp, hq, hs, hg)]
        self.h_in = torch.cat([t.reshape(-1) for t in (hp, hq, hs, hg)]).pin_memory()
        self.d_in = torch.empty_like(self.h_in, device=self.dev)
        self.out_sizes = [1, B * V * V, B * N * 3, B * 4, B]
        self.d_out = torch.empty(sum(self.out_sizes), dtype=torch.float32, device=self.dev)
        self.h_out = torch.empty(sum(self.out_sizes), dtype=torch.float32).pin_memory()

    def e2e_step(self):
        """Through the public API with HOST buffers: one H2D of the step's inputs from pinned memory,
        forward, loss, backward (autograd), one D2H of loss + silhouettes + gradients."""
        from dpc_b200.util import point_cloud as pcm
        if not hasattr(self, "h_in"):
            self._e2e_setup()
        self.d_in.copy_(self.h_in, non_blocking=True)
        parts = torch.split(self.d_in, self.in_sizes)
        pc, q, sc, gt = [p.reshape(sh) for p, sh in zip(parts, self.in_shapes)]
        pc, q, sc = pc.requires_grad_(True), q.requires_grad_(True), sc.requires_grad_(True)
        out = pcm.pointcloud_project_fast(self.cfg, pc, q, None, None, self.kernel, sc)
        loss = ((gt - out["proj"]) ** 2).sum() / 2 / B
        gpc, gq, gsc = torch.autograd.grad(loss, (pc, q, sc))
        dst = torch.split(self.d_out, self.out_sizes)
        for d, t in zip(dst, (loss.detach(), out["proj"].detach(), gpc, gq, gsc)):
            d.copy_(t.reshape(-1))
        self.h_out.copy_(self.d_out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.h_out[0]), self.h_in.numel() * 4, self.h_out.numel() * 4


def e2e_pipelined(pipe, steps):
    """End-to-end throughput with HOST buffers: every step copies its inputs from pinned host memory
    to the device and its results (loss, silhouettes, gradients) back, but the copies of step i+1 /
    i-1 run on their own streams underneath the compute of step i (double buffered)."""
    from dpc_b200.util import point_cloud as pcm
    dev = pipe.dev
    if not hasattr(pipe, "h_in"):
        pipe._e2e_setup()
    s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    d_in = [torch.empty_like(pipe.d_in) for _ in range(2)]
    d_out = [torch.empty_like(pipe.d_out) for _ in range(2)]
    h_out = [torch.empty_like(pipe.h_out).pin_memory() for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]     # compute no longer reads d_in[k]

    def h2d(i):
        k = i & 1
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[k])
            d_in[k].copy_(pipe.h_in, non_blocking=True)
            ev_in[k].record(s_in)

    for k in range(2):
        ev_free[k].record(s_cmp)
        ev_out[k].record(s_out)
    torch.cuda.synchronize()
    loss = 0.0
    t0 = time.perf_counter()
    h2d(0)
    for i in range(steps):
        k = i & 1
        if i + 1 < steps:
            h2d(i + 1)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[k])
            s_cmp.wait_event(ev_out[k])            # d_out[k] has been drained by the copy of step i-2
            parts = torch.split(d_in[k], pipe.in_sizes)
            pc, q, sc, gt = [p.reshape(sh) for p, sh in zip(parts, pipe.in_shapes)]
            pc, q, sc = pc.requires_grad_(True), q.requires_grad_(True), sc.requires_grad_(True)
            out = pcm.pointcloud_project_fast(pipe.cfg, pc, q, None, None, pipe.kernel, sc)
            l = ((gt - out["proj"]) ** 2).sum() / 2 / B
            gpc, gq, gsc = torch.autograd.grad(l, (pc, q, sc))
            for d, t in zip(torch.split(d_out[k], pipe.out_sizes), (l.detach(), out["proj"].detach(), gpc, gq, gsc)):
                d.copy_(t.reshape(-1))
            ev_free[k].record(s_cmp)
            ev_cmp[k].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[k])
            h_out[k].copy_(d_out[k], non_blocking=True)
            ev_out[k].record(s_out)
        if i >= 1:
            ev_out[(i - 1) & 1].synchronize()       # the host reads step i-1's results
            loss = float(h_out[(i - 1) & 1][0])
    ev_out[(steps - 1) & 1].synchronize()
    loss = float(h_out[(steps - 1) & 1][0])
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    return dt, loss


def e2e_graphed(pipe, steps, nbuf=4, host_inputs=True):
    """End to end with HOST buffers, the way a throughput-minded caller would run it: ONE CUDA graph per buffer set holds
    the whole step -- H2D of the step's inputs from pinned memory, the step a user builds with the public API
    (pointcloud_project_fast -> loss -> autograd), D2H of loss + silhouettes + gradients into pinned memory -- and
    `nbuf` such graphs are in flight, each on its own stream.  Per step the host does one graph launch and (nbuf - 1
    steps later) one event wait; the copies of one step run under the kernels of its neighbours.  The eager path is
    bounded by ~0.5 ms of Python/autograd time per step, separate copy/compute streams with per-step event bookkeeping
    by ~0.15 ms of host time (profiles/r01_k_pcie_probe.txt: H2D 3.6 MB 0.11 ms, D2H 0.07 ms, bookkeeping 0.03 ms)."""
    from dpc_b200.util import point_cloud as pcm
    from dpc_b200.util.losses import proj_l2_loss
    dev = pipe.dev
    if not hasattr(pipe, "h_in"):
        pipe._e2e_setup()
    torch.cuda.synchronize()
    pcm.release_scratch()            # scratch grids are cached per stream: drop those of streams that are gone
    streams = [torch.cuda.Stream(dev) for _ in range(nbuf)]
    d_in = [torch.empty_like(pipe.d_in) for _ in range(nbuf)]
    h_out = [torch.empty_like(pipe.h_out).pin_memory() for _ in range(nbuf)]
    h_parts = [torch.split(h, pipe.out_sizes) for h in h_out]

    if not host_inputs:
        # what a training step does: the decoder's points never leave HBM and only the loss value is read back
        for k in range(nbuf):
            d_in[k].copy_(pipe.h_in)

    def step_fn(k):
        if host_inputs:
            d_in[k].copy_(pipe.h_in, non_blocking=True)                               # H2D, 3.6 MB
        parts = torch.split(d_in[k], pipe.in_sizes)
        pc, q, sc, gt = [p.reshape(sh) for p, sh in zip(parts, pipe.in_shapes)]
        pc, q, sc = pc.detach().requires_grad_(True), q.detach().requires_grad_(True), sc.detach().requires_grad_(True)
        out = pcm.pointcloud_project_fast(pipe.cfg, pc, q, None, None, pipe.kernel, sc)
        l = proj_l2_loss(gt, out["proj"], B)              # sum((gt - proj)^2) / 2 / B, model_pc.py:414-415
        gpc, gq, gsc = torch.autograd.grad(l, (pc, q, sc))
        if host_inputs:
            for h, t in zip(h_parts[k], (l.detach(), out["proj"].detach(), gpc, gq, gsc)):    # D2H, 3.6 MB
                h.copy_(t.reshape(-1), non_blocking=True)
        else:
            h_parts[k][0].copy_(l.detach().reshape(-1), non_blocking=True)                   # D2H, 4 bytes

    graphs = []
    torch.cuda.synchronize()
    for k in range(nbuf):
        with torch.cuda.stream(streams[k]):
            for _ in range(3):
                step_fn(k)
    torch.cuda.synchronize()
    for k in range(nbuf):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=streams[k]):
            step_fn(k)
        graphs.append(g)
    torch.cuda.synchronize()
    done = [torch.cuda.Event() for _ in range(nbuf)]

    lag = nbuf - 1
    loss = 0.0
    t0 = time.perf_counter()
    for i in range(steps):
        k = i % nbuf
        if i >= nbuf:
            done[k].synchronize()                 # the host consumes step i - nbuf's results before its buffers are reused
            loss = float(h_out[k][0])
        with torch.cuda.stream(streams[k]):
            graphs[k].replay()
            done[k].record(streams[k])
    for i in range(max(0, steps - nbuf), steps):
        done[i % nbuf].synchronize()
    loss = float(h_out[(steps - 1) % nbuf][0])
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    return dt, loss


def fused_gather_route(pipe):
    """True when the experimental backward route (gathers inside the x/y pass, knob 4) is in effect for this pipeline."""
    return bool(pipe.p.tr_pc) and "4=1" in os.environ.get("DPC_KNOBS", "").split(",")


def parity_block(pipe, graph, ref):
    """Max abs differences between the timed step's results (re-run once so the buffers hold that step) and the oracle on
    the same seeded inputs -- the north star's gate (tr_pc / indices bit-exact, 1e-5 abs elsewhere), at the headline shape."""
    if graph is not None:
        graph.replay()
    else:
        pipe.step()
    torch.cuda.synchronize()

    def mad(dev_t, ref_t):
        return float((dev_t.detach().cpu().double().reshape(-1) - ref_t.double().reshape(-1)).abs().max())

    proj = pipe.proj.cpu()
    loss = float(((pipe.gt3.cpu() - proj) ** 2).sum() / 2 / B)
    return {
        "checker": "oracle (torch-CPU restatement of the reference), same seeded inputs as the timed step",
        "tr_pc_bit_exact": bool(torch.equal(pipe.tr_pc.cpu(), ref["tr_pc"])),
        "max_abs_proj": mad(pipe.proj, ref["proj"]),
        "mean_abs_proj": float((proj.double().reshape(-1) - ref["proj"].double().reshape(-1)).abs().mean()),
        "max_abs_voxels": mad(pipe.vox, ref["voxels"]),
        "max_abs_d_pc": mad(pipe.d_pc, ref["d_pc"]), "max_abs_d_q": mad(pipe.d_q, ref["d_q"]),
        "max_abs_d_scale": mad(pipe.d_sc, ref["d_scale"]),
        "grad_magnitude": {"d_pc": float(ref["d_pc"].abs().max()), "d_q": float(ref["d_q"].abs().max()),
                           "d_scale": float(ref["d_scale"].abs().max())},
        "loss": loss, "loss_oracle": ref["loss"], "tolerance_abs": 1e-5,
    }


def variant_rows(dev, steps, flush):
    """SURVEY 8(d)'s other rows of the same workload, timed like the headline (graph replay, in-graph events, L2
    evicted between steps): init-clustered cloud (decoder init, stddev 0.025: every point of a sample in ~27 voxels,
    worst-case contention), the end of the sigma schedule (0.2), max projection instead of DRC; plus the fraction of
    points that land inside the cube."""
    rows = {}
    for name, kw in (("spread_sigma3_drc", {}), ("clustered_init", {"clustered": True}), ("sigma_0.2", {"sigma": 0.2}),
                     ("max_projection", {"max_projection": True})):
        pipe = Pipeline(dev, 0, **kw)
        for _ in range(3):
            pipe.step()
        torch.cuda.synchronize()
        tr = pipe.tr_pc
        valid = float(((tr >= -0.5) & (tr <= 0.5)).all(dim=-1).float().mean())
        e0 = torch.cuda.Event(enable_timing=True, external=True)
        e1 = torch.cuda.Event(enable_timing=True, external=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            e0.record()
            pipe.step()
            e1.record()
        total = 0.0
        for it in range(steps + 2):
            if flush is not None:
                flush.fill_(it & 0xff)
            g.replay()
            e1.synchronize()
            if it >= 2:
                total += e0.elapsed_time(e1)
        rows[name] = {"ms_per_step": total / steps, "projections_per_s": B * steps / (total / 1000.0), "valid_point_fraction": valid}
        del g, pipe
    return rows


def back_to_back_row(dev, rank, steps, nsets=4):
    """The same C-ABI step run back to back WITHOUT the per-step event pair and the eviction fill: `nsets` independent
    buffer sets (inputs, scratch, saved, outputs: ~105 MB each, together well beyond the 126 MB L2) take turns, so a set's
    lines have been pushed out by the other sets' traffic when its turn comes again; one event pair around all steps.
    The difference to `value` is what the two event-record nodes of every timed step cost (~5 us)."""
    pipes = [Pipeline(dev, seed_shift=rank + 101 * k) for k in range(nsets)]
    graphs = []
    for pipe in pipes:
        for _ in range(3):
            pipe.step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            pipe.step()
        graphs.append(g)
    for k in range(2 * nsets):
        graphs[k % nsets].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        graphs[k % nsets].replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    set_bytes = sum(t.numel() * t.element_size() for t in (pipes[0].pc, pipes[0].scratch, pipes[0].saved, pipes[0].tr_pc,
                                                             pipes[0].vox, pipes[0].proj, pipes[0].g_proj, pipes[0].d_pc))
    del graphs, pipes
    torch.cuda.empty_cache()
    return ms, set_bytes


def run_ours(args, rank, local_rank, world):
    from dpc_b200 import distributed as D
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback for the product path); "
                           "use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pipe = Pipeline(dev, seed_shift=rank)
    flush = L2Flush(dev, args.l2_flush_mode) if args.l2_flush else None
    for _ in range(max(3, args.warmup)):
        pipe.step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    D.barrier()
    torch.cuda.synchronize()
    # The step (6 kernels, 5 memsets, 2 loss ops) is launch-bound from Python: ~90 us of host time to enqueue
    # ~80 us of GPU work.  It is therefore captured once in a CUDA graph (same C-ABI calls, same buffers) and the
    # timed region replays it; --no-graph times the eager calls instead.
    graph = None
    g_ev0 = g_ev1 = None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                pipe.step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            # the timing events are NODES of the graph (external events), recorded right in front of the step's first
            # node and right behind its last one: the step's device time, without the graph's launch latency behind the
            # untimed L2-eviction fills (that latency, ~10 us, overlaps the previous step when steps run back to back)
            try:
                g_ev0 = torch.cuda.Event(enable_timing=True, external=True)
                g_ev1 = torch.cuda.Event(enable_timing=True, external=True)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    g_ev0.record()
                    pipe.step()
                    g_ev1.record()
                graph.replay()
                torch.cuda.synchronize()
                if not (0.0 < g_ev0.elapsed_time(g_ev1) < 50.0):
                    raise RuntimeError("in-graph events returned %r ms" % g_ev0.elapsed_time(g_ev1))
            except Exception as exc:
                print("in-graph timing events unavailable (%r): events around the graph launch" % (exc,), file=sys.stderr)
                g_ev0 = g_ev1 = None
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    pipe.step()
                graph.replay()
                torch.cuda.synchronize()
        except Exception as exc:
            print("step graph capture failed, timing eager calls: %r" % (exc,), file=sys.stderr)
            graph = None
    sampler.start()
    evs = []
    ms_in_graph = 0.0
    for it in range(args.steps):
        if flush is not None:
            flush.fill_(it & 0xff)          # evict L2 between steps (outside the timed events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if graph is not None:
            graph.replay()
        else:
            pipe.step()
        e1.record()
        evs.append((e0, e1))
        if g_ev0 is not None:
            g_ev1.synchronize()             # the in-graph events are re-recorded by every replay: read them per step
            ms_in_graph += g_ev0.elapsed_time(g_ev1)
    torch.cuda.synchronize()
    D.barrier()
    clocks = sampler.stop()
    ms_outside = sum(a.elapsed_time(b) for a, b in evs)
    ms_outside = D.reduce_scalar(ms_outside, "max", dev)
    ms_total = ms_in_graph if g_ev0 is not None else sum(a.elapsed_time(b) for a, b in evs)
    ms_total = D.reduce_scalar(ms_total, "max", dev)
    units = D.reduce_scalar(float(B * args.steps), "sum", dev)
    value = units / (ms_total / 1000.0)

    # end-to-end through the Python API with host buffers
    for _ in range(3):
        pipe.e2e_step()
    torch.cuda.synchronize()
    D.barrier()
    t0 = time.perf_counter()
    n_e2e = max(5, min(args.steps, 50))
    for _ in range(n_e2e):
        loss, h2d, d2h = pipe.e2e_step()
    t_e2e = D.reduce_scalar(time.perf_counter() - t0, "max", dev)
    e2e_serial = world * B * n_e2e / t_e2e
    serial_ms = 1000.0 * t_e2e / n_e2e
    # the same per-step copies, overlapped with the neighbouring steps' compute on separate streams
    e2e_pipelined(pipe, 5)
    D.barrier()
    t_pipe, loss = e2e_pipelined(pipe, n_e2e)
    t_pipe = D.reduce_scalar(t_pipe, "max", dev)
    e2e_value = world * B * n_e2e / t_pipe
    t_e2e_serial, t_e2e = t_e2e, t_pipe
    e2e_mode = "eager API calls; copies of neighbouring steps overlap compute (3 streams, double buffered)"
    e2e_eager = e2e_value
    try:
        n_g = max(n_e2e, 100)
        e2e_graphed(pipe, 10)
        D.barrier()
        t_g, loss_g = e2e_graphed(pipe, n_g)
        t_g = D.reduce_scalar(t_g, "max", dev)
        if abs(loss_g - loss) <= 1e-3 * max(1.0, abs(loss)):
            e2e_value, t_e2e, n_e2e_used = world * B * n_g / t_g, t_g, n_g
            e2e_mode = ("one CUDA graph per buffer set = H2D of the inputs + the API-built step (pointcloud_project_fast -> loss -> "
                        "autograd) + D2H of loss/silhouettes/gradients; 4 graphs in flight on 4 streams, the host reads step "
                        "i-4's results before relaunching its graph")
            n_e2e = n_g
    except Exception as exc:  # capture not possible: keep the eager number
        print("graph capture failed: %r" % (exc,), file=sys.stderr)
    # second, clearly labelled row: the same API-built step with DEVICE-resident inputs and only the loss read back (what the
    # train step does: the decoder's points never leave HBM) -- separates the kernels + API from the host fabric
    e2e_dev = None
    try:
        e2e_graphed(pipe, 10, host_inputs=False)
        D.barrier()
        t_d, loss_d = e2e_graphed(pipe, max(n_e2e, 100), host_inputs=False)
        t_d = D.reduce_scalar(t_d, "max", dev)
        e2e_dev = {"value": world * B * max(n_e2e, 100) / t_d, "unit": UNIT, "ms_per_step": 1000.0 * t_d / max(n_e2e, 100),
                   "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4, "loss": loss_d,
                   "mode": "same graphs, inputs resident in HBM, only the loss value crosses PCIe per step"}
    except Exception as exc:
        print("device-resident e2e failed: %r" % (exc,), file=sys.stderr)

    b2b = None
    if graph is not None:
        try:
            n_b2b = max(args.steps, 200)
            ms_b2b, set_bytes = back_to_back_row(dev, rank, n_b2b)
            ms_b2b = D.reduce_scalar(ms_b2b, "max", dev)
            b2b = {"ms_per_step": ms_b2b / n_b2b, "value": world * B * n_b2b / (ms_b2b / 1000.0), "unit": UNIT, "steps": n_b2b,
                   "buffer_sets": 4, "bytes_per_set": set_bytes,
                   "l2": "no eviction fill: 4 independent buffer sets take turns (4 x %.0f MB > 126 MB L2)" % (set_bytes / 1e6),
                   "note": "same graph-replayed C-ABI step as `value`, one event pair around all steps instead of one per "
                           "step; `value` stays the per-step, explicitly evicted number"}
        except Exception as exc:
            print("back-to-back row failed: %r" % (exc,), file=sys.stderr)

    line = None
    if rank == 0:
        import json as _json
        peaks = {}
        try:
            peaks = _json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        stages = pipe.stage_times(min(args.steps, 20), flush)
        step_gbs = FULL_PATH_BYTES * B / ((ms_total / args.steps) * 1e-3) / 1e9 if world == 1 else None
        busy = pipe.kernel_timeline(graph, flush, reps=50)
        full_shape = (B == 32 and N == 8000 and V == 64)
        roofline_busy = None
        if busy:
            # each kernel's duration INSIDE the timed step (globaltimer stamps of its own CTAs, first CTA past its grid
            # dependency -> last exit; PDL chaining and graph replay intact, L2 evicted before the step): the only clock
            # whose kernel durations add up to less than the step.  `achieved` = ALGORITHMIC bytes / that duration: the
            # grids the previous kernel left in L2 are not re-read from HBM (ncu: DRAM traffic ~0.5x algorithmic for the
            # smoothing kernels), so a fraction is "algorithmic GB/s over HBM peak, L2-assisted", not DRAM utilisation.
            sbytes = dict(STAGE_BYTES)
            if fused_gather_route(pipe):
                # the gathers of the splat backward run inside the x/y pass of the backward: that kernel reads dL/d(raw)
                # twice algorithmically (pipeline + 32N of corners) and writes dL/d(tr_pc); the chain-rule kernel moves
                # only point data (pc, dL/d(tr_pc) in, d_pc out)
                sbytes["conv_xy_bwd"] = 2 * G_BYTES + G_BYTES // 32 + 32 * N + 2 * 12 * N
                sbytes["splat_bwd"] = 3 * 12 * N
            per = {k: {"busy_us": busy[k], "achieved": sbytes[k] * B / (busy[k] * 1e-6) / 1e9,
                       "frac": sbytes[k] * B / (busy[k] * 1e-6) / 1e9 / peak} for k in sbytes if k in busy}
            slow = max(per, key=lambda k: per[k]["busy_us"])
            roofline_busy = {"bound": "hbm", "unit": "GB/s", "peak": peak, "dominant_kernel": slow,
                             "achieved": per[slow]["achieved"], "frac": per[slow]["frac"], "per_kernel": per,
                             "meaning": "algorithmic bytes / in-step kernel duration / HBM peak (L2-assisted: inputs left in "
                                        "L2 by the previous kernel are not re-read from HBM)"}
            roofline = {"bound": "hbm", "kernel": slow, "achieved": per[slow]["achieved"], "peak": peak, "unit": "GB/s",
                        "frac": per[slow]["frac"], "traffic": NCU_TRAFFIC_B32.get(slow) if full_shape else None,
                        "traffic_source": NCU_TRAFFIC_SOURCE,
                        "duration_us": busy[slow],
                        "duration_source": "in-step: %globaltimer stamps of the kernel's own CTAs (first CTA past its grid "
                                           "dependency -> last exit), mean of 50 graph replays of the timed step, L2 evicted "
                                           "before each; the longest kernel of the step by that clock",
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": sbytes[slow] * B}
        else:
            dom = max((k for k in stages if k in STAGE_BYTES), key=stages.get)
            achieved = STAGE_BYTES[dom] * B / (stages[dom] * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": NCU_TRAFFIC_B32.get(dom) if full_shape else None,
                        "traffic_source": NCU_TRAFFIC_SOURCE,
                        "duration_source": "CUDA events recorded by the library around the stage (serialises the kernels)",
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": STAGE_BYTES[dom] * B}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_total / args.steps, "ms_per_step_events_around_launch": ms_outside / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "config_detail": {
                "launch": (("C-ABI step captured once in a CUDA graph, replayed per timed step; the timing events are "
                            "nodes of that graph (first / last), so the graph's launch latency behind the untimed L2 "
                            "eviction is not charged to the step -- ms_per_step_events_around_launch has it")
                           if g_ev0 is not None else
                           "C-ABI step captured once in a CUDA graph, replayed per timed step, events around the launch"
                           if graph is not None else "eager C-ABI calls"),
                "l2": flush.describe() if args.l2_flush else "no flush"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1000.0 * t_e2e / n_e2e, "loss": loss,
                    "mode": e2e_mode, "eager_pipelined_value": e2e_eager,
                    "serial_value": e2e_serial, "serial_ms_per_step": serial_ms,
                    "note": "e2e may exceed `value`: four steps are in flight (their kernels fill each other's gaps) and L2 "
                            "is not evicted between them, while `value` is one step at a time behind an L2 eviction"},
            "e2e_device_resident_inputs": e2e_dev,
            "back_to_back": b2b,
            "gpu_launches": Pipeline.LAUNCHES_PER_STEP * args.steps,
            "roofline": roofline,
            "roofline_step": {"algorithmic_bytes_per_projection": FULL_PATH_BYTES, "achieved": step_gbs,
                              "frac": (step_gbs / peak) if step_gbs else None, "unit": "GB/s"},
            "train": None,
            "stages_ms": stages,
            "kernel_busy_us": busy,
            "roofline_in_step": roofline_busy,
        }
        if world == 1 and g_ev0 is not None:
            try:
                line["variants"] = variant_rows(dev, max(args.steps, 50), flush)
            except Exception as exc:
                print("variant rows failed: %r" % (exc,), file=sys.stderr)
        if not args.no_cpu_baseline and world == 1:
            threads = pick_threads()
            kept = {}
            total, n = time_oracle(args.ref_batch, 3, 1, threads, keep=kept)
            line["cpu_baseline"] = {"value": args.ref_batch * n / total, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d fwd+bwd steps of a B=%d batch of the same workload (oracle, torch-CPU)"
                                              % (n, args.ref_batch)}
            if args.ref_batch == B:
                # the oracle ran on the very inputs of the timed step (make_inputs(B), rank 0): compare the step's results
                line["parity"] = parity_block(pipe, graph, kept)
    D.barrier()
    # BASELINE configs 3 / 4 (the train step around this path) under the same launch: every rank takes part (all-reduce).
    # Last phase, under a watchdog: a hang inside a collective cannot be caught, and the projection line must get out
    # whatever happens here.
    if not args.no_train and args.scaling == "weak":
        finished = threading.Event()

        def watchdog():
            if not finished.wait(timeout=args.train_timeout):
                if rank == 0:
                    line["train"] = {"error": "train phase did not finish within %d s (watchdog)" % args.train_timeout}
                    emit(line)
                os._exit(0)

        threading.Thread(target=watchdog, daemon=True).start()
        train = train_rows(args, rank, local_rank, world)
        finished.set()
        if rank == 0:
            line["train"] = train
    if rank == 0:
        emit(line)
    D.barrier()


def train_rows(args, rank, local_rank, world, steps=20):
    """BASELINE configs 3 / 4 under the same launch: the full train step (CNN encoder / decoder / pose ensemble under bf16
    autocast -> fp32 renderer -> losses -> one all-reduce of the flat gradient buffer -> fused Adam) as a replayed CUDA
    graph.  chair_camera_supervision: 8 objects x 4 views per GPU (weak scaling); chair_unsupervised: 16 objects in total
    (config 4: "B=16 on 8xB200"), 4 views x 4 pose candidates each, split over the ranks.  Device time (CUDA events,
    max over ranks) of `steps` steps after the warm-up / capture steps."""
    from dpc_b200 import distributed as D
    from dpc_b200.train import Trainer, synthetic_batch
    from dpc_b200.util.config import experiment_config
    dev = torch.device("cuda", local_rank)
    rows = {}
    for name in ("chair_camera_supervision", "chair_unsupervised"):
        row = {}
        try:
            cfg = experiment_config(name)
            if name == "chair_unsupervised":
                if 16 % world != 0:
                    raise RuntimeError("16 objects do not split over %d ranks" % world)
                cfg.batch_size = 16 // world
            torch.manual_seed(0)
            batch = synthetic_batch(cfg, dev, seed=rank)
            for mode in ("graph", "eager"):
                tr = Trainer(cfg, dev, ddp=world > 1, bf16=True, graph=(mode == "graph"))
                for _ in range(max(3, args.warmup)):
                    loss = tr.step(batch)
                torch.cuda.synchronize()
                D.barrier()
                n = steps if mode == "graph" else max(3, steps // 4)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    loss = tr.step(batch)
                e1.record()
                torch.cuda.synchronize()
                ms = D.reduce_scalar(e0.elapsed_time(e1), "max", dev) / n
                views = cfg.batch_size * cfg.step_size * world
                row[mode] = {"ms_per_step": ms, "object_views_per_s": views / (ms * 1e-3), "loss": float(loss)}
                if mode == "graph":
                    row.update(objects_per_gpu=cfg.batch_size, views=cfg.step_size, pose_candidates=cfg.pose_predict_num_candidates,
                               renderer_batch_per_gpu=cfg.batch_size * cfg.step_size * cfg.pose_predict_num_candidates,
                               points_kept_by_dropout=tr.n_keep(), parameters=tr.flat_p.numel(),
                               allreduce_bytes_per_step=tr.comm_bytes)
                    if world > 1:          # the step's one collective on its own: NCCL all-reduce of the flat gradient buffer
                        import torch.distributed as dist
                        for _ in range(3):
                            dist.all_reduce(tr.flat_g)
                        torch.cuda.synchronize()
                        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a0.record()
                        for _ in range(10):
                            dist.all_reduce(tr.flat_g)
                        a1.record()
                        torch.cuda.synchronize()
                        ar = D.reduce_scalar(a0.elapsed_time(a1), "max", dev) / 10
                        row["allreduce_ms"] = ar
                        row["allreduce_busbw_gbs"] = tr.comm_bytes * 2 * (world - 1) / world / (ar * 1e-3) / 1e9
                del tr
            row["dtype"] = "bf16 (CNN, autocast) + f32 (renderer, losses, optimizer state)"
            row["parallelism"] = "dp%d: objects sharded over ranks, one NCCL all-reduce of the flat fp32 gradient buffer per step" % world
        except Exception as exc:  # keep the projection line even if a train phase fails
            row["error"] = repr(exc)
            print("train row %s failed: %r" % (name, exc), file=sys.stderr)
        rows[name] = row
        torch.cuda.empty_cache()
    return rows


def run_train(args, rank, local_rank, world):
    """--workload train: only the train rows (BASELINE configs 3 / 4), as their own JSON line."""
    torch.cuda.set_device(local_rank)
    rows = train_rows(args, rank, local_rank, world, steps=args.steps)
    if rank == 0:
        sup = rows.get("chair_camera_supervision", {}).get("graph", {})
        emit({"metric": "train step: object-views/sec (chair_camera_supervision)", "value": sup.get("object_views_per_s"),
              "unit": "object-views/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
              "ms_per_step": sup.get("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "bf16 (CNN) + f32 (renderer)", "data": "synthetic", "config": {"workload": "train"}, "train": rows})
    from dpc_b200 import distributed as D
    D.barrier()


_REAL_STDOUT_FD = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner with
    printf when the box sets NCCL_DEBUG, whatever NCCL_DEBUG_FILE says), so file descriptor 1 points at stderr for the
    whole run and is put back only for the JSON line (emit())."""
    global _REAL_STDOUT_FD
    if _REAL_STDOUT_FD is None:
        sys.stdout.flush()
        _REAL_STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT_FD is not None:
        os.dup2(_REAL_STDOUT_FD, 1)
    print(json.dumps(line))
    sys.stdout.flush()
    if _REAL_STDOUT_FD is not None:
        os.dup2(2, 1)


def main():
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="projection", choices=["projection", "train"])
    ap.add_argument("--no-train", action="store_true", help="skip the train-step rows (BASELINE configs 3 / 4) of the JSON line")
    ap.add_argument("--train-timeout", type=int, default=240, help="watchdog of the train-step phase (seconds)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (the driver's contract): B=32 per GPU; strong: the B=32 batch is split over the ranks (SURVEY 8e)")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-batch", type=int, default=32, help="batch of the bounded CPU sample")
    ap.add_argument("--no-l2-flush", dest="l2_flush", action="store_false")
    ap.add_argument("--l2-flush-mode", default="write+read", choices=["write", "write+read"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager C-ABI calls instead of a CUDA-graph replay of the step")
    args = ap.parse_args()
    from dpc_b200 import distributed as D
    rank, local_rank, world = D.env_world()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    rank, local_rank, world = D.init()
    if args.scaling == "strong" and args.workload == "projection":
        if 32 % world != 0:
            raise SystemExit("--scaling strong needs a world size that divides 32")
        globals()["B"] = 32 // world          # every rank renders 32 / world samples of its own (synthetic), B = 32 in total
    if args.workload != "projection":
        run_train(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)
    if world > 1:
        # the line is out; never let the tear-down of the process group keep the job alive
        t = threading.Timer(30.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
