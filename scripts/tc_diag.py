#!/usr/bin/env python
"""GPU diagnostics of the tensor-core smoothing kernels (dpc_smooth_tc.cuh): each check runs the public
conv entry points with the tensor-core path on (knob 8 = 1) and off (FFMA2 / generic kernels) and compares
both with a torch reference.  Prints where a delta lands when a layout is wrong."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpc_b200 import _capi  # noqa: E402

L = _capi.lib()
dev = torch.device("cuda:0")
TC = int(os.environ.get("DPC_TC", "1"))
P = _capi.ptr
st = torch.cuda.current_stream().cuda_stream


def gauss(K, sig):
    x = torch.arange(-(K // 2) + (1 if K % 2 == 0 else 0), K // 2 + 1, dtype=torch.float64)
    x = torch.arange(K, dtype=torch.float64) - (K - 1) // 2
    w = torch.exp(-x * x / (2 * sig * sig))
    return (w / w.sum()).float()


def ref_conv_axis(v, taps, axis, pl):
    """zero-padded cross-correlation along `axis` of [B,Z,Y,X] (fp64)."""
    K = taps.numel()
    v = v.double().movedim(axis, -1)
    sh = v.shape
    x = F.pad(v.reshape(-1, 1, sh[-1]), (pl, K - 1 - pl))
    y = F.conv1d(x, taps.double().view(1, 1, K).to(x.device))
    return y.reshape(sh).movedim(-1, axis)


def conv_xy(v, taps, tc, clip_in=0, pl=None):
    K = taps.numel()
    pl = (K - 1) // 2 if pl is None else pl
    B, Z, Y, X = v.shape
    out = torch.full_like(v, float("nan"))
    L.dpc_debug_set(8, tc)
    t = taps.to(dev)
    _capi.check(L.dpc_conv_xy(P(v), P(out), P(t), K, pl, P(t), K, pl, B, Z, X, clip_in, None, None, st))
    torch.cuda.synchronize()
    return out


def conv_z(v, taps, tc, scale=None, mode=-1, pl=None):
    K = taps.numel()
    pl = (K - 1) // 2 if pl is None else pl
    B, Z, Y, X = v.shape
    out = torch.full_like(v, float("nan"))
    proj = torch.full((B, Y, X), float("nan"), device=dev)
    mask2 = torch.zeros(B * Y * X * 2, dtype=torch.int32, device=dev)
    L.dpc_debug_set(8, tc)
    t = taps.to(dev)
    _capi.check(L.dpc_conv_z_fwd(P(v), P(t), K, pl, P(scale) if scale is not None else None, mode, 1e-5, 2.0, 10.0, 0,
                                 B, Z, X, P(out), P(mask2) if scale is not None else None,
                                 P(proj) if mode != -1 else None, None, None, st))
    torch.cuda.synchronize()
    return out, proj


def report(name, got, want, tol=1e-5):
    d = (got.double() - want.double()).abs()
    bad = int((~(d <= tol)).sum())
    print("%-44s max|d| = %.3e  bad = %d / %d  %s" % (name, float(d[torch.isfinite(d)].max()) if torch.isfinite(d).any() else float("nan"),
                                                     bad, d.numel(), "OK" if bad == 0 else "FAIL"), flush=True)
    return bad == 0


def where(t, n=12):
    nz = (t.abs() > 1e-6).nonzero()
    return [tuple(int(a) for a in r) + (round(float(t[tuple(r)]), 4),) for r in nz[:n]]


def main():
    torch.manual_seed(0)
    ok = True
    B = 6   # 192 depth-pass tiles: more than one per SM for the persistent kernels
    v = torch.rand(B, 64, 64, 64, device=dev)
    ident = torch.ones(1)
    # 1. identity taps: layouts only
    o, _ = conv_z(v, ident, TC)
    ok &= report("conv_z identity (tc)", o, v, 1e-6)
    if not torch.allclose(o, v, atol=1e-6):
        d = torch.zeros(B, 64, 64, 64, device=dev); d[0, 5, 3, 7] = 1.0
        od, _ = conv_z(d, ident, TC)
        print("  delta at (z5,y3,x7) lands at", where(od[0:1]))
    o = conv_xy(v, ident, TC)
    ok &= report("conv_xy identity (tc)", o, v, 1e-6)
    if not torch.allclose(o, v, atol=1e-6):
        d = torch.zeros(B, 64, 64, 64, device=dev); d[0, 5, 3, 7] = 1.0
        od = conv_xy(d, ident, TC)
        print("  delta at (z5,y3,x7) lands at", where(od[0:1]))
    # 2. shift taps
    sh = torch.tensor([1.0, 0.0, 0.0])
    o, _ = conv_z(v, sh, TC)
    ok &= report("conv_z shift taps (tc)", o, ref_conv_axis(v, sh, 1, 1), 1e-6)
    o = conv_xy(v, sh, TC)
    ok &= report("conv_xy shift taps (tc)", o, ref_conv_axis(ref_conv_axis(v, sh, 3, 1), sh, 2, 1), 1e-6)
    # 3. Gaussians, odd / even K, against fp64 and against the CUDA-core kernels
    for K, sig in ((21, 3.0), (11, 1.5), (21, 0.2), (8, 2.0), (63, 9.0)):
        t = gauss(K, sig)
        pl = (K - 1) // 2
        want = ref_conv_axis(v, t, 1, pl)
        o1, _ = conv_z(v, t, TC)
        o0, _ = conv_z(v, t, 0)
        ok &= report("conv_z K=%d sig=%g (tc vs fp64)" % (K, sig), o1, want, 2e-6)
        report("conv_z K=%d sig=%g (cuda cores vs fp64)" % (K, sig), o0, want, 2e-6)
        want = ref_conv_axis(ref_conv_axis(v, t, 3, pl), t, 2, pl)
        o1 = conv_xy(v, t, TC)
        o0 = conv_xy(v, t, 0)
        ok &= report("conv_xy K=%d sig=%g (tc vs fp64)" % (K, sig), o1, want, 2e-6)
        report("conv_xy K=%d sig=%g (cuda cores vs fp64)" % (K, sig), o0, want, 2e-6)
    # 4. scale + clip + DRC projection
    t = gauss(21, 3.0)
    sc = torch.tensor([0.7, 1.9, 1.0, 0.3, 2.5, 1.2], device=dev)
    for mode in (0, 1, 2):
        o1, p1 = conv_z(v, t, TC, scale=sc, mode=mode)
        o0, p0 = conv_z(v, t, 0, scale=sc, mode=mode)
        ok &= report("conv_z+proj mode %d voxels (tc vs cuda cores)" % mode, o1, o0, 2e-6)
        ok &= report("conv_z+proj mode %d proj   (tc vs cuda cores)" % mode, p1, p0, 5e-6)
    # 5. large magnitudes (gradients): relative accuracy
    g = torch.randn(B, 64, 64, 64, device=dev) * 1e3
    want = ref_conv_axis(ref_conv_axis(g, t, 3, 10), t, 2, 10)
    o1 = conv_xy(g, t, TC)
    ok &= report("conv_xy randn*1e3 (tc vs fp64, tol 2e-3)", o1, want, 2e-3)
    L.dpc_debug_set(8, 0)
    print("TC DIAG level %d" % TC, "PASS" if ok else "FAIL", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
