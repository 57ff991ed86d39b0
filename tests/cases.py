"""Shared helpers: load a golden fixture, run an implementation of the reference API on it
(the oracle on CPU or the product on CUDA), and compare with the parity tolerances."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INPUT_KEYS = ("point_cloud", "transform", "predicted_translation", "all_rgb", "scaling_factor", "focal_length")
OUTPUT_KEYS = ("proj", "voxels", "tr_pc", "voxels_rgb", "proj_rgb", "drc_probs", "proj_depth")

# parity tolerances (north star): indices and tr_pc bit-exact; grids / silhouettes / gradients 1e-5 abs
ATOL = 1e-5


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["cfg_over"] = json.loads(str(d.pop("cfg_json")))
    d["opt"] = json.loads(str(d.pop("opt_json")))
    return d


def make_cfg(over):
    from dpc_b200.util.config import default_config
    return default_config(**over)


def run_impl(impl, fx, device="cpu", need_grad=True, dtype=torch.float32, host_sigma=False):
    """impl: module-like with smoothing_kernel() and pointcloud_project_fast().
    host_sigma: hand sigma over as a Python float (the product then also keeps the taps on the host).
    Returns (outputs dict of detached cpu tensors, grads dict)."""
    cfg = make_cfg(fx["cfg_over"])
    leaves, args = {}, {}
    for k in INPUT_KEYS:
        a = fx.get("in_" + k)
        if a is None:
            args[k] = None
            continue
        t = torch.from_numpy(a).to(device=device, dtype=dtype)
        if need_grad:
            t.requires_grad_(True)
            leaves[k] = t
        args[k] = t
    kernel = None
    if "in_sigma" in fx:
        sig = float(fx["in_sigma"])
        kernel = impl.smoothing_kernel(cfg, sig if host_sigma else torch.tensor(sig, dtype=torch.float32, device=device))
        if dtype != torch.float32:
            kernel = [k.to(dtype) for k in kernel]
    out = impl.pointcloud_project_fast(cfg, args["point_cloud"], args["transform"], args["predicted_translation"],
                                       args["all_rgb"], kernel, args["scaling_factor"], args["focal_length"])
    grads = {}
    if need_grad:
        ups = {k[3:]: torch.from_numpy(v).to(device=device, dtype=dtype) for k, v in fx.items() if k.startswith("up_")}
        keys = [k for k in ups if out.get(k) is not None]
        names = list(leaves)
        gs = torch.autograd.grad([out[k] for k in keys], [leaves[n] for n in names],
                                 grad_outputs=[ups[k] for k in keys], allow_unused=True)
        grads = {n: g.detach().cpu() for n, g in zip(names, gs) if g is not None}
    outs = {k: (None if v is None else v.detach().cpu()) for k, v in out.items()}
    return outs, grads


def nan_equal_bits(a, b):
    """bit-exact comparison that treats any NaN as equal to any NaN (TF and CUDA propagate
    different NaN payloads) and -0.0 as different from +0.0 only if both finite paths differ."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    both_nan = np.isnan(a) & np.isnan(b)
    return np.array_equal(np.where(both_nan, 0, a), np.where(both_nan, 0, b))


def max_abs_diff(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m = ~(np.isnan(a) & np.isnan(b))
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m])))


def assert_parity(fx, outs, grads, atol=ATOL, exact_transform=None):
    """The north-star gate: tr_pc bit-exact (quaternion pose), every grid / silhouette / gradient
    within 1e-5 abs of the reference (gradients: relative to max(1, |g|_inf) for the per-sample sums)."""
    if exact_transform is None:
        exact_transform = fx["cfg_over"].get("pose_quaternion", True)
    if exact_transform:
        assert nan_equal_bits(outs["tr_pc"].numpy(), fx["out_tr_pc"]), "tr_pc must be bit-exact"
    else:
        assert max_abs_diff(outs["tr_pc"].numpy(), fx["out_tr_pc"]) <= 1e-6
    for k in OUTPUT_KEYS:
        ref = fx.get("out_" + k)
        if ref is None:
            assert outs[k] is None, k
            continue
        assert outs[k] is not None, k
        assert tuple(outs[k].shape) == ref.shape, (k, tuple(outs[k].shape), ref.shape)
        d = max_abs_diff(outs[k].numpy(), ref)
        # proj_depth is a depth in camera units (up to max_depth = 10): 1e-5 is applied relative to
        # its magnitude (one fp32 ulp at 10 is already 1e-6); everything else lives in [0,1].
        tol = atol * (max(1.0, float(np.nanmax(np.abs(ref)))) if k == "proj_depth" else 1.0)
        assert d <= tol, "%s: max abs diff %.3g" % (k, d)
    for k, v in fx.items():
        if not k.startswith("grad_"):
            continue
        name = k[5:]
        assert name in grads, "missing gradient for " + name
        g = grads[name].numpy()
        finite = np.isfinite(v)
        if not finite.all():
            # the reference propagates NaN from a NaN input point into the per-sample sums
            assert np.array_equal(np.isnan(g), np.isnan(v)) or name in ("transform",), name
        scale = max(1.0, float(np.max(np.abs(v[finite])))) if finite.any() else 1.0
        d = max_abs_diff(np.where(finite, g, 0), np.where(finite, v, 0))
        assert d <= atol * scale, "grad %s: max abs diff %.3g (scale %.3g)" % (name, d, scale)
