"""Generates tests/golden/*.npz by running the REFERENCE'S OWN SOURCE
(/root/reference/dpc/util/{point_cloud,drc,gauss_kernel,quaternion,camera}.py, unmodified)
over the TF1 shim in oracle/tf1_shim, forward and (through torch autograd underneath the
shim) backward.  Run here, in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

The fixtures travel to the GPU box; the reference does not.  Each .npz holds the case's
config overrides (json), its inputs, every output of pointcloud_project_fast, the upstream
gradients used, and the resulting input gradients.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import run_reference  # noqa: E402
from dpc_b200.util.config import default_config  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def quat_to_matrix(q, d):
    """4x4 extrinsic whose rows are ordered like the reference's internal camera
    (row 0 = depth axis) so that the matrix branch sees the cloud `d` in front."""
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    r = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)
    m = torch.zeros(q.shape[0], 4, 4)
    m[:, :3, :3] = r
    m[:, 0, 3] = d
    m[:, 3, 3] = 1.0
    return m


CASES = {
    # name: (cfg overrides, dict of input options)
    "cfg1_drc_k11": (dict(vox_size=32, pc_gauss_kernel_size=11), dict(B=2, N=1000, sigma=1.0, scale=True)),
    "cfg1_drc_k21_sigma3": (dict(vox_size=32, pc_gauss_kernel_size=21), dict(B=2, N=1000, sigma=3.0, scale=True)),
    "v64_small": (dict(vox_size=64, pc_gauss_kernel_size=21), dict(B=1, N=500, sigma=3.0, scale=True)),
    "sigma_end": (dict(vox_size=16, pc_gauss_kernel_size=21), dict(B=2, N=300, sigma=0.2, scale=True)),
    "max_proj": (dict(vox_size=16, pc_gauss_kernel_size=5, ptn_max_projection=True), dict(B=2, N=300, sigma=1.0, scale=True)),
    "trans_focal": (dict(vox_size=16, pc_gauss_kernel_size=5), dict(B=2, N=300, sigma=0.8, scale=True, trans=True, focal=True)),
    "matrix_pose": (dict(vox_size=16, pc_gauss_kernel_size=5, pose_quaternion=False), dict(B=2, N=300, sigma=0.8, scale=True, matrix=True)),
    "rgb": (dict(vox_size=16, pc_gauss_kernel_size=5, pc_rgb=True), dict(B=2, N=300, sigma=0.8, scale=True, rgb=True)),
    "rgb_options": (dict(vox_size=16, pc_gauss_kernel_size=5, pc_rgb=True, pc_rgb_stop_points_gradient=True,
                         pc_rgb_clip_after_conv=True, pc_rgb_divide_by_occupancies=True),
                    dict(B=2, N=300, sigma=0.8, scale=True, rgb=True)),
    "vox_z": (dict(vox_size=16, vox_size_z=8, pc_gauss_kernel_size=7), dict(B=2, N=300, sigma=1.0, scale=True)),
    "no_kernel_no_scale": (dict(vox_size=16), dict(B=2, N=300, sigma=None, scale=False)),
    "even_kernel": (dict(vox_size=16, pc_gauss_kernel_size=10), dict(B=2, N=300, sigma=1.5, scale=True)),
    "drc_prod": (dict(vox_size=16, pc_gauss_kernel_size=5, drc_logsum=False), dict(B=2, N=300, sigma=0.8, scale=True)),
    "clustered_init": (dict(vox_size=16, pc_gauss_kernel_size=5), dict(B=2, N=400, sigma=0.8, scale=True, spread=0.025)),
    "edge_points": (dict(vox_size=16, pc_gauss_kernel_size=5), dict(B=2, N=12, sigma=0.8, scale=True, edge=True)),
    "extra_upstream": (dict(vox_size=16, pc_gauss_kernel_size=5), dict(B=2, N=300, sigma=0.8, scale=True, upstream_all=True)),
    "single_point": (dict(vox_size=16, pc_gauss_kernel_size=5), dict(B=1, N=1, sigma=0.8, scale=True)),
    # 64^3 grid: exercises the shape-specialised FFMA2 kernels (K = 21 and K = 11, max projection,
    # gradients arriving at voxels / tr_pc / drc_probs / proj_depth as well as at proj)
    "v64_k11_max": (dict(vox_size=64, pc_gauss_kernel_size=11, ptn_max_projection=True), dict(B=1, N=400, sigma=1.0, scale=True)),
    "v64_extra_upstream": (dict(vox_size=64, pc_gauss_kernel_size=21), dict(B=1, N=400, sigma=2.0, scale=True, upstream_all=True)),
}


def build_inputs(cfg, opt, seed0=1234):
    b, n = opt["B"], opt["N"]
    spread = opt.get("spread", 0.5)
    inp = {}
    pc = torch.tanh(spread * torch.randn(b, n, 3, generator=_gen(seed0))) / 2
    q = torch.randn(b, 4, generator=_gen(seed0 + 2))
    if opt.get("edge"):
        # identity pose so the listed coordinates reach the splat untouched (up to (z+2)-2)
        q = torch.tensor([[1.0, 0, 0, 0]]).repeat(b, 1)
        pts = torch.tensor([
            [0.0, 0.0, 0.0],          # grid centre
            [-0.5, 0.0, 0.0],         # depth exactly on the lower face: valid, idx 0, frac 0
            [0.0, -0.5 * 2.5 / 1.875, 0.0],   # y lands near the face
            [0.0, 0.0, 0.6],          # projects outside -> invalid
            [float("nan"), 0.0, 0.0],  # NaN fails both comparisons -> invalid
            [0.25, 0.1, -0.1], [0.25, 0.1, -0.1], [0.25, 0.1, -0.1], [0.25, 0.1, -0.1],  # same voxel
            [0.4999, 0.0, 0.0],
            [-0.49999, 0.2, 0.2],
            [0.1, 0.2, 0.3],
        ])
        pc = pts.unsqueeze(0).repeat(b, 1, 1)
        pc[1, :, 1] *= 0.5
    inp["point_cloud"] = pc
    if opt.get("matrix"):
        inp["transform"] = quat_to_matrix(q, cfg.camera_distance)
    else:
        inp["transform"] = q
    inp["predicted_translation"] = 0.05 * torch.randn(b, 3, generator=_gen(seed0 + 5)) if opt.get("trans") else None
    inp["focal_length"] = (1.875 + 0.2 * torch.randn(b, 1, generator=_gen(seed0 + 6))) if opt.get("focal") else None
    inp["all_rgb"] = torch.rand(b, n, 3, generator=_gen(seed0 + 7)) if opt.get("rgb") else None
    inp["scaling_factor"] = torch.sigmoid(torch.randn(b, 1, generator=_gen(seed0 + 3))) if opt.get("scale") else None
    inp["sigma"] = None if opt.get("sigma") is None else torch.tensor(opt["sigma"], dtype=torch.float32)
    return inp


def run_case(ns, name):
    tf = ns.tf
    over, opt = CASES[name]
    cfg = default_config(**over)
    inp = build_inputs(cfg, opt)
    leaves = {}

    def leaf(key):
        v = inp[key]
        if v is None:
            return None
        t = v.clone().requires_grad_(True)
        leaves[key] = t
        return tf.Tensor(t)

    pc, tr = leaf("point_cloud"), leaf("transform")
    trans, focal = leaf("predicted_translation"), leaf("focal_length")
    rgb, scale = leaf("all_rgb"), leaf("scaling_factor")
    kernel = ns.gauss_kernel.smoothing_kernel(cfg, tf.Tensor(inp["sigma"])) if inp["sigma"] is not None else None
    out = ns.point_cloud.pointcloud_project_fast(cfg, pc, tr, trans, rgb, kernel, scale, focal)

    b, v = opt["B"], cfg.vox_size
    outs, ups = {}, {}
    for k, val in out.items():
        outs[k] = None if val is None else val.t
    gt = (torch.rand(b, v, v, 1, generator=_gen(1238)) > 0.5).float()
    ups["proj"] = ((outs["proj"] - gt) / b).detach()  # d/dproj of sum((gt-proj)^2)/2/B  (model_pc.py:414-415)
    if opt.get("rgb"):
        gt_rgb = torch.rand(b, v, v, 3, generator=_gen(1239))
        ups["proj_rgb"] = ((outs["proj_rgb"] - gt_rgb) / b).detach()
    if opt.get("upstream_all"):
        for k2, sc in (("voxels", 1e-2), ("tr_pc", 1.0), ("drc_probs", 1e-2), ("proj_depth", 1e-1)):
            ups[k2] = sc * torch.randn(outs[k2].shape, generator=_gen(1300 + len(k2)))
    keys = list(ups)
    names = list(leaves)
    grads = torch.autograd.grad([outs[k] for k in keys], [leaves[n] for n in names],
                                grad_outputs=[ups[k] for k in keys], allow_unused=True)
    # the integer voxel indices the reference computes (point_cloud.py:79-82), for the bit-exact gate
    vs = cfg.vox_size_z if cfg.vox_size_z != -1 else cfg.vox_size
    size = torch.tensor([[[vs, cfg.vox_size, cfg.vox_size]]], dtype=torch.float32)
    tr_pc = outs["tr_pc"].detach()
    grid = (tr_pc + 0.5) * (size - 1)
    idx = torch.floor(grid).to(torch.int32)
    valid = ((tr_pc >= -0.5) & (tr_pc <= 0.5)).all(-1)

    blob = {"cfg_json": np.array(json.dumps(over)), "opt_json": np.array(json.dumps(opt))}
    for k, val in inp.items():
        if val is not None:
            blob["in_" + k] = val.numpy()
    for k, val in outs.items():
        if val is not None:
            blob["out_" + k] = val.detach().numpy()
    if kernel is not None:
        for i, kk in enumerate(kernel):
            blob["kernel%d" % i] = kk.t.numpy()
    for k, val in ups.items():
        blob["up_" + k] = val.numpy()
    for n, g in zip(names, grads):
        if g is not None:
            blob["grad_" + n] = g.numpy()
    blob["idx"] = idx.numpy()
    blob["valid"] = valid.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
    return blob


def main():
    ns = run_reference.load()
    total = 0
    only = sys.argv[1:]
    for name in (only or CASES):
        blob = run_case(ns, name)
        path = os.path.join(OUT, name + ".npz")
        total += os.path.getsize(path)
        print("%-24s %8d bytes  valid=%d/%d  proj[min,max]=[%.4f,%.4f]" % (
            name, os.path.getsize(path), int(blob["valid"].sum()), blob["valid"].size,
            float(blob["out_proj"].min()), float(blob["out_proj"].max())))
    print("total", total)


if __name__ == "__main__":
    main()
