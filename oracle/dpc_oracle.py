"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (torch-CPU eager, one rounding per op, gradients from torch autograd) of
the reference's differentiable point-cloud projection path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file; the product package (dpc_b200) never does and fails loudly without its CUDA
library.

Parity status: the reference ships no tests, golden vectors or fixtures, and TensorFlow 1.x
cannot be installed here, so the TF *runtime* numerics are "parity unpinned".  What IS
pinned: this restatement is checked (tests/test_oracle_vs_reference.py) against the
reference's own, unmodified Python source executed over oracle/tf1_shim (every tf.* op
eager fp32), bit-exact for tr_pc / voxel indices / taps and to 1e-6 elsewhere, and against
the fixtures that run produced (tests/golden/*.npz, generator committed), plus the
hand-derived known answers of SURVEY.md Appendix B.

Reference lines followed (all under /root/reference/dpc/util/):
  quaternion.py:32-45,62-83,96-117   rotate()
  point_cloud.py:157-216             pc_perspective_transform()
  camera.py:5-13                     intrinsic_matrix()
  point_cloud.py:60-136              pointcloud2voxels3d_fast()
  gauss_kernel.py:5-11,27-54         gauss_kernel_1d(), smoothing_kernel()
  point_cloud.py:139-154             smoothen_voxels3d(), convolve_rgb()
  drc.py:47-153                      drc_* projections
  point_cloud.py:229-290             pointcloud_project_fast()
  point_cloud.py:293-319             pc_point_dropout()

The functions keep the reference's names and signatures so the parity tests can call the
oracle and the product the same way.  Tensors are torch tensors (any float dtype: run in
float64 for finite-difference checks); `cfg` is any attribute-style object.
"""
import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- quaternion
def _hamilton(a, b):
    """Quaternion product with the reference's association (quaternion.py:72-78)."""
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    w = aw * bw - ax * bx - ay * by - az * bz
    x = aw * bx + ax * bw + ay * bz - az * by
    y = aw * by + ay * bw + az * bx - ax * bz
    z = aw * bz + az * bw + ax * by - ay * bx
    return torch.stack((w, x, y, z), dim=-1)


def _sqrt_rn(x):
    """Correctly rounded square root.  torch's CPU float32 sqrt is NOT IEEE-exact on this build (about
    0.7% of inputs are off by one ulp; found when a whole sample's tr_pc disagreed with the GPU's
    sqrt.rn); TensorFlow/Eigen use the hardware sqrtps, which is.  sqrt in float64 followed by one
    rounding to float32 is exact (53 >= 2*24+2)."""
    if x.dtype == torch.float32:
        return torch.sqrt(x.double()).float()
    return torch.sqrt(x)


def quaternion_rotate(pc, q):
    """q * (0,p) * conj(q) with q normalised first (quaternion.py:96-117).
    |q| = sqrt(((q0^2+q1^2)+q2^2)+q3^2): squares summed left to right -- this order
    DEFINES parity for tf.norm."""
    sq = q * q
    nrm = _sqrt_rn(sq[..., 0] + sq[..., 1] + sq[..., 2] + sq[..., 3]).unsqueeze(-1)
    qn = (q / nrm).unsqueeze(1)  # [B,1,4]
    conj = qn * torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=q.dtype)
    p4 = F.pad(pc, (1, 0))  # (0, x, y, z)  quaternion.py:45
    r = _hamilton(_hamilton(qn, p4), conj)
    return r[..., 1:4]


def intrinsic_matrix(cfg, dims=4):
    m = np.eye(dims, dtype=np.float32)
    m[1, 1] = float(cfg.focal_length)
    m[2, 2] = float(cfg.focal_length)
    return m


# --------------------------------------------------------------------------- camera transform
def pc_perspective_transform(cfg, point_cloud, transform, predicted_translation=None, focal_length=None):
    d = cfg.camera_distance
    f = cfg.focal_length if focal_length is None else focal_length.unsqueeze(-1)  # [B,1,1]
    if cfg.pose_quaternion:
        p = quaternion_rotate(point_cloud, transform)
        if predicted_translation is not None:
            p = p + predicted_translation.unsqueeze(1)
        xs, ys, zs = p[..., 2:3], p[..., 1:2], p[..., 0:1]
        zs = zs + d
        xs = xs * f
        ys = ys * f
    else:
        ones = torch.ones_like(point_cloud[..., :1])
        xyz1 = torch.cat([point_cloud, ones], dim=-1)
        k = torch.from_numpy(intrinsic_matrix(cfg, 4)).to(point_cloud.dtype)
        cam = torch.matmul(k.unsqueeze(0).expand(transform.shape[0], 4, 4), transform)
        p = torch.matmul(xyz1, cam.transpose(1, 2))
        xs, ys, zs = p[..., 2:3], p[..., 1:2], p[..., 0:1]
    xs = xs / zs
    ys = ys / zs
    zs = zs - d
    if predicted_translation is not None:
        zs = zs - predicted_translation.unsqueeze(1)[..., 0:1]
    return torch.cat([zs, ys, xs], dim=2)


# --------------------------------------------------------------------------- splat
def _grid_dims(cfg):
    v = int(cfg.vox_size)
    vz = int(cfg.vox_size_z) if int(cfg.vox_size_z) != -1 else v
    return vz, v


def voxel_indices(cfg, pc):
    """(valid[B,N] bool, idx[B,N,3] int32, frac[B,N,3]) -- point_cloud.py:76-92."""
    vz, v = _grid_dims(cfg)
    valid = torch.logical_and(pc >= -0.5, pc <= 0.5).all(dim=-1)
    size = torch.tensor([[[vz, v, v]]], dtype=pc.dtype)
    g = (pc + 0.5) * (size - 1)
    fl = torch.floor(g)
    return valid, fl.to(torch.int32), g - fl


def pointcloud2voxels3d_fast(cfg, pc, rgb):
    """Trilinear splat built the way the reference builds it: eight dense per-corner
    grids (each a masked scatter-add in point order) summed in k,j,i order
    (point_cloud.py:94-134).  Returns (voxels [B,Vz,V,V], voxels_rgb [B,Vz,V,V,3] | None)."""
    vz, v = _grid_dims(cfg)
    b, n = pc.shape[0], pc.shape[1]
    valid, idx, r = voxel_indices(cfg, pc)
    rr = (1.0 - r, r)
    flat_valid = valid.reshape(-1)
    bi = torch.arange(b).unsqueeze(1).expand(b, n).reshape(-1)[flat_valid]
    base = idx.reshape(-1, 3).long()[flat_valid]
    total, total_rgb = None, None
    for k in range(2):
        for j in range(2):
            for i in range(2):
                w_full = rr[k][..., 0] * rr[j][..., 1] * rr[i][..., 2]
                w = w_full.reshape(-1)[flat_valid]
                cz, cy, cx = base[:, 0] + k, base[:, 1] + j, base[:, 2] + i
                # A coordinate of exactly +0.5 has base index V-1 and an upper neighbour V whose
                # weight is exactly 0.  TF-CPU scatter_nd raises on it, TF-GPU drops the update;
                # the oracle (and the CUDA path) follow TF-GPU -- value-identical either way.
                inb = (cz < vz) & (cy < v) & (cx < v)
                where = (bi[inb], cz[inb], cy[inb], cx[inb])
                w = w[inb]
                grid = torch.zeros(b, vz, v, v, dtype=pc.dtype).index_put(where, w, accumulate=True)
                total = grid if total is None else total + grid
                if rgb is not None:
                    wsrc = w_full.detach() if cfg.pc_rgb_stop_points_gradient else w_full
                    wrgb = (wsrc.unsqueeze(-1) * rgb).reshape(-1, 3)[flat_valid][inb]
                    g3 = torch.zeros(b, vz, v, v, 3, dtype=pc.dtype).index_put(where, wrgb, accumulate=True)
                    total_rgb = g3 if total_rgb is None else total_rgb + g3
    return total, total_rgb


# --------------------------------------------------------------------------- smoothing taps
def gauss_kernel_1d(l, sig):
    """gauss_kernel.py:5-11.  Unary minus binds before //: l=21 -> -10..10, l=10 -> -4..5.
    `sig` may be a python number (2*sig^2 is then formed in double and rounded once, as TF
    does for a python constant) or a float tensor (fp32 ops) -- the production case,
    model_pc.py:35-40."""
    lo, hi = (-l) // 2 + 1.0, l // 2 + 1.0
    xx = torch.arange(lo, hi, dtype=torch.float32)
    if torch.is_tensor(sig):
        xx = xx.to(sig.dtype)
    k = torch.exp(-xx ** 2 / (2.0 * sig ** 2))
    return k / k.sum()


def smoothing_kernel(cfg, sigma):
    """[k1 (1,1,K,1,1), k2 (1,K,1,1,1), k3 (Kz,1,1,1,1)] -- gauss_kernel.py:35-54."""
    fsz = int(cfg.pc_gauss_kernel_size)
    k1d = gauss_kernel_1d(fsz, sigma)
    if int(cfg.vox_size_z) != -1:
        ratio = cfg.vox_size_z / cfg.vox_size
        fsz_z = int(np.floor(fsz * ratio))
        if fsz_z % 2 == 0:
            fsz_z += 1
        kz = gauss_kernel_1d(fsz_z, sigma * ratio)
    else:
        if not cfg.pc_separable_gauss_filter:
            raise NotImplementedError("non-separable smoothing is dead code in the reference (gauss_kernel.py:51-54)")
        kz, fsz_z = k1d, fsz
    return [k1d.reshape(1, 1, fsz, 1, 1), k1d.reshape(1, fsz, 1, 1, 1), kz.reshape(fsz_z, 1, 1, 1, 1)]


def _conv3d_same(x, filt):
    """tf.nn.conv3d, NDHWC, stride 1, zero 'SAME' padding ((k-1)//2 low, rest high)."""
    xin = x.permute(0, 4, 1, 2, 3)
    w = filt.to(x.dtype).permute(4, 3, 0, 1, 2).contiguous()
    pads = []
    for kk in (w.shape[4], w.shape[3], w.shape[2]):
        pads += [(kk - 1) // 2, (kk - 1) - (kk - 1) // 2]
    y = F.conv3d(F.pad(xin, pads), w)
    return y.permute(0, 2, 3, 4, 1)


def smoothen_voxels3d(cfg, voxels, kernel):
    if not cfg.pc_separable_gauss_filter:
        return _conv3d_same(voxels, kernel)
    for filt in kernel:  # axis 3 (W), then 2 (H), then 1 (D): point_cloud.py:141-142
        voxels = _conv3d_same(voxels, filt)
    return voxels


def convolve_rgb(cfg, voxels_rgb, kernel):
    chans = [voxels_rgb[..., c:c + 1] for c in range(3)]
    for filt in kernel:
        chans = [_conv3d_same(ch, filt) for ch in chans]
    return torch.cat(chans, dim=4)


# --------------------------------------------------------------------------- DRC
def drc_event_probabilities(voxels, cfg):
    """p [Z+1,B,H,W,1] -- drc.py:47-102.  Quirk kept: the log-space 'unity' is
    clip_val, not 0 (drc.py:58-59), so p_0 and p_Z carry a factor e^clip_val."""
    u = voxels.permute(1, 0, 2, 3, 4)
    eps = cfg.drc_logsum_clip_val
    one = torch.ones((1,) + tuple(u.shape[1:]), dtype=voxels.dtype)
    if cfg.drc_logsum:
        u = torch.clamp(u, eps, 1.0 - eps)
        y, x = torch.log(u), torch.log(1.0 - u)
        combine, unity, scan = torch.add, one * eps, torch.cumsum
    else:
        y, x = u, 1.0 - u
        combine, unity, scan = torch.mul, one, torch.cumprod
    if cfg.drc_tf_cumulative:
        r = scan(x, dim=0)
    else:  # drc.py:81-90, running op slice by slice
        acc = [x[0:1]]
        for i in range(1, x.shape[0]):
            acc.append(combine(x[i:i + 1], acc[-1]))
        r = torch.cat(acc, 0)
    p = combine(torch.cat([unity, r], 0), torch.cat([y, unity], 0))
    if cfg.drc_logsum:
        p = torch.exp(p)
    return p


def drc_projection(voxels, cfg):
    """(proj [B,H,W,1], p) -- drc.py:110-123: sum of all termination events but the last."""
    p = drc_event_probabilities(voxels, cfg)
    c = torch.cat([torch.ones_like(p[:-1]), torch.zeros_like(p[:1])], 0)
    return (p * c).sum(0), p


def drc_depth_projection(p, cfg):
    z = p.shape[0] - 1
    zf = torch.tensor(float(z), dtype=torch.float32)
    di = torch.arange(0, z, dtype=torch.float32) / zf - 0.5 + cfg.camera_distance
    psi = torch.cat([di, torch.tensor([cfg.max_depth], dtype=torch.float32)]).to(p.dtype)
    return (p * psi.reshape(-1, 1, 1, 1, 1)).sum(0)


def project_volume_rgb_integral(cfg, p, rgb):
    c = rgb.permute(1, 0, 2, 3, 4)
    bg = torch.ones((1,) + tuple(c.shape[1:]), dtype=rgb.dtype)
    return (p * torch.cat([c, bg], 0)).sum(0)


# --------------------------------------------------------------------------- pipeline
def pointcloud_project_fast(cfg, point_cloud, transform, predicted_translation, all_rgb,
                            kernel=None, scaling_factor=None, focal_length=None):
    """point_cloud.py:229-290; returns the same 7-key dict."""
    tr_pc = pc_perspective_transform(cfg, point_cloud, transform, predicted_translation, focal_length)
    raw, vox_rgb = pointcloud2voxels3d_fast(cfg, tr_pc, all_rgb)
    raw = raw.unsqueeze(-1)
    vox = torch.clamp(raw, 0.0, 1.0)
    if kernel is not None:
        vox = smoothen_voxels3d(cfg, vox, kernel)
        if vox_rgb is not None:
            if not cfg.pc_rgb_clip_after_conv:
                vox_rgb = torch.clamp(vox_rgb, 0.0, 1.0)
            vox_rgb = convolve_rgb(cfg, vox_rgb, kernel)
    if scaling_factor is not None:
        vox = torch.clamp(vox * scaling_factor.reshape(-1, 1, 1, 1, 1), 0.0, 1.0)
    if vox_rgb is not None:
        if cfg.pc_rgb_divide_by_occupancies:
            div = smoothen_voxels3d(cfg, raw.detach(), kernel)
            vox_rgb = vox_rgb / (div + cfg.pc_rgb_divide_by_occupancies_epsilon)
        if cfg.pc_rgb_clip_after_conv:
            vox_rgb = torch.clamp(vox_rgb, 0.0, 1.0)
    if cfg.ptn_max_projection:
        proj = torch.amax(vox, dim=1)
        drc_probs, proj_depth = None, None
    else:
        proj, drc_probs = drc_projection(vox, cfg)
        drc_probs = torch.flip(drc_probs, [2])
        proj_depth = drc_depth_projection(drc_probs, cfg)
    proj = torch.flip(proj, [1])
    if vox_rgb is not None:
        vox_rgb = torch.flip(vox_rgb, [2])
        # NB the reference has no guard here: rgb together with ptn_max_projection
        # dereferences drc_probs=None and fails (point_cloud.py:275-277).
        proj_rgb = project_volume_rgb_integral(cfg, drc_probs, vox_rgb)
    else:
        proj_rgb = None
    return {"proj": proj, "voxels": vox, "tr_pc": tr_pc, "voxels_rgb": vox_rgb,
            "proj_rgb": proj_rgb, "drc_probs": drc_probs, "proj_depth": proj_depth}


def pc_point_dropout_with_indices(points, rgb, selected):
    """Deterministic half of pc_point_dropout (point_cloud.py:312-318): gather rows
    `selected [B,n_keep]` (the reference draws them with np.random.choice on the host)."""
    b = points.shape[0]
    bi = torch.arange(b).unsqueeze(1).expand_as(selected)
    out = points[bi, selected]
    out_rgb = rgb[bi, selected] if rgb is not None else None
    return out, out_rgb


def num_points_after_dropout(num_input_points, keep_prob):
    """tf.cast(num_input_points * keep_prob, int32): fp32 product, truncation (point_cloud.py:298)."""
    return int(np.float32(num_input_points) * np.float32(keep_prob))


def projection_loss(proj, gt):
    """sum((gt-proj)^2)/2 / B -- model_pc.py:414-415 (tf.nn.l2_loss / num_samples)."""
    d = gt - proj
    return (d * d).sum() / 2 / proj.shape[0]


# ------------------------------------------------------------------ f-4: chamfer evaluation
def point_cloud_distance(Vs, Vt, chunk=512):
    """Restates util/point_cloud_distance.py:26-39: for every source point the closest target.
    diff = Vt - Vs (`:33`), dist = sqrt(sum(diff**2, axis=2)) with the three squares added left to right
    (`:34`; the order inside TF's reduce_sum is unpinned, like tf.norm), idx = first argmin over the ROUNDED
    distances (`:35`), proj = Vt[idx] (`:36`), minDist = dist[i, idx] (`:37`).  numpy in the inputs' own
    precision (np.sqrt is the correctly rounded hardware root); the [VsN, VtN] matrix is formed `chunk` sources
    at a time.  Returns torch tensors (proj, minDist, idx int32)."""
    vs = Vs.detach().cpu().numpy() if isinstance(Vs, torch.Tensor) else np.asarray(Vs)
    vt = Vt.detach().cpu().numpy() if isinstance(Vt, torch.Tensor) else np.asarray(Vt)
    assert vs.dtype == vt.dtype and vs.dtype in (np.float32, np.float64)
    ns = vs.shape[0]
    idx = np.empty(ns, dtype=np.int32)
    md = np.empty(ns, dtype=vs.dtype)
    for a in range(0, ns, chunk):
        d = vt[None, :, :] - vs[a:a + chunk, None, :]
        q = d * d
        dist = np.sqrt((q[..., 0] + q[..., 1]) + q[..., 2])
        j = np.argmin(dist, axis=1)
        idx[a:a + chunk] = j
        md[a:a + chunk] = dist[np.arange(dist.shape[0]), j]
    return torch.from_numpy(vt[idx]), torch.from_numpy(md), torch.from_numpy(idx)

