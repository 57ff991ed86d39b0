"""Drop-in for the reference's `dpc/util/point_cloud.py` hot path on B200.

Same function names, argument order and dict-of-tensors result as
  /root/reference/dpc/util/point_cloud.py:157  pc_perspective_transform
  /root/reference/dpc/util/point_cloud.py:60   pointcloud2voxels3d_fast
  /root/reference/dpc/util/point_cloud.py:139  smoothen_voxels3d
  /root/reference/dpc/util/point_cloud.py:148  convolve_rgb
  /root/reference/dpc/util/point_cloud.py:229  pointcloud_project_fast
  /root/reference/dpc/util/point_cloud.py:293  pc_point_dropout
with torch CUDA tensors in place of TF graph tensors.  All arithmetic runs in the hand-written
sm_100a kernels of csrc/ through the C-ABI (include/dpc_b200.h); this file only allocates
outputs, picks the fused or the composed route and wires autograd.  There is no CPU path.

Differences a caller can observe (all documented in DESIGN.md):
  * a coordinate of exactly +0.5 is dropped like TF-GPU does (TF-CPU raises);
  * gradients w.r.t. the smoothing taps (i.e. dL/dsigma) are produced when the kernel's tensors require them
    (sigma.requires_grad); the call then takes the composed route (separate kernels) -- sigma is a pure
    function of the step counter in the reference (model_pc.py:35-40), nothing there consumes it;
  * `pc_point_dropout` draws its subsets on the device (`dropout_indices`: a keyed pseudo-random permutation per
    sample, no sort) instead of np.random.choice on the host; the gather itself is identical.
Extension (a superset of the reference signature): `pointcloud_project_fast(..., point_indices=sel)` consumes a dropout
index list inside the splat's load stage, so the dropped points are never read or copied.
"""
import ctypes

import torch

from .. import _capi
from .point_cloud_distance import point_cloud_distance  # noqa: F401  (re-exported like point_cloud.py:7 does)
from .._capi import (POSE_MATRIX, POSE_NONE, POSE_QUAT, PROJ_DRC, PROJ_DRC_PROD, PROJ_MAX, PROJ_NONE,
                     ProjectParams, check, f32c, ptr, stream_of)


# --------------------------------------------------------------------------- helpers
def _grid_dims(cfg):
    v = int(cfg.vox_size)
    vz = int(cfg.vox_size_z) if int(cfg.vox_size_z) != -1 else v
    return vz, v


def _proj_mode(cfg):
    if cfg.ptn_max_projection:
        return PROJ_MAX
    return PROJ_DRC if cfg.drc_logsum else PROJ_DRC_PROD


def _pose_kind(cfg):
    return POSE_QUAT if cfg.pose_quaternion else POSE_MATRIX


def _check_pose(cfg, point_cloud, transform, predicted_translation, focal_length):
    if point_cloud.dim() != 3 or point_cloud.shape[-1] != 3:
        raise ValueError("point_cloud must be [B,N,3], got %s" % (tuple(point_cloud.shape),))
    b = point_cloud.shape[0]
    if cfg.pose_quaternion:
        if tuple(transform.shape) != (b, 4):
            # quaternion.validate_shape (quaternion.py:22-29)
            raise ValueError("Can't create a quaternion from a tensor with shape %s. The last dimension must be 4."
                             % (tuple(transform.shape),))
    else:
        if tuple(transform.shape) != (b, 4, 4):
            raise ValueError("camera matrix must be [B,4,4], got %s" % (tuple(transform.shape),))
        if predicted_translation is not None:
            raise ValueError("predicted_translation needs cfg.pose_quaternion (the reference fails in tf.slice here)")
    if predicted_translation is not None and tuple(predicted_translation.shape) != (b, 3):
        raise ValueError("predicted_translation must be [B,3]")
    if focal_length is not None and focal_length.numel() != b:
        raise ValueError("focal_length must be [B,1]")


def _taps_1d(k):
    return f32c(k.reshape(-1))


class SeparableKernel(list):
    """[k1, k2, k3] as the reference builds it (gauss_kernel.py:27-32,46-50) that also remembers
    the 1-D taps, so the fused path knows x and y share one tap vector.

    host_taps_xy / host_taps_z: the same taps as CPU fp32 tensors when sigma was known on the host
    (a float, or a CPU tensor): the fused path then hands them to the kernels as launch parameters
    (include/dpc_b200.h, taps_xy_host).  The device copies are made once per device and cached."""
    taps_xy = None
    taps_z = None
    host_taps_xy = None
    host_taps_z = None
    _dev_cache = None

    def device_taps(self, device):
        """(taps_xy, taps_z) on `device`."""
        if self.taps_xy.device == device:
            return self.taps_xy, self.taps_z
        if self._dev_cache is None:
            self._dev_cache = {}
        hit = self._dev_cache.get(device)
        if hit is None:
            txy = self.taps_xy.to(device)
            tz = txy if self.taps_z is self.taps_xy else self.taps_z.to(device)
            hit = self._dev_cache[device] = (txy, tz)
        return hit


def _split_kernel(kernel, device=None):
    """-> (taps_x, taps_y, taps_z, shared_xy) as 1-D fp32 tensors (on `device` when given)."""
    if isinstance(kernel, SeparableKernel) and kernel.taps_xy is not None:
        txy, tz = kernel.device_taps(device) if device is not None else (kernel.taps_xy, kernel.taps_z)
        return txy, txy, tz, True
    if not isinstance(kernel, (list, tuple)) or len(kernel) != 3:
        raise ValueError("kernel must be the list of three separable filters returned by smoothing_kernel "
                         "(the reference's non-separable branch is dead code, gauss_kernel.py:51-54)")
    kx, ky, kz = (_taps_1d(k) for k in kernel)
    shared = kx.data_ptr() == ky.data_ptr() and kx.numel() == ky.numel()
    return kx, ky, kz, shared


def _dev_taps(t, device):
    return t if t.device == device else t.to(device)


def _taps_need_grad(kernel):
    return kernel is not None and any(torch.is_tensor(k) and k.requires_grad for k in kernel)


def _grad_taps(kernel, device):
    """(tx, ty, tz) as 1-D fp32 tensors on `device` that stay connected to autograd (sigma -> taps), for _SmoothFn."""
    if _taps_need_grad(kernel):
        return tuple(_dev_taps(k.reshape(-1).to(torch.float32), device).contiguous() for k in kernel)
    tx, ty, tz, _ = _split_kernel(kernel, device)
    return _dev_taps(tx, device), _dev_taps(ty, device), _dev_taps(tz, device)


# --------------------------------------------------------------------------- K1: transform + splat
class _SplatFn(torch.autograd.Function):
    """(pc, pose, trans, focal, rgb) -> (tr_pc, vox, vox_rgb) through dpc_splat_fwd/bwd."""

    @staticmethod
    @_capi.on_tensor_device
    def forward(ctx, pc, pose, trans, focal, rgb, pose_kind, vz, v, focal_const, cam_dist, rgb_stop_grad, want_vox):
        ctx.set_materialize_grads(False)   # an output nobody differentiates arrives as None, not as a grid of zeros
        L = _capi.lib()
        pc = f32c(pc)
        pose, trans, rgb = f32c(pose), f32c(trans), f32c(rgb)
        focal = f32c(focal.reshape(-1)) if focal is not None else None
        b, n = pc.shape[0], pc.shape[1]
        tr_pc = torch.empty_like(pc) if pose_kind != POSE_NONE else None
        vox = torch.zeros(b, vz, v, v, dtype=torch.float32, device=pc.device) if want_vox else None
        vox_rgb = torch.zeros(b, vz, v, v, 3, dtype=torch.float32, device=pc.device) if (want_vox and rgb is not None) else None
        check(L.dpc_splat_fwd(ptr(pc), ptr(pose), pose_kind, ptr(trans), ptr(focal), focal_const, cam_dist, ptr(rgb),
                              b, n, vz, v, ptr(tr_pc), ptr(vox), ptr(vox_rgb), None, None, stream_of(pc)))
        ctx.save_for_backward(pc, pose, trans, focal, rgb)
        ctx.meta = (pose_kind, vz, v, focal_const, cam_dist, rgb_stop_grad)
        outs = (tr_pc if tr_pc is not None else pc.new_empty(0),
                vox if vox is not None else pc.new_empty(0),
                vox_rgb if vox_rgb is not None else pc.new_empty(0))
        ctx.mark_non_differentiable(*[o for o in outs if o.numel() == 0])
        return outs

    @staticmethod
    @_capi.on_tensor_device
    def backward(ctx, g_tr, g_vox, g_rgbvox):
        L = _capi.lib()
        pc, pose, trans, focal, rgb = ctx.saved_tensors
        pose_kind, vz, v, focal_const, cam_dist, rgb_stop_grad = ctx.meta
        b, n = pc.shape[0], pc.shape[1]
        g_tr = f32c(g_tr) if (g_tr is not None and g_tr.numel()) else None
        g_vox = f32c(g_vox) if (g_vox is not None and g_vox.numel()) else None
        g_rgbvox = f32c(g_rgbvox) if (g_rgbvox is not None and g_rgbvox.numel()) else None
        need = ctx.needs_input_grad
        d_pc = torch.empty_like(pc) if need[0] else None
        d_pose = torch.zeros_like(pose) if (pose is not None and need[1]) else None
        d_trans = torch.zeros_like(trans) if (trans is not None and need[2]) else None
        d_focal = torch.zeros_like(focal) if (focal is not None and need[3]) else None
        d_rgb = torch.empty_like(rgb) if (rgb is not None and need[4]) else None
        check(L.dpc_splat_bwd(ptr(pc), ptr(pose), pose_kind, ptr(trans), ptr(focal), focal_const, cam_dist,
                              ptr(rgb) if g_rgbvox is not None else None, int(rgb_stop_grad), b, n, vz, v,
                              ptr(g_vox), ptr(g_rgbvox), ptr(g_tr),
                              ptr(d_pc), ptr(d_pose), ptr(d_trans), ptr(d_focal),
                              ptr(d_rgb) if g_rgbvox is not None else None, stream_of(pc)))
        if d_rgb is not None and g_rgbvox is None:
            d_rgb.zero_()
        if d_focal is not None:
            d_focal = d_focal.reshape(b, 1)
        return (d_pc, d_pose, d_trans, d_focal, d_rgb) + (None,) * 7


def pc_perspective_transform(cfg, point_cloud, transform, predicted_translation=None, focal_length=None):
    """
    :param point_cloud: [B, N, 3]
    :param transform: [B, 4] if quaternion (unnormalised, w first) or [B, 4, 4] if camera matrix
    :param predicted_translation: [B, 3] translation vector
    :return: [B, N, 3], channels (depth, y, x)
    """
    _check_pose(cfg, point_cloud, transform, predicted_translation, focal_length)
    vz, v = _grid_dims(cfg)
    tr_pc, _, _ = _SplatFn.apply(point_cloud, transform, predicted_translation, focal_length, None,
                                 _pose_kind(cfg), vz, v, float(cfg.focal_length), float(cfg.camera_distance),
                                 False, False)
    return tr_pc


def pointcloud2voxels3d_fast(cfg, pc, rgb):  # [B,N,3]
    """Trilinear splat of camera-space points. Returns (voxels [B,Vz,V,V], voxels_rgb [B,Vz,V,V,3] | None)."""
    if pc.dim() != 3 or pc.shape[-1] != 3:
        raise ValueError("pc must be [B,N,3]")
    vz, v = _grid_dims(cfg)
    _, vox, vox_rgb = _SplatFn.apply(pc, None, None, None, rgb, POSE_NONE, vz, v, float(cfg.focal_length),
                                     float(cfg.camera_distance), bool(cfg.pc_rgb_stop_points_gradient), True)
    return vox, (vox_rgb if rgb is not None else None)


# --------------------------------------------------------------------------- K2: separable smoothing
def _conv_xy(L, x, tx, ty, plx, ply, clip_in=False, mask_out=None, mask_in=None):
    b, vz, v = x.shape[0], x.shape[1], x.shape[2]
    out = torch.empty_like(x)
    check(L.dpc_conv_xy(ptr(x), ptr(out), ptr(tx), tx.numel(), plx, ptr(ty), ty.numel(), ply, b, vz, v,
                        int(clip_in), ptr(mask_out), ptr(mask_in), stream_of(x)))
    return out


def _tap_corr(L, a, g, axis, k, pad_lo):
    """dL/d(taps) of one zero-padded correlation pass: a = its input, g = the gradient at its output (dpc_tap_corr)."""
    b, vz, v = a.shape[0], a.shape[1], a.shape[2]
    out = torch.empty(k, dtype=torch.float32, device=a.device)
    check(L.dpc_tap_corr(ptr(a), ptr(g), axis, b, vz, v, k, pad_lo, ptr(out), stream_of(a)))
    return out


class _SmoothFn(torch.autograd.Function):
    """[B,Vz,V,V] -> same: correlate along x, then y, then depth (point_cloud.py:141-142).  Differentiable w.r.t. the
    grid and -- when a tap tensor requires a gradient, i.e. when sigma does -- w.r.t. the taps of every pass (N1)."""

    @staticmethod
    @_capi.on_tensor_device
    def forward(ctx, vox, tx, ty, tz):
        L = _capi.lib()
        vox = f32c(vox)
        b, vz, v = vox.shape[0], vox.shape[1], vox.shape[2]
        tmp = _conv_xy(L, vox, tx, ty, (tx.numel() - 1) // 2, (ty.numel() - 1) // 2)
        out = torch.empty_like(vox)
        kz = tz.numel()
        check(L.dpc_conv_z_fwd(ptr(tmp), ptr(tz), kz, (kz - 1) // 2, None, PROJ_NONE, 0.0, 0.0, 0.0, 0, b, vz, v,
                               ptr(out), None, None, None, None, stream_of(vox)))
        want_taps = any(ctx.needs_input_grad[1:4])
        ctx.want_taps = want_taps
        # the tap gradients need the input of every pass: the grid itself and the x/y-smoothed grid (the x-smoothed one
        # is recomputed in the backward)
        ctx.save_for_backward(tx, ty, tz, *((vox, tmp) if want_taps else ()))
        return out

    @staticmethod
    @_capi.on_tensor_device
    def backward(ctx, g):
        L = _capi.lib()
        tx, ty, tz = ctx.saved_tensors[:3]
        g = f32c(g)
        b, vz, v = g.shape[0], g.shape[1], g.shape[2]
        rx, ry, rz = tx.flip(0).contiguous(), ty.flip(0).contiguous(), tz.flip(0).contiguous()
        kx, ky, kz = tx.numel(), ty.numel(), tz.numel()
        h2 = torch.empty_like(g)
        # mode NONE: the "voxels" operand is never read for values, only d(out) = g flows
        check(L.dpc_conv_z_bwd(ptr(g), None, None, ptr(rz), kz, kz - 1 - (kz - 1) // 2, PROJ_NONE, 0.0, 0.0, 0.0, 0,
                               b, vz, v, None, ptr(g), None, None, ptr(h2), None, stream_of(g)))
        if not ctx.want_taps:
            d = _conv_xy(L, h2, rx, ry, kx - 1 - (kx - 1) // 2, ky - 1 - (ky - 1) // 2)
            return d, None, None, None
        a0, a2 = ctx.saved_tensors[3:5]
        one = torch.ones(1, dtype=torch.float32, device=g.device)          # a 1-tap identity filter
        a1 = _conv_xy(L, a0, tx, one, (kx - 1) // 2, 0)                      # input of the y pass
        gx = _conv_xy(L, h2, one, ry, 0, ky - 1 - (ky - 1) // 2)             # gradient at the output of the x pass
        d = _conv_xy(L, gx, rx, one, kx - 1 - (kx - 1) // 2, 0)
        need = ctx.needs_input_grad
        dtz = _tap_corr(L, a2, g, 0, kz, (kz - 1) // 2) if need[3] else None
        dty = _tap_corr(L, a1, h2, 1, ky, (ky - 1) // 2) if need[2] else None
        dtx = _tap_corr(L, a0, gx, 2, kx, (kx - 1) // 2) if need[1] else None
        return d, dtx, dty, dtz


def smoothen_voxels3d(cfg, voxels, kernel):
    """voxels [B,Vz,V,V,1] -> same shape; three separable zero-padded correlations."""
    if not cfg.pc_separable_gauss_filter:
        raise NotImplementedError("only the separable filter exists (the reference's dense branch is unreachable, "
                                  "gauss_kernel.py:51-54)")
    if voxels.dim() != 5 or voxels.shape[-1] != 1:
        raise ValueError("voxels must be [B,Vz,V,V,1]")
    dev = voxels.device
    tx, ty, tz = _grad_taps(kernel, dev)
    out = _SmoothFn.apply(voxels.squeeze(-1), tx, ty, tz)
    return out.unsqueeze(-1)


def convolve_rgb(cfg, voxels_rgb, kernel):
    """[B,Vz,V,V,3] -> same: each colour channel smoothed separately (point_cloud.py:148-154)."""
    dev = voxels_rgb.device
    tx, ty, tz = _grad_taps(kernel, dev)
    b = voxels_rgb.shape[0]
    chans = voxels_rgb.permute(4, 0, 1, 2, 3).reshape((3 * b,) + tuple(voxels_rgb.shape[1:4])).contiguous()
    out = _SmoothFn.apply(chans, tx, ty, tz)
    return out.reshape((3, b) + tuple(voxels_rgb.shape[1:4])).permute(1, 2, 3, 4, 0).contiguous()


# --------------------------------------------------------------------------- K3: projection alone
class _ProjectFn(torch.autograd.Function):
    """voxels [B,Vz,V,V] -> (proj [B,V,V], probs [Vz+1,B,V,V] | empty, depth | empty), no smoothing."""

    @staticmethod
    @_capi.on_tensor_device
    def forward(ctx, vox, mode, eps, cam_dist, max_depth, flip_y, want_probs, want_depth):
        ctx.set_materialize_grads(False)
        L = _capi.lib()
        vox = f32c(vox)
        b, vz, v = vox.shape[0], vox.shape[1], vox.shape[2]
        dev = vox.device
        one = torch.ones(1, dtype=torch.float32, device=dev)
        out = torch.empty_like(vox)
        proj = torch.empty(b, v, v, dtype=torch.float32, device=dev)
        drc = mode in (PROJ_DRC, PROJ_DRC_PROD)
        probs = torch.empty(vz + 1, b, v, v, dtype=torch.float32, device=dev) if (drc and want_probs) else None
        depth = torch.empty(b, v, v, dtype=torch.float32, device=dev) if (drc and want_depth) else None
        check(L.dpc_conv_z_fwd(ptr(vox), ptr(one), 1, 0, None, mode, eps, cam_dist, max_depth, int(flip_y), b, vz, v,
                               ptr(out), None, ptr(proj), ptr(probs), ptr(depth), stream_of(vox)))
        ctx.save_for_backward(vox, one)
        ctx.meta = (mode, eps, cam_dist, max_depth, flip_y)
        e = vox.new_empty(0)
        outs = (proj, probs if probs is not None else e, depth if depth is not None else e)
        ctx.mark_non_differentiable(*[o for o in outs if o.numel() == 0])
        return outs

    @staticmethod
    @_capi.on_tensor_device
    def backward(ctx, g_proj, g_probs, g_depth):
        L = _capi.lib()
        vox, one = ctx.saved_tensors
        mode, eps, cam_dist, max_depth, flip_y = ctx.meta
        b, vz, v = vox.shape[0], vox.shape[1], vox.shape[2]
        g_proj = f32c(g_proj) if g_proj is not None else None
        g_probs = f32c(g_probs) if (g_probs is not None and g_probs.numel()) else None
        g_depth = f32c(g_depth) if (g_depth is not None and g_depth.numel()) else None
        d = torch.empty_like(vox)
        check(L.dpc_conv_z_bwd(ptr(vox), None, None, ptr(one), 1, 0, mode, eps, cam_dist, max_depth, int(flip_y),
                               b, vz, v, ptr(g_proj), None, ptr(g_probs), ptr(g_depth), ptr(d), None, stream_of(vox)))
        return (d,) + (None,) * 7


# --------------------------------------------------------------------------- the fused pipeline
# Scratch grids of the fused path (raw + one intermediate, 2 x B*Vz*V*V*4 bytes) carry no state
# between calls, so one buffer per (device, stream, size) is kept and reused instead of being
# allocated per call.  Stream-ordered reuse is safe; a different stream gets its own buffer.
_SCRATCH = {}
_SCRATCH_MAX_BYTES = 2 << 30      # 2 GiB of cached scratch at most (one B=32, 64^3 entry is 67 MiB)


def _scratch_for(device, stream, nbytes):
    key = (device.index if device.type == "cuda" else -1, stream, nbytes)
    if device.type == "cuda" and torch.cuda.is_current_stream_capturing():
        # inside a CUDA-graph capture the buffer comes from the graph's own memory pool and is NOT cached: the graph bakes
        # the pointer in, so its lifetime has to be the graph's, not this cache's
        return key, torch.empty(nbytes, dtype=torch.uint8, device=device)
    buf = _SCRATCH.get(key)
    if buf is None:
        # eager calls: streams and batch shapes come and go -- bound the cache by entries and by bytes (dropping an entry
        # is safe: it was allocated on the stream it is used on, the caching allocator orders its reuse)
        held = sum(b.numel() for b in _SCRATCH.values())
        if len(_SCRATCH) >= 32 or held + nbytes > _SCRATCH_MAX_BYTES:
            _SCRATCH.clear()
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _SCRATCH[key] = buf
    return key, buf


def release_scratch():
    """Drop the cached scratch grids (they are re-created, zeroed, on the next call)."""
    _SCRATCH.clear()


class _ProjectFastFn(torch.autograd.Function):
    """pointcloud_project_fast without rgb: K1 -> K2a -> K2b+K3 in three launches, backward in
    three; the clip masks travel as bit planes in a small per-call `saved` buffer."""

    @staticmethod
    @_capi.on_tensor_device
    def forward(ctx, pc, pose, trans, focal, scale, taps_xy, taps_z, params, sel=None):
        # Without this autograd hands the backward a 32 MiB grid of zeros for `voxels` and zeros for `tr_pc` whenever
        # the loss uses only `proj` (the training case): two fill launches, and the general backward kernel instead of
        # the silhouette-only one (34 vs 14 us at B=32; found in the ncu launch list of the e2e step).
        ctx.set_materialize_grads(False)
        L = _capi.lib()
        pc, pose, trans = f32c(pc), f32c(pose), f32c(trans)
        focal = f32c(focal.reshape(-1)) if focal is not None else None
        scale = f32c(scale.reshape(-1)) if scale is not None else None
        dev = pc.device
        b, n, vz, v = params.B, params.N, params.Vz, params.V
        scratch_bytes = L.dpc_project_fast_scratch_bytes(ctypes.byref(params))
        saved_bytes = L.dpc_project_fast_saved_bytes(ctypes.byref(params))
        if scratch_bytes < 0 or saved_bytes < 0:
            raise ValueError("dpc_b200: unsupported shape for the fused path: B=%d N=%d Vz=%d V=%d K=%d"
                             % (b, n, vz, v, params.K))
        stream = stream_of(pc)
        key, scratch = _scratch_for(dev, stream, scratch_bytes)
        params.flags = 0   # the forward zeroes the raw grid itself (the memset also warms L2 for the splat)
        saved = torch.empty(saved_bytes, dtype=torch.uint8, device=dev)
        if sel is not None:          # f-2: the splat reads pc[b, sel[b, i]]; params.N = sel.shape[1], params.N_src = pc.shape[1]
            params.sel, params.N_src = ptr(sel), pc.shape[1]
        tr_pc = torch.empty(b, n, 3, dtype=torch.float32, device=dev)
        voxels = torch.empty(b, vz, v, v, dtype=torch.float32, device=dev)
        proj = torch.empty(b, v, v, dtype=torch.float32, device=dev)
        # drc_probs / proj_depth are NOT written here: the training loss consumes only `proj`
        # (default_config.yaml:111-115) and the event tensor is another full grid of HBM traffic.
        # ProjectionOutputs derives them from `voxels` on first access.
        try:
            check(L.dpc_project_fast_fwd(ctypes.byref(params), ptr(pc), ptr(pose), ptr(trans), ptr(focal), ptr(scale),
                                         ptr(taps_xy), ptr(taps_z), ptr(tr_pc), ptr(voxels), ptr(proj), None,
                                         None, ptr(scratch), scratch_bytes, ptr(saved), saved_bytes, stream))
        except Exception:
            _SCRATCH.pop(key, None)
            raise
        ctx.save_for_backward(pc, pose, trans, focal, scale, taps_xy, taps_z, voxels, saved, tr_pc, sel)
        ctx.params = params
        return tr_pc, voxels, proj

    @staticmethod
    @_capi.on_tensor_device
    def backward(ctx, g_tr, g_vox, g_proj):
        L = _capi.lib()
        pc, pose, trans, focal, scale, taps_xy, taps_z, voxels, saved, tr_pc, sel = ctx.saved_tensors
        params = ctx.params
        params.tr_pc = ptr(tr_pc)      # the cells the forward used: the splat backward prefetches its gathers from them
        b = params.B

        def opt(g):
            return f32c(g) if (g is not None and g.numel()) else None

        g_tr, g_vox, g_proj = opt(g_tr), opt(g_vox), opt(g_proj)
        need = ctx.needs_input_grad
        d_pc = torch.empty_like(pc) if need[0] else None
        d_pose = torch.empty_like(pose) if (pose is not None and need[1]) else None
        d_trans = torch.empty_like(trans) if (trans is not None and need[2]) else None
        d_focal = torch.empty_like(focal) if (focal is not None and need[3]) else None
        d_scale = torch.empty_like(scale) if (scale is not None and need[4]) else None
        stream = stream_of(pc)
        scratch_bytes = L.dpc_project_fast_scratch_bytes(ctypes.byref(params))
        _, scratch = _scratch_for(pc.device, stream, scratch_bytes)
        check(L.dpc_project_fast_bwd(ctypes.byref(params), ptr(pc), ptr(pose), ptr(trans), ptr(focal), ptr(scale),
                                     ptr(taps_xy), ptr(taps_z), ptr(voxels),
                                     ptr(g_proj), ptr(g_vox), ptr(g_tr), None, None,
                                     ptr(d_pc), ptr(d_pose), ptr(d_trans), ptr(d_focal), ptr(d_scale),
                                     ptr(scratch), scratch_bytes, ptr(saved), saved.numel(), stream))
        if d_focal is not None:
            d_focal = d_focal.reshape(b, 1)
        if d_scale is not None:
            d_scale = d_scale.reshape(b, 1)
        return d_pc, d_pose, d_trans, d_focal, d_scale, None, None, None, None


class ProjectionOutputs(dict):
    """The reference's seven-key result dict.  In the TF graph unused outputs cost nothing; here
    `drc_probs` and `proj_depth` are computed from `voxels` (projection kernel, autograd-connected)
    the first time either key is read, so a caller that only uses `proj` never pays for them."""
    _LAZY = ("drc_probs", "proj_depth")

    def __init__(self, base, thunk):
        super().__init__(base)
        self._thunk = thunk

    def _materialize(self):
        if self._thunk is not None:
            thunk, self._thunk = self._thunk, None
            probs, depth = thunk()
            super().__setitem__("drc_probs", probs)
            super().__setitem__("proj_depth", depth)

    def __getitem__(self, key):
        if key in self._LAZY:
            self._materialize()
        return super().__getitem__(key)

    def get(self, key, default=None):
        if key in self._LAZY:
            self._materialize()
        return super().get(key, default)

    def items(self):
        self._materialize()
        return super().items()

    def values(self):
        self._materialize()
        return super().values()

    def copy(self):
        self._materialize()
        return dict(self)

    # dict(out), {**out}, out.keys(): CPython copies a dict subclass through keys() + __getitem__ only when keys() is
    # overridden; materialise there so that a copy never carries the lazy placeholders (None) of the fused route
    def keys(self):
        self._materialize()
        return super().keys()

    def __iter__(self):
        self._materialize()
        return super().__iter__()


def _fused_supported(cfg, point_cloud, kernel_parts):
    vz, v = _grid_dims(cfg)
    if (v * v) % 32 != 0 or v > _capi.MAX_V or vz > _capi.MAX_V:
        return False
    if kernel_parts is not None:
        tx, ty, tz, shared = kernel_parts
        if not shared or tx.numel() > _capi.MAX_TAPS or tz.numel() > _capi.MAX_TAPS:
            return False
    return True


def pointcloud_project_fast(cfg, point_cloud, transform, predicted_translation,
                            all_rgb, kernel=None, scaling_factor=None, focal_length=None, point_indices=None):
    """The reference's projection pipeline (point_cloud.py:229-290): camera transform -> trilinear
    splat -> clip -> separable smoothing -> * scaling_factor -> clip -> DRC (or max) projection
    along depth -> row flip.  Returns the same dict: proj [B,V,V,1], voxels [B,Vz,V,V,1] (not
    flipped), tr_pc [B,N,3], voxels_rgb, proj_rgb, drc_probs [Vz+1,B,V,V,1], proj_depth [B,V,V,1]
    (None where the reference returns None).
    point_indices (extension, [B,k] integer, distinct per sample; see dropout_indices): render only the points
    point_cloud[b, point_indices[b]] -- pc_point_dropout folded into the splat's load stage; tr_pc is then [B,k,3] and
    the gradient of a dropped point is zero."""
    _check_pose(cfg, point_cloud, transform, predicted_translation, focal_length)
    b, n = point_cloud.shape[0], point_cloud.shape[1]
    sel = None
    if point_indices is not None:
        if point_indices.dim() != 2 or point_indices.shape[0] != b or point_indices.shape[1] > n:
            raise ValueError("point_indices must be [B,k] with k <= N")
        if all_rgb is not None or not _fused_supported(cfg, point_cloud, None) or _taps_need_grad(kernel):
            # routes without an indexed load stage: materialise the subset first (the reference's own order of operations)
            point_cloud, all_rgb = pc_point_dropout(point_cloud, all_rgb, None, selected_indices=point_indices)
        else:
            sel = point_indices.to(device=point_cloud.device, dtype=torch.int32).contiguous()
        n = point_indices.shape[1]
    vz, v = _grid_dims(cfg)
    dev = point_cloud.device
    if scaling_factor is not None and scaling_factor.numel() != b:
        raise ValueError("scaling_factor must be [B,1]")
    parts = None
    host_xy = host_z = None
    if kernel is not None:
        if not cfg.pc_separable_gauss_filter:
            raise NotImplementedError("only the separable filter exists (gauss_kernel.py:51-54)")
        tx, ty, tz, shared = _split_kernel(kernel, dev)
        parts = (_dev_taps(tx, dev), _dev_taps(ty, dev), _dev_taps(tz, dev), shared)
        host_xy = getattr(kernel, "host_taps_xy", None)
        host_z = getattr(kernel, "host_taps_z", None)
    if _taps_need_grad(kernel):
        # N1, dL/dsigma: the composed route (its smoothing Function returns the gradients of the taps of every pass)
        return _project_composed(cfg, point_cloud, transform, predicted_translation, all_rgb, kernel,
                                 _grad_taps(kernel, dev) + (False,), scaling_factor, focal_length)
    if all_rgb is None and _fused_supported(cfg, point_cloud, parts):
        params = ProjectParams(B=b, N=n, Vz=vz, V=v, pose_kind=_pose_kind(cfg), mode=_proj_mode(cfg),
                               K=parts[0].numel() if parts else 0, Kz=parts[2].numel() if parts else 0,
                               focal_const=float(cfg.focal_length), cam_dist=float(cfg.camera_distance),
                               clip_eps=float(cfg.drc_logsum_clip_val), max_depth=float(cfg.max_depth))
        if parts and host_xy is not None and host_z is not None:
            # CPU copies of the taps: read by the C-ABI during each call, kept alive on `params`
            params.taps_xy_host, params.taps_z_host = host_xy.data_ptr(), host_z.data_ptr()
            params._host_taps = (host_xy, host_z)
        tr_pc, voxels, proj = _ProjectFastFn.apply(
            point_cloud, transform, predicted_translation, focal_length, scaling_factor,
            parts[0] if parts else None, parts[2] if parts else None, params, sel)
        base = {"proj": proj.unsqueeze(-1), "voxels": voxels.unsqueeze(-1), "tr_pc": tr_pc,
                "voxels_rgb": None, "proj_rgb": None, "drc_probs": None, "proj_depth": None}
        if cfg.ptn_max_projection:
            return base
        mode, eps = _proj_mode(cfg), float(cfg.drc_logsum_clip_val)
        cd, md = float(cfg.camera_distance), float(cfg.max_depth)

        def events():
            _, probs, depth = _ProjectFn.apply(voxels, mode, eps, cd, md, True, True, True)
            return probs.unsqueeze(-1), depth.unsqueeze(-1)

        return ProjectionOutputs(base, events)
    return _project_composed(cfg, point_cloud, transform, predicted_translation, all_rgb, kernel, parts,
                             scaling_factor, focal_length)


def _project_composed(cfg, point_cloud, transform, predicted_translation, all_rgb, kernel, parts,
                      scaling_factor, focal_length):
    """Same pipeline from the individual kernels (rgb products, x/y taps that differ, odd grid
    sizes).  Element-wise glue between kernels is plain torch on the device."""
    from . import drc as drc_mod
    b = point_cloud.shape[0]
    vz, v = _grid_dims(cfg)
    has_rgb = all_rgb is not None
    tr_pc, raw, vox_rgb = _SplatFn.apply(point_cloud, transform, predicted_translation, focal_length, all_rgb,
                                         _pose_kind(cfg), vz, v, float(cfg.focal_length), float(cfg.camera_distance),
                                         bool(cfg.pc_rgb_stop_points_gradient), True)
    vox = torch.clamp(raw, 0.0, 1.0)
    if parts is not None:
        vox = _SmoothFn.apply(vox, parts[0], parts[1], parts[2])
        if has_rgb:
            if not cfg.pc_rgb_clip_after_conv:
                vox_rgb = torch.clamp(vox_rgb, 0.0, 1.0)
            vox_rgb = convolve_rgb(cfg, vox_rgb, kernel)
    if scaling_factor is not None:
        vox = torch.clamp(vox * scaling_factor.reshape(b, 1, 1, 1), 0.0, 1.0)
    if has_rgb:
        if cfg.pc_rgb_divide_by_occupancies:
            div = _SmoothFn.apply(raw.detach(), parts[0], parts[1], parts[2])
            vox_rgb = vox_rgb / (div.unsqueeze(-1) + cfg.pc_rgb_divide_by_occupancies_epsilon)
        if cfg.pc_rgb_clip_after_conv:
            vox_rgb = torch.clamp(vox_rgb, 0.0, 1.0)
    mode = _proj_mode(cfg)
    drc = mode != PROJ_MAX
    proj, probs, depth = _ProjectFn.apply(vox, mode, float(cfg.drc_logsum_clip_val), float(cfg.camera_distance),
                                          float(cfg.max_depth), True, drc, drc)
    proj_rgb = None
    if has_rgb:
        vox_rgb = torch.flip(vox_rgb, [2])
        if not drc:
            raise TypeError("rgb projection needs the DRC probabilities; the reference fails the same way with "
                            "ptn_max_projection (point_cloud.py:275-277)")
        proj_rgb = drc_mod.project_volume_rgb_integral(cfg, probs.unsqueeze(-1), vox_rgb)
    return {"proj": proj.unsqueeze(-1), "voxels": vox.unsqueeze(-1), "tr_pc": tr_pc,
            "voxels_rgb": vox_rgb if has_rgb else None, "proj_rgb": proj_rgb,
            "drc_probs": probs.unsqueeze(-1) if drc else None,
            "proj_depth": depth.unsqueeze(-1) if drc else None}


# --------------------------------------------------------------------------- f-2: point dropout
class _GatherFn(torch.autograd.Function):
    @staticmethod
    @_capi.on_tensor_device
    def forward(ctx, x, sel):
        L = _capi.lib()
        x = f32c(x)
        b, n, c = x.shape
        k = sel.shape[1]
        out = torch.empty(b, k, c, dtype=torch.float32, device=x.device)
        check(L.dpc_gather_points(ptr(x), ptr(sel), b, n, k, c, ptr(out), stream_of(x)))
        ctx.save_for_backward(sel)
        ctx.shape = (b, n, c)
        return out

    @staticmethod
    @_capi.on_tensor_device
    def backward(ctx, g):
        L = _capi.lib()
        (sel,) = ctx.saved_tensors
        b, n, c = ctx.shape
        g = f32c(g)
        d = torch.empty(b, n, c, dtype=torch.float32, device=g.device)
        check(L.dpc_gather_points_bwd(ptr(g), ptr(sel), b, n, sel.shape[1], c, ptr(d), stream_of(g)))
        return d, None


def num_points_after_dropout(num_input_points, keep_prob):
    """tf.cast(num_input_points * keep_prob, tf.int32) in fp32 (point_cloud.py:298)."""
    kp = torch.as_tensor(keep_prob, dtype=torch.float32).reshape(()).cpu()
    return int((torch.tensor(float(num_input_points), dtype=torch.float32) * kp).item())


_DRAWS = [0]


def dropout_indices(batch, num_points, n_keep, device, seed=None, draw=None, state=None):
    """[batch, n_keep] int32 on `device`: for every sample a uniformly keyed pseudo-random subset of n_keep DISTINCT
    indices of [0, num_points) -- what np.random.choice(N, n_keep, replace=False) draws per sample in the reference
    (point_cloud.py:296-311), generated by `dpc_dropout_indices` (Philox-keyed Feistel permutation; no noise tensor,
    no sort).  seed defaults to torch's initial seed, draw to a process-wide counter; state (device uint64 [2] =
    {seed, draw}) makes the kernel read both on the device (CUDA-graph replays)."""
    L = _capi.lib()
    dev = torch.device(device)
    sel = torch.empty(batch, n_keep, dtype=torch.int32, device=dev)
    if seed is None:
        seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    if draw is None:
        draw = _DRAWS[0]
        _DRAWS[0] += 1
    stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else None
    check(L.dpc_dropout_indices(int(seed), int(draw), ptr(state), batch, num_points, n_keep, ptr(sel), stream))
    return sel


def pc_point_dropout(points, rgb, keep_prob, generator=None, selected_indices=None):
    """Keep a random subset of int(N*keep_prob) points per sample, without replacement, same subset
    for points and rgb (point_cloud.py:293-319).  The reference draws with np.random.choice in a
    tf.py_func on the host; here the subsets come from `dropout_indices` on the device (pass
    `selected_indices` [B,n_keep] to inject a specific subset; `generator` seeds the draw).  This is the materialising
    form (a gather kernel); the training step hands the index list to pointcloud_project_fast instead."""
    b, n, _ = points.shape
    if selected_indices is None:
        k = num_points_after_dropout(n, keep_prob)
        seed = generator.initial_seed() if generator is not None else None
        selected_indices = dropout_indices(b, n, k, points.device, seed=seed)
    sel = selected_indices.to(device=points.device, dtype=torch.int64).contiguous()
    out_points = _GatherFn.apply(points, sel)
    out_rgb = _GatherFn.apply(rgb, sel) if rgb is not None else None
    return out_points, out_rgb


def subsample_points(xyz, num_points):
    idxs = torch.randint(0, xyz.shape[0], (num_points,), device=xyz.device)
    return xyz[idxs, :]
