"""Drop-in for the reference's `dpc/util/drc.py` (ray-termination "DRC" projections).

  /root/reference/dpc/util/drc.py:105  drc_event_probabilities
  /root/reference/dpc/util/drc.py:110  drc_projection
  /root/reference/dpc/util/drc.py:126  project_volume_rgb_integral
  /root/reference/dpc/util/drc.py:139  drc_depth_grid
  /root/reference/dpc/util/drc.py:146  drc_depth_projection

drc_projection / drc_event_probabilities run in the projection kernel (one thread per ray, the
scan along depth held in registers).  The kernel evaluates p_i = u_i * prod_{j<i}(1-u_j) in
product space; the reference's log-space form (clip to [eps,1-eps], cumsum of log(1-u), exp) is
the same function including its quirks: the clip, and the factor e^eps on the first and the
terminal event because the reference's "log of one" is eps rather than 0 (drc.py:58-59).
"""
import torch

from .. import _capi
from .point_cloud import _ProjectFn, _proj_mode


def _drc_mode(cfg):
    return _capi.PROJ_DRC if cfg.drc_logsum else _capi.PROJ_DRC_PROD


def _run(voxels, cfg, want_probs):
    if voxels.dim() != 5 or voxels.shape[-1] != 1:
        raise ValueError("voxels must be [B,Vz,V,V,1]")
    proj, probs, _ = _ProjectFn.apply(voxels.squeeze(-1), _drc_mode(cfg), float(cfg.drc_logsum_clip_val),
                                      float(cfg.camera_distance), float(cfg.max_depth), False, want_probs, False)
    return proj, probs


def drc_event_probabilities(voxels, cfg):
    """p [Vz+1,B,V,V,1]: probability that the ray ends in voxel i; the last entry is 'escapes'."""
    _, probs = _run(voxels, cfg, True)
    return probs.unsqueeze(-1)


def drc_projection(voxels, cfg):
    """(silhouette [B,V,V,1], p [Vz+1,B,V,V,1]) -- sum of all events except 'escapes'."""
    proj, probs = _run(voxels, cfg, True)
    return proj.unsqueeze(-1), probs.unsqueeze(-1)


def drc_depth_grid(cfg, z_size):
    """[i/Z - 0.5 + camera_distance for i < Z] ++ [max_depth], fp32 (drc.py:139-143)."""
    z = torch.as_tensor(float(z_size), dtype=torch.float32)
    di = torch.arange(0, int(z_size), dtype=torch.float32) / z - 0.5 + cfg.camera_distance
    return torch.cat([di, torch.tensor([cfg.max_depth], dtype=torch.float32)])


def drc_depth_projection(p, cfg):
    """sum_i p_i * psi_i over the event axis; p [Vz+1,B,V,V,1]."""
    psi = drc_depth_grid(cfg, p.shape[0] - 1).to(p.device).reshape(-1, 1, 1, 1, 1)
    return (p * psi).sum(0)


class _ProjectRgbFn(torch.autograd.Function):
    @staticmethod
    @_capi.on_tensor_device
    def forward(ctx, p, rgb):
        L = _capi.lib()
        p, rgb = _capi.f32c(p), _capi.f32c(rgb)
        b, vz, v = rgb.shape[0], rgb.shape[1], rgb.shape[2]
        out = torch.empty(b, v, v, 3, dtype=torch.float32, device=rgb.device)
        _capi.check(L.dpc_project_rgb_fwd(_capi.ptr(p), _capi.ptr(rgb), b, vz, v, _capi.ptr(out), _capi.stream_of(rgb)))
        ctx.save_for_backward(p, rgb)
        return out

    @staticmethod
    @_capi.on_tensor_device
    def backward(ctx, g):
        L = _capi.lib()
        p, rgb = ctx.saved_tensors
        b, vz, v = rgb.shape[0], rgb.shape[1], rgb.shape[2]
        g = _capi.f32c(g)
        d_p = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        d_rgb = torch.empty_like(rgb) if ctx.needs_input_grad[1] else None
        if d_p is None and d_rgb is None:
            return None, None
        _capi.check(L.dpc_project_rgb_bwd(_capi.ptr(p), _capi.ptr(rgb), _capi.ptr(g), b, vz, v, _capi.ptr(d_p), _capi.ptr(d_rgb),
                                          _capi.stream_of(rgb)))
        return d_p, d_rgb


def project_volume_rgb_integral(cfg, p, rgb):
    """sum_i p_i * rgb_i with a white background for the 'escapes' event (drc.py:126-136); p [Vz+1,B,V,V,1],
    rgb [B,Vz,V,V,3] -> [B,V,V,3].  One kernel (dpc_project_rgb_fwd/bwd): the reference materialises p * [rgb, 1] first."""
    if p.dim() != 5 or rgb.dim() != 5 or rgb.shape[-1] != 3 or p.shape[0] != rgb.shape[1] + 1:
        raise ValueError("p must be [Vz+1,B,V,V,1] and rgb [B,Vz,V,V,3]")
    return _ProjectRgbFn.apply(p.squeeze(-1), rgb)
