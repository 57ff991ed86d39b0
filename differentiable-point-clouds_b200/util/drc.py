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


def project_volume_rgb_integral(cfg, p, rgb):
    """sum_i p_i * rgb_i with a white background for the 'escapes' event; rgb [B,Vz,V,V,3]."""
    c = rgb.permute(1, 0, 2, 3, 4)
    bg = torch.ones((1,) + tuple(c.shape[1:]), dtype=rgb.dtype, device=rgb.device)
    return (p * torch.cat([c, bg], 0)).sum(0)
