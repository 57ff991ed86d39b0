"""Loss terms of the training step (reference: dpc/util/losses.py) and the silhouette loss as one kernel.

  /root/reference/dpc/util/losses.py:25-33   drc_loss
  /root/reference/dpc/util/losses.py:53-69   add_drc_loss
  /root/reference/dpc/util/losses.py:72-93   add_proj_rgb_loss
  /root/reference/dpc/util/losses.py:116-140 add_proj_depth_loss
Same names and argument order (cfg, inputs, outputs, weight_scale, ...); `add_summary` is accepted and ignored (there
is no summary writer here).  The resizes are the TF 1.x ones (align_corners=False, no half-pixel centres), for which an
integer reduction factor is exact sub-sampling; other factors are not needed by the reference's configs and raise.

Mirrors `proj_loss = tf.nn.l2_loss(gt - pred); proj_loss /= tf.to_float(num_samples)`
(/root/reference/dpc/models/model_pc.py:414-415; `tf.nn.l2_loss(x) = sum(x**2) / 2`): forward value and the gradient
w.r.t. the prediction come out of the same pass (`dpc_proj_l2_loss` of the C-ABI), so the step between the forward and
the backward of the projection path is one PDL-aware launch instead of a handful of element-wise ones.
"""
import torch

from .. import _capi

_WORK = {}


def _workspace(device, stream):
    """Partial sums + completion counter: one per (device, stream) -- calls on one stream are ordered and may share
    it, calls on different streams may run at the same time and may not."""
    key = (device.index if device.type == "cuda" else -1, stream)
    w = _WORK.get(key)
    if w is None:
        n = int(_capi.lib().dpc_proj_l2_loss_workspace_bytes())
        w = torch.zeros((n + 3) // 4, dtype=torch.int32, device=device)      # zeroed once; the kernel leaves it reusable
        _WORK[key] = w
    return w


class _ProjL2LossFn(torch.autograd.Function):
    @staticmethod
    @_capi.on_tensor_device
    def forward(ctx, pred, gt, num_samples):
        L = _capi.lib()
        p, g = _capi.f32c(pred), _capi.f32c(gt)
        if p.shape != g.shape:
            raise ValueError("proj_l2_loss: prediction %s and ground truth %s differ in shape" % (tuple(p.shape), tuple(g.shape)))
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        g_pred = torch.empty_like(p) if need else None
        stream = _capi.stream_of(p)
        w = _workspace(p.device, stream)
        _capi.check(L.dpc_proj_l2_loss(_capi.ptr(p), _capi.ptr(g), p.numel(), 1.0 / float(num_samples), _capi.ptr(loss),
                                       _capi.ptr(g_pred), _capi.ptr(w), w.numel() * 4, stream))
        if need:
            ctx.save_for_backward(g_pred)
        return loss

    @staticmethod
    @_capi.on_tensor_device
    def backward(ctx, g_loss):
        (g_pred,) = ctx.saved_tensors
        d = g_pred * g_loss
        return (d if ctx.needs_input_grad[0] else None), ((-d) if ctx.needs_input_grad[1] else None), None


def proj_l2_loss(gt, pred, num_samples=None):
    """sum((gt - pred)**2) / 2 / num_samples (num_samples defaults to the batch size, pred.shape[0])."""
    if num_samples is None:
        num_samples = pred.shape[0]
    return _ProjL2LossFn.apply(pred, gt, num_samples)


# ------------------------------------------------------------------ the reference's loss terms (dpc/util/losses.py)
def resize_tf1(img, size, method="bilinear"):
    """tf.image.resize_images(img [B,H,W,C], [size, size]) of TF 1.x for an integer reduction factor (identity when the
    size already matches): source coordinate = i * in / out is an integer, so bilinear and nearest are the same
    exact sub-sampling.  (Bicubic -- cfg.bicubic_gt_downsampling -- has no implementation here.)"""
    h = img.shape[1]
    if h == size:
        return img
    if method == "bicubic":
        raise NotImplementedError("bicubic_gt_downsampling is not implemented (the reference's experiments use bilinear)")
    if h < size or h % size != 0:
        raise NotImplementedError("resize %d -> %d: only integer reduction factors are implemented" % (h, size))
    s = h // size
    return img[:, ::s, ::s, :]


def _filtered_gt(cfg, gt, sigma, enabled):
    """pc_gauss_filter_gt(_rgb) with the switch-off at sigma < 1 (losses.py:82-87, model_pc.py:399-405)."""
    if not enabled:
        return gt
    from .gauss_kernel import gauss_smoothen_image
    smoothed = gauss_smoothen_image(cfg, gt, sigma)
    if cfg.pc_gauss_filter_gt_switch_off:
        return gt if float(sigma) < 1.0 else smoothed
    return smoothed


def drc_loss(cfg, probs, gt_proj):
    """sum(probs * psi), psi = [1 - gt] * vox_size ++ [gt]  (losses.py:25-33); probs [Vz+1,B,V,V,1]."""
    g = gt_proj.unsqueeze(0)
    psi = torch.cat([(1 - g).expand(int(cfg.vox_size), -1, -1, -1, -1), g], 0)
    return (probs * psi).sum()


def add_drc_loss(cfg, inputs, outputs, weight_scale, add_summary=True):
    gt, pred = inputs["masks"], outputs["drc_probs"]
    gt = resize_tf1(gt, pred.shape[2])
    return drc_loss(cfg, pred, gt) / gt.shape[0] * weight_scale


def add_proj_rgb_loss(cfg, inputs, outputs, weight_scale, add_summary=True, sigma=None):
    gt, pred = inputs["images"], outputs["projs_rgb"]
    gt = resize_tf1(gt, pred.shape[1])
    gt = _filtered_gt(cfg, gt, sigma, cfg.pc_gauss_filter_gt_rgb)
    return ((gt - pred) ** 2).sum() / 2 / pred.shape[0] * weight_scale


def add_proj_depth_loss(cfg, inputs, outputs, weight_scale, sigma_rel, add_summary=True):
    gt, pred = inputs["depths"], outputs["projs_depth"]
    if cfg.max_depth != cfg.max_dataset_depth:
        far = gt == cfg.max_dataset_depth
        gt = (~far).to(gt.dtype) * gt + far.to(gt.dtype) * cfg.max_depth
    gt = resize_tf1(gt, pred.shape[1], "nearest")
    if cfg.pc_gauss_filter_gt:
        from .gauss_kernel import gauss_smoothen_image
        gt = gauss_smoothen_image(cfg, gt, sigma_rel)
    return ((gt - pred) ** 2).sum() / 2 / pred.shape[0] * weight_scale
