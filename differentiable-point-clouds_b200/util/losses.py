"""The silhouette loss of the training step as one kernel.

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
    def backward(ctx, g_loss):
        (g_pred,) = ctx.saved_tensors
        d = g_pred * g_loss
        return (d if ctx.needs_input_grad[0] else None), ((-d) if ctx.needs_input_grad[1] else None), None


def proj_l2_loss(gt, pred, num_samples=None):
    """sum((gt - pred)**2) / 2 / num_samples (num_samples defaults to the batch size, pred.shape[0])."""
    if num_samples is None:
        num_samples = pred.shape[0]
    return _ProjL2LossFn.apply(pred, gt, num_samples)
