"""Nearest-neighbour projection between point sets -- the chamfer evaluation's inner op.

Mirrors /root/reference/dpc/util/point_cloud_distance.py:26-39 (`point_cloud_distance(Vs, Vt)`), which
`util/point_cloud.py:7` re-exports and `run/eval_chamfer.py:11,55` calls on float64 placeholders.  Same call
signature and results `(proj [VsN,3], minDist [VsN], idx [VsN] int32)` on torch CUDA tensors (float32 or
float64, like the inputs), through dpc_point_cloud_distance_f32/_f64 of the C-ABI.  The reference materialises
[VsN, VtN, 3] and therefore feeds the source in `pc_eval_chamfer_num_parts` pieces (`eval_chamfer.py:18-34`);
the kernel streams the targets through shared memory, so `compute_distance` below takes the whole set at once
(the `num_parts` argument is accepted and ignored).  Evaluation only: the results carry no gradient.
"""
import torch

from .. import _capi


def point_cloud_distance(Vs, Vt):
    """For each point in Vs: the closest point in Vt, its distance and its index (first minimum)."""
    if Vs.dim() != 2 or Vt.dim() != 2 or Vs.shape[1] != 3 or Vt.shape[1] != 3:
        raise ValueError("point_cloud_distance expects [VsN,3] and [VtN,3], got %s and %s" % (tuple(Vs.shape), tuple(Vt.shape)))
    if Vs.dtype != Vt.dtype or Vs.dtype not in (torch.float32, torch.float64):
        raise ValueError("point_cloud_distance: both sets must be float32 or both float64")
    if Vs.shape[0] < 1 or Vt.shape[0] < 1:
        raise ValueError("point_cloud_distance: empty point set")   # tf.argmin over an empty axis raises too
    L = _capi.lib()
    vs, vt = Vs.detach().contiguous(), Vt.detach().contiguous()
    ns, nt = vs.shape[0], vt.shape[0]
    esz = vs.element_size()
    proj = torch.empty_like(vs)
    min_dist = torch.empty(ns, dtype=vs.dtype, device=vs.device)
    idx = torch.empty(ns, dtype=torch.int32, device=vs.device)
    wbytes = L.dpc_point_cloud_distance_workspace_bytes(ns, nt, esz)
    work = torch.empty((wbytes + 7) // 8, dtype=torch.int64, device=vs.device)
    fn = L.dpc_point_cloud_distance_f32 if esz == 4 else L.dpc_point_cloud_distance_f64
    _capi.check(fn(_capi.ptr(vs), ns, _capi.ptr(vt), nt, _capi.ptr(proj), _capi.ptr(min_dist), _capi.ptr(idx),
                   _capi.ptr(work), wbytes, _capi.stream_of(vs)))
    return proj, min_dist, idx


def compute_distance(source, target, num_parts=None):
    """`eval_chamfer.py:18-34` (compute projection from source to target): (min_dist, idx) of every source point.
    The reference cuts the source into `num_parts` pieces to bound its [VsN/num_parts, VtN, 3] temporaries; not needed."""
    _, min_dist, idx = point_cloud_distance(source, target)
    return min_dist, idx


def chamfer_pair(pred, gt):
    """The two directed means of `eval_chamfer.py:111-114`: (mean pred->gt, mean gt->pred) as Python floats."""
    pred_to_gt, _ = compute_distance(pred, gt)
    gt_to_pred, _ = compute_distance(gt, pred)
    return float(pred_to_gt.double().mean()), float(gt_to_pred.double().mean())
