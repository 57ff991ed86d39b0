"""Quaternion helpers of the hot path (reference: dpc/util/quaternion.py:62-117).

`quaternion_rotate` is the first op of the projection path and runs inside the splat kernel;
the small algebra helpers the model code uses around it are plain torch.
"""
import torch


def quaternion_multiply(a, b):
    """Hamilton product, last dimension 4 (w first); a 3-vector is treated as (0,x,y,z)."""
    if a.shape[-1] == 3:
        a = torch.nn.functional.pad(a, (1, 0))
    if b.shape[-1] == 3:
        b = torch.nn.functional.pad(b, (1, 0))
    if a.shape[-1] != 4 or b.shape[-1] != 4:
        raise ValueError("Can't create a quaternion from a tensor with shape %s. The last dimension must be 4."
                         % (tuple(a.shape),))
    w1, x1, y1, z1 = a.unbind(-1)
    w2, x2, y2, z2 = b.unbind(-1)
    return torch.stack((w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                        w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                        w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                        w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2), dim=-1)


def quaternion_conjugate(q):
    # (w, -x, -y, -z) without a host-built constant: a host -> device copy cannot be captured in a CUDA graph
    return torch.cat([q[..., :1], -q[..., 1:]], dim=-1)


def quaternion_normalise(q):
    return q / torch.linalg.vector_norm(q, dim=-1, keepdim=True)


def quaternion_rotate(pc, q, inverse=False):
    """q * pc * q' for pc [B,N,3], q [B,4] (normalised inside).  On the projection path this
    rotation is fused into the splat kernel (dpc_math.cuh, dpc_quat_rotate); this stand-alone
    form serves the model code that rotates clouds outside the renderer (model_pc.py:84)."""
    qn = quaternion_normalise(q).unsqueeze(1)
    qc = quaternion_conjugate(qn)
    if inverse:
        qn, qc = qc, qn
    r = quaternion_multiply(quaternion_multiply(qn, pc), qc)
    return r[..., 1:4]
