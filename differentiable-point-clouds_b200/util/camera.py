"""Camera helpers on the hot path (reference: dpc/util/camera.py:5-13)."""
import numpy as np


def intrinsic_matrix(cfg, dims=3, inverse=False):
    """diag(1, f, f[, 1]) -- the pin-hole intrinsics the matrix-pose branch multiplies into the
    extrinsic (point_cloud.py:193-198).  The splat kernel applies it itself (dpc_math.cuh,
    dpc_pose_load); this function exists for callers that build camera matrices."""
    val = float(cfg.focal_length)
    if inverse:
        val = 1.0 / val
    m = np.eye(dims, dtype=np.float32)
    m[1, 1] = val
    m[2, 2] = val
    return m
