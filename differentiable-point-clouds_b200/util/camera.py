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


# ---- f-3: camera conversion of the input pipeline (reference: dpc/util/camera.py:16-60, dpc/util/euler.py:139-175,
# used by models/model_base.py:74-97 on every batch through tf.py_func).  Host-side numpy like the reference (a few
# dozen floats per view); vectorised over the batch instead of a Python loop per matrix.
def camera_from_blender(their):
    """Blender extrinsic [..., 4, 4] -> this code base's camera convention (camera.py:16-38): rows / columns permuted
    and sign-flipped, row 3 = (0, 0, 0, their[3, 3])."""
    their = np.asarray(their)
    our = np.zeros(their.shape[:-2] + (4, 4), dtype=np.float32)
    our[..., 0, 0] = -their[..., 2, 0]
    our[..., 0, 1] = their[..., 2, 2]
    our[..., 0, 2] = their[..., 2, 1]
    our[..., 1, 0] = their[..., 1, 0]
    our[..., 1, 1] = -their[..., 1, 2]
    our[..., 1, 2] = -their[..., 1, 1]
    our[..., 2, 0] = -their[..., 0, 0]
    our[..., 2, 1] = their[..., 0, 2]
    our[..., 2, 2] = their[..., 0, 1]
    our[..., 0, 3] = their[..., 2, 3]
    our[..., 1, 3] = their[..., 1, 3]
    our[..., 2, 3] = their[..., 0, 3]
    our[..., 3, 3] = their[..., 3, 3]
    return our


def ypr_from_campos(cx, cy, cz):
    """(yaw, pitch, roll) of a camera at (cx, cy, cz) looking at the origin (euler.py:139-154); arrays broadcast."""
    cx, cy, cz = (np.asarray(v, dtype=np.float64) for v in (cx, cy, cz))
    dist = np.sqrt(cx * cx + cy * cy + cz * cz)
    cx, cy, cz = cx / dist, cy / dist, cz / dist
    t = np.sqrt(cx * cx + cy * cy)
    tx, ty = cx / t, cy / t
    yaw = np.arccos(tx)
    yaw = np.where(ty > 0, 2 * np.pi - yaw, yaw)
    return yaw, np.arcsin(cz), np.zeros_like(yaw)


def _axis_angle_quaternion(angle, axis):
    q = np.zeros(np.shape(angle) + (4,), dtype=np.float64)
    q[..., 0] = np.cos(angle / 2)
    q[..., 1:4] = np.sin(angle / 2)[..., None] * np.asarray(axis, dtype=np.float64)
    return q


def _q_mul(a, b):
    w1, x1, y1, z1 = (a[..., i] for i in range(4))
    w2, x2, y2, z2 = (b[..., i] for i in range(4))
    return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2, w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2], axis=-1)


def quaternion_from_campos(cam_pos):
    """Camera position(s) [..., 3] -> quaternion(s) [..., 4] (w first), float32: yaw (+pi, the Blender convention of
    camera.py:49-55) about y, pitch about z, roll = 0 about x, composed roll * (pitch * yaw) (euler.py:170-175)."""
    cam_pos = np.asarray(cam_pos)
    yaw, pitch, roll = ypr_from_campos(cam_pos[..., 0], cam_pos[..., 1], cam_pos[..., 2])
    yaw = yaw + np.pi
    q = _q_mul(_axis_angle_quaternion(roll, [1, 0, 0]),
               _q_mul(_axis_angle_quaternion(pitch, [0, 0, 1]), _axis_angle_quaternion(yaw, [0, 1, 0])))
    return q.astype(np.float32)


def preprocess_cameras(extrinsic, cam_pos):
    """The camera half of ModelBase.preprocess (model_base.py:74-97): batches of Blender extrinsics [N,4,4] and camera
    positions [N,3] -> (matrices [N,4,4], camera_quaternion [N,4]), the two pose inputs of the projection path."""
    return camera_from_blender(extrinsic), quaternion_from_campos(cam_pos)
