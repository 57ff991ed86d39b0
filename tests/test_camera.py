"""f-3: the camera conversion of the input pipeline (util/camera.py) against values produced by the reference's own
helpers (tests/golden/make_golden_camera.py)."""
import os

import numpy as np

from dpc_b200.util import camera

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "camera", "camera.npz"))


def test_camera_from_blender_batched_and_single():
    assert np.array_equal(camera.camera_from_blender(Z["extr"]), Z["ours"])
    assert np.array_equal(camera.camera_from_blender(Z["extr"][2]), Z["ours"][2])


def test_quaternion_from_campos():
    q = camera.quaternion_from_campos(Z["pos"])
    assert q.dtype == np.float32 and q.shape == Z["quat"].shape
    assert np.abs(q - Z["quat"]).max() <= 1e-6
    assert np.abs(np.linalg.norm(q, axis=-1) - 1).max() <= 1e-6
    m, qq = camera.preprocess_cameras(Z["extr"], Z["pos"][:6])
    assert m.shape == (6, 4, 4) and qq.shape == (6, 4)
