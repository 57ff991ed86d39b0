"""Generates tests/golden/camera.npz from the reference's own camera helpers (dpc/util/camera.py:16-60,
dpc/util/euler.py:139-175), run here where /root/reference exists:  python tests/golden/make_golden_camera.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import run_reference  # noqa: E402

ns = run_reference.load()
cam = ns.camera
importlib.import_module("util.euler")
rng = np.random.RandomState(5)
extr = rng.randn(6, 4, 4).astype(np.float32)
extr[:, 3, :] = [0, 0, 0, 1]
pos = rng.randn(9, 3).astype(np.float32) * 2
pos[0] = [1.0, 0.0, 0.5]           # ty == 0 branch
pos[1] = [0.3, -0.7, -1.2]
ours = np.stack([cam.camera_from_blender(e) for e in extr])
quat = np.stack([cam.quaternion_from_campos(p) for p in pos]).astype(np.float32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "camera", "camera.npz"), extr=extr, pos=pos, ours=ours, quat=quat)
print("camera.npz written", ours.shape, quat.shape)
