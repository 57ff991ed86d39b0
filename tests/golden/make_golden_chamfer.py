"""Generates tests/golden/chamfer/*.npz by running the REFERENCE'S OWN point_cloud_distance
(/root/reference/dpc/util/point_cloud_distance.py, unmodified) over the TF1 shim, in fp32 and in fp64 (the evaluation's
precision, run/eval_chamfer.py:50-51).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden_chamfer.py

Each fixture: source / target sets (with exact duplicates and an exact hit, so that "first minimum" matters) and the
reference's (proj, minDist, idx)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import run_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "chamfer")


def sets(ns, nt, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    vs = (torch.tanh(0.5 * torch.randn(ns, 3, generator=g, dtype=torch.float64)) / 2).to(dtype)
    vt = (torch.tanh(0.5 * torch.randn(nt, 3, generator=g, dtype=torch.float64)) / 2).to(dtype)
    if nt >= 8:
        vt[nt // 2] = vt[3]
        vt[nt - 1] = vt[3]
        vs[0] = vt[3]
    return vs, vt


def main():
    ref = run_reference.load()
    tf = ref.tf
    os.makedirs(OUT, exist_ok=True)
    for name, ns, nt, dtype in (("f32_400x1500", 400, 1500, torch.float32), ("f64_400x1500", 400, 1500, torch.float64),
                                ("f64_37x5", 37, 5, torch.float64), ("f32_1x1", 1, 1, torch.float32)):
        vs, vt = sets(ns, nt, dtype, seed=ns * 7 + nt)
        proj, md, idx = ref.point_cloud_distance.point_cloud_distance(tf.constant(vs.numpy(), dtype=dtype),
                                                                      tf.constant(vt.numpy(), dtype=dtype))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), vs=vs.numpy(), vt=vt.numpy(), proj=np.asarray(proj.numpy()),
                            min_dist=np.asarray(md.numpy()), idx=np.asarray(idx.numpy()).astype(np.int32))
        print("wrote", name)


if __name__ == "__main__":
    main()
