#!/usr/bin/env python
"""One forward+backward step of the headline shape through the C-ABI (for compute-sanitizer / quick hang checks).
usage: one_step.py [B] [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bench.B = b
pipe = bench.Pipeline(torch.device("cuda:0"), 0)
for _ in range(reps):
    pipe.step()
torch.cuda.synchronize()
print("one_step ok: B=%d reps=%d |d_pc|max %.4g d_q[0] %s d_scale[0] %.6g" % (
    b, reps, float(pipe.d_pc.abs().max()), pipe.d_q[0].tolist(), float(pipe.d_sc[0])))
