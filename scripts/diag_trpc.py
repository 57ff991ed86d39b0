"""Diagnostic: where does tr_pc differ from the oracle?"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpc_b200.util.point_cloud as pcm
from dpc_b200.util.config import default_config
from oracle import dpc_oracle as O

cfg = default_config(vox_size=32, pc_gauss_kernel_size=21)
for (b, n, seed) in ((8, 2000, 77), (4, 32000, 77), (32, 32768, 5), (8, 2000, 78), (8, 1999, 77), (8, 2048, 77)):
    g = torch.Generator().manual_seed(seed)
    pc = torch.tanh(0.5 * torch.randn(b, n, 3, generator=g)) / 2
    q = torch.randn(b, 4, generator=g)
    a = pcm.pc_perspective_transform(cfg, pc.cuda(), q.cuda()).cpu()
    r = O.pc_perspective_transform(cfg, pc, q)
    bad = (a != r) & ~(torch.isnan(a) & torch.isnan(r))
    idx = bad.nonzero()
    print("case", b, n, seed, "mismatches", idx.shape[0], "of", a.numel())
    for row in idx[:12].tolist():
        bi, ni, ci = row
        av, rv = a[bi, ni, ci].item(), r[bi, ni, ci].item()
        print("   b=%d n=%d c=%d cuda=%r (%s) oracle=%r (%s) pc=%s q=%s" % (
            bi, ni, ci, av, np.float32(av).view(np.uint32).item().__format__('08x'), rv,
            np.float32(rv).view(np.uint32).item().__format__('08x'), pc[bi, ni].tolist(), q[bi].tolist()))
    if idx.shape[0]:
        print("   per-sample mismatch counts", bad.sum(dim=(1, 2)).tolist())
        print("   point-index range of mismatches", int(idx[:, 1].min()), int(idx[:, 1].max()))
