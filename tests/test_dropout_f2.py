"""f-2: point dropout consumed by the splat's load stage (reference point_cloud.py:293-319).
CPU: the real kernel sources under the CUDA-model emulator (tests/emu); GPU: the same assertions on the device at the
experiment's shape (N = 8000 -> 560 kept)."""
import numpy as np
import pytest
import torch

import dpc_b200.util.gauss_kernel as gk
import dpc_b200.util.point_cloud as pcm
from dpc_b200.util.config import default_config
from oracle import dpc_oracle as O
from tests.emu_support import emu  # noqa: F401


def _check_indices(sel, n, k):
    s = sel.cpu().numpy()
    assert s.shape[1] == k and s.dtype == np.int32
    assert s.min() >= 0 and s.max() < n
    for row in s:
        assert len(set(row.tolist())) == k, "indices of a sample must be distinct"


def _index_properties(device):
    n, k, b = 8000, 560, 6
    a = pcm.dropout_indices(b, n, k, device, seed=7, draw=0)
    _check_indices(a, n, k)
    assert torch.equal(a, pcm.dropout_indices(b, n, k, device, seed=7, draw=0)), "same (seed, draw) -> same subsets"
    assert not torch.equal(a, pcm.dropout_indices(b, n, k, device, seed=7, draw=1))
    assert not torch.equal(a, pcm.dropout_indices(b, n, k, device, seed=8, draw=0))
    assert not torch.equal(a[0], a[1]), "samples draw different subsets"
    # every index is kept with probability k / n: 40 draws x 6 samples x 560 picks over 8000 indices
    cnt = np.zeros(n, dtype=np.int64)
    for d in range(40):
        s = pcm.dropout_indices(b, n, k, device, seed=11, draw=d).cpu().numpy()
        np.add.at(cnt, s.reshape(-1), 1)
    mean = 40 * b * k / n                       # 16.8 expected hits per index
    assert abs(cnt.mean() - mean) < 1e-9
    assert cnt.max() < mean + 6 * np.sqrt(mean) and cnt.min() > max(0.0, mean - 6 * np.sqrt(mean))
    assert 0.8 * mean < cnt.var() < 1.2 * mean  # binomial-like spread, no structure
    # odd sizes, n_keep == n (a permutation), tiny clouds
    for nn, kk in ((1000, 1000), (37, 5), (2, 1), (5000, 4999)):
        _check_indices(pcm.dropout_indices(3, nn, kk, device, seed=3, draw=2), nn, kk)


def _fused_matches_oracle(device, b, n, k, v, ksize, atol=1e-5):
    cfg = default_config(vox_size=v, pc_gauss_kernel_size=ksize)
    g = torch.Generator().manual_seed(21)
    pc = torch.tanh(0.5 * torch.randn(b, n, 3, generator=g)) / 2
    q = torch.randn(b, 4, generator=g)
    sc = torch.sigmoid(torch.randn(b, 1, generator=g))
    gt = (torch.rand(b, v, v, 1, generator=g) > 0.5).float()
    sel = pcm.dropout_indices(b, n, k, device, seed=5, draw=9)
    # oracle: gather first (the reference's order of operations), then project
    op, oq, osc = [t.clone().requires_grad_(True) for t in (pc, q, sc)]
    sub, _ = O.pc_point_dropout_with_indices(op, None, sel.cpu().long())
    ref = O.pointcloud_project_fast(cfg, sub, oq, None, None, O.smoothing_kernel(cfg, torch.tensor(1.5)), osc)
    (((gt - ref["proj"]) ** 2).sum() / 2 / b).backward()
    # product: the index list goes to the splat
    cp, cq, csc = [t.clone().to(device).requires_grad_(True) for t in (pc, q, sc)]
    out = pcm.pointcloud_project_fast(cfg, cp, cq, None, None, gk.smoothing_kernel(cfg, 1.5), csc, point_indices=sel)
    (((gt.to(device) - out["proj"]) ** 2).sum() / 2 / b).backward()
    assert out["tr_pc"].shape == (b, k, 3)
    assert torch.equal(out["tr_pc"].detach().cpu(), ref["tr_pc"].detach()), "tr_pc of the kept points must be bit-exact"
    assert float((out["proj"].detach().cpu() - ref["proj"].detach()).abs().max()) <= atol
    assert float((out["voxels"].detach().cpu() - ref["voxels"].detach()).abs().max()) <= atol
    d_pc = cp.grad.cpu()
    assert d_pc.shape == pc.shape
    assert float((d_pc - op.grad).abs().max()) <= atol
    kept = torch.zeros(b, n, dtype=torch.bool)
    kept.scatter_(1, sel.cpu().long(), True)
    assert torch.all(d_pc[~kept] == 0), "dropped points get exactly zero gradient"
    for a, r in ((cq.grad.cpu(), oq.grad), (csc.grad.cpu(), osc.grad)):
        assert float((a - r).abs().max()) <= atol * max(1.0, float(r.abs().max()))
    # the materialising form gives the same points
    mp, _ = pcm.pc_point_dropout(pc.to(device), None, None, selected_indices=sel)
    assert torch.equal(mp.cpu(), sub.detach())


def test_dropout_indices_emulated(emu):  # noqa: F811
    _index_properties("cpu")


def test_fused_dropout_matches_oracle_emulated(emu):  # noqa: F811
    _fused_matches_oracle("cpu", b=2, n=600, k=42, v=16, ksize=5)


@pytest.mark.parametrize("knobs", [{19: 1}, {0: 8, 20: 6}])
def test_fused_dropout_kernel_variants_emulated(emu, knobs):  # noqa: F811
    """The index list in the software-pipelined splat kernels with several (ragged) tiles per warp (knob 19 = 1: grids sized
    for one warp per SM of the emulator's four), and in the tile-per-CTA kernels (knobs 0 = 8, 20 = 6)."""
    for k, v in knobs.items():
        emu.dpc_debug_set(k, v)
    try:
        _fused_matches_oracle("cpu", b=2, n=600, k=333, v=16, ksize=5)
    finally:
        for k in knobs:
            emu.dpc_debug_set(k, {0: 4, 19: 0, 20: 0}[k])


@pytest.mark.gpu
def test_dropout_indices_gpu():
    _index_properties("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 8000, 560, 64, 21), (2, 8000, 7999, 64, 21), (3, 1000, 70, 32, 11)])
def test_fused_dropout_matches_oracle_gpu(shape):
    _fused_matches_oracle("cuda:0", *shape)


@pytest.mark.gpu
def test_graph_replay_draws_fresh_subsets():
    """state on the device: a captured graph reads {seed, draw} at replay time."""
    dev = torch.device("cuda:0")
    state = torch.tensor([123, 0], dtype=torch.int64, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        sel = pcm.dropout_indices(2, 8000, 560, dev, state=state)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            sel = pcm.dropout_indices(2, 8000, 560, dev, state=state)
            state[1] += 1
    g.replay(); torch.cuda.synchronize(); a = sel.clone()
    g.replay(); torch.cuda.synchronize(); b = sel.clone()
    assert not torch.equal(a, b)
    assert torch.equal(a, pcm.dropout_indices(2, 8000, 560, dev, seed=123, draw=0))
    assert torch.equal(b, pcm.dropout_indices(2, 8000, 560, dev, seed=123, draw=1))
