"""CPU: hand-derived known answers and invariants of the oracle."""
import numpy as np
import pytest
import torch

from dpc_b200.util.config import default_config
from oracle import dpc_oracle as O

Q_ID = torch.tensor([[1.0, 0.0, 0.0, 0.0]])


def f32(x):
    return np.float32(x)


def test_transform_known_answers():
    cfg = default_config()
    pc = torch.tensor([[[0.0, 0, 0], [0, 0.2, 0], [-0.5, 0.45, 0], [0.1, 0.2, 0.3]]])
    tr = O.pc_perspective_transform(cfg, pc, Q_ID).numpy()[0]
    assert np.array_equal(tr[0], [0, 0, 0])
    assert np.array_equal(tr[1], np.array([0, 0.1875, 0], np.float32))
    assert np.array_equal(tr[2], np.array([-0.5, 0.5625, 0], np.float32))
    # (z+2)-2 rounding of 0.1 in fp32; x,y divided by z+2
    z = f32(f32(f32(0.1) + f32(2.0)) - f32(2.0))
    y = f32(f32(f32(0.2) * f32(1.875)) / f32(f32(0.1) + f32(2.0)))
    x = f32(f32(f32(0.3) * f32(1.875)) / f32(f32(0.1) + f32(2.0)))
    assert np.array_equal(tr[3], np.array([z, y, x], np.float32))
    assert abs(float(tr[3][0]) - 0.0999999) < 1e-7
    q = torch.tensor([[0.5 ** 0.5, 0.0, 0.5 ** 0.5, 0.0]])
    tr2 = O.pc_perspective_transform(cfg, pc[:, 3:], q).numpy()[0, 0]
    assert np.allclose(tr2, [0.29999995, 0.1630435, -0.08152173], atol=2e-7)


def test_origin_point_splats_eight_eighths():
    cfg = default_config()
    vox, _ = O.pointcloud2voxels3d_fast(cfg, torch.zeros(1, 1, 3), None)
    nz = torch.nonzero(vox)
    assert nz.shape[0] == 8
    assert set(nz[:, 1:].flatten().tolist()) == {31, 32}
    assert torch.all(vox[vox > 0] == 0.125)
    assert float(vox.sum()) == 1.0


def test_invalid_point_contributes_nothing_and_gets_zero_grad():
    cfg = default_config()
    pc = torch.tensor([[[-0.5, 0.45, 0.0], [0.0, 0.0, float("nan")]]], requires_grad=True)
    out = O.pointcloud_project_fast(cfg, pc, Q_ID, None, None, O.smoothing_kernel(cfg, 1.0), torch.tensor([[0.5]]))
    assert float(out["voxels"].abs().sum()) == 0.0
    out["proj"].sum().backward()
    g = pc.grad
    assert torch.all(g[0, 0] == 0)


def test_plus_half_is_dropped_not_an_error():
    cfg = default_config(vox_size=16)
    pc = torch.tensor([[[0.5, 0.0, 0.0]]])  # reaches the splat as exactly +0.5 in depth
    tr = O.pc_perspective_transform(cfg, pc, Q_ID)
    assert float(tr[0, 0, 0]) == 0.5
    vox, _ = O.pointcloud2voxels3d_fast(cfg, tr, None)
    assert abs(float(vox.sum()) - 1.0) < 1e-6
    assert float(vox[0, 15].sum()) == float(vox.sum())


def test_drc_known_answer():
    cfg = default_config()
    v = torch.tensor([0.0, 0.5, 1.0, 0.25]).reshape(1, 4, 1, 1, 1)
    proj, p = O.drc_projection(v, cfg)
    expect = [1.0000099e-05, 4.9999499e-01, 4.9998999e-01, 1.2516854e-06, 3.7550901e-06]
    assert np.allclose(p.flatten().numpy(), expect, rtol=2e-6, atol=0)
    assert abs(float(proj) - 0.99999624) < 1e-6


@pytest.mark.parametrize("l,lo,hi", [(21, -10, 10), (11, -5, 5), (10, -4, 5), (5, -2, 2)])
def test_tap_support(l, lo, hi):
    k = O.gauss_kernel_1d(l, 1.3)
    assert k.shape[0] == hi - lo + 1 == l
    xx = np.arange(lo, hi + 1, dtype=np.float64)
    ref = np.exp(-xx ** 2 / (2 * 1.3 ** 2))
    ref /= ref.sum()
    assert np.allclose(k.numpy(), ref, atol=1e-7)


def test_mass_conservation_and_range():
    cfg = default_config(vox_size=32, pc_gauss_kernel_size=11)
    g = torch.Generator().manual_seed(7)
    pc = torch.tanh(0.5 * torch.randn(3, 700, 3, generator=g)) / 2
    q = torch.randn(3, 4, generator=g)
    tr = O.pc_perspective_transform(cfg, pc, q)
    valid, _, _ = O.voxel_indices(cfg, tr)
    vox, _ = O.pointcloud2voxels3d_fast(cfg, tr, None)
    assert torch.allclose(vox.sum(dim=(1, 2, 3)), valid.sum(1).float(), atol=1e-3)
    out = O.pointcloud_project_fast(cfg, pc, q, None, None, O.smoothing_kernel(cfg, 1.0), torch.sigmoid(torch.randn(3, 1, generator=g)))
    assert float(out["proj"].min()) >= 0.0 and float(out["proj"].max()) <= 1.0
    assert out["drc_probs"].shape == (33, 3, 32, 32, 1)
    # termination probabilities of a ray sum to ~1 (up to the e^eps quirk)
    s = out["drc_probs"].sum(0)
    assert float((s - 1).abs().max()) < 1e-3


def test_fp64_finite_difference_gradients():
    cfg = default_config(vox_size=8, pc_gauss_kernel_size=5)
    g = torch.Generator().manual_seed(3)
    pc = (torch.tanh(0.5 * torch.randn(1, 20, 3, generator=g, dtype=torch.float64)) / 2).requires_grad_(True)
    q = torch.randn(1, 4, generator=g, dtype=torch.float64).requires_grad_(True)
    sc = torch.tensor([[0.6]], dtype=torch.float64, requires_grad=True)
    ker = [k.double() for k in O.smoothing_kernel(cfg, 0.9)]
    w = torch.randn(1, 8, 8, 1, generator=g, dtype=torch.float64)

    def fn(pc_, q_, sc_):
        return (O.pointcloud_project_fast(cfg, pc_, q_, None, None, ker, sc_)["proj"] * w).sum()

    assert torch.autograd.gradcheck(fn, (pc, q, sc), eps=1e-7, atol=1e-5, rtol=1e-4)
