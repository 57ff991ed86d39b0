"""Property tests (hypothesis) of the path, SURVEY section 4 item 1: mass conservation of the splat, range of the
silhouette, equivariance of the projection under a flip of the image-plane axes, linearity of the smoothing.  They run
on the oracle (CPU) and, on the same drawn inputs, on the real kernel sources under the CPU emulation (tests/emu)."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import dpc_oracle as O
from dpc_b200.util.config import default_config
from tests.emu_support import emu  # noqa: F401

CFG = default_config(vox_size=16, pc_gauss_kernel_size=5)
SET = dict(max_examples=20, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


def _cloud(seed, n, spread):
    g = torch.Generator().manual_seed(seed)
    return torch.tanh(spread * torch.randn(2, n, 3, generator=g)) * 0.6      # some points leave the cube


def _impls(emu):
    import dpc_b200.util.gauss_kernel as gk
    import dpc_b200.util.point_cloud as pcm
    return (("oracle", O, O), ("kernels", pcm, gk))


@settings(**SET)
@given(seed=st.integers(0, 2 ** 20), n=st.integers(1, 70), spread=st.sampled_from([0.02, 0.5, 2.0]))
def test_splat_conserves_mass(emu, seed, n, spread):
    pc = _cloud(seed, n, spread)
    valid = ((pc >= -0.5) & (pc <= 0.5)).all(-1)
    for name, mp, _ in _impls(emu):
        vox, _ = mp.pointcloud2voxels3d_fast(CFG, pc, None)
        got = vox.sum(dim=(1, 2, 3))
        # every valid point's 8 weights sum to 1; a coordinate of exactly +0.5 loses only zero-weight corners
        assert torch.allclose(got, valid.sum(1).float(), atol=1e-4), (name, got, valid.sum(1))
        assert float(vox.min()) >= 0.0


@settings(**SET)
@given(seed=st.integers(0, 2 ** 20), n=st.integers(1, 60), sigma=st.sampled_from([0.3, 1.0, 2.5]))
def test_silhouette_range_and_flip_equivariance(emu, seed, n, sigma):
    """Mirroring the cloud along an image-plane axis of camera space mirrors the silhouette (identity pose, no
    perspective skew of that symmetry: x -> -x commutes with the projection)."""
    pc = _cloud(seed, n, 0.5)
    g = torch.Generator().manual_seed(seed + 1)
    q = torch.tensor([[1.0, 0.0, 0.0, 0.0]]).repeat(2, 1)
    sc = torch.sigmoid(torch.randn(2, 1, generator=g))
    pc_m = pc.clone()
    pc_m[..., 2] = -pc_m[..., 2]              # input channel 2 becomes the image x axis (point_cloud.py:181-183)
    for name, mp, mg in _impls(emu):
        kern = mg.smoothing_kernel(CFG, torch.tensor(sigma))
        a = mp.pointcloud_project_fast(CFG, pc, q, None, None, kern, sc)["proj"]
        b = mp.pointcloud_project_fast(CFG, pc_m, q, None, None, kern, sc)["proj"]
        assert float(a.min()) >= 0.0 and float(a.max()) <= 1.0 + 1e-6, name
        # voxel centres are symmetric about 0 ((i + 0.5)/V - 0.5 is not, but g = (p + 0.5)(V - 1) is): flip along W
        assert torch.allclose(a.flip(2), b, atol=2e-5), (name, float((a.flip(2) - b).abs().max()))


@settings(**SET)
@given(seed=st.integers(0, 2 ** 20), alpha=st.floats(0.1, 3.0))
def test_smoothing_is_linear_and_preserves_interior_mass(emu, seed, alpha):
    g = torch.Generator().manual_seed(seed)
    v = CFG.vox_size
    x = torch.rand(1, v, v, v, 1, generator=g)
    y = torch.rand(1, v, v, v, 1, generator=g)
    for name, mp, mg in _impls(emu):
        kern = mg.smoothing_kernel(CFG, torch.tensor(1.0))
        sx, sy = mp.smoothen_voxels3d(CFG, x, kern), mp.smoothen_voxels3d(CFG, y, kern)
        sxy = mp.smoothen_voxels3d(CFG, x + alpha * y, kern)
        assert torch.allclose(sxy, sx + alpha * sy, atol=1e-5), name
        # a unit impulse well inside the grid keeps its mass (the taps sum to one along every axis)
        imp = torch.zeros(1, v, v, v, 1)
        imp[0, v // 2, v // 2, v // 2, 0] = 1.0
        assert abs(float(mp.smoothen_voxels3d(CFG, imp, kern).sum()) - 1.0) <= 1e-5, name
