"""CPU: the product's host code + the real kernel sources, executed by the CUDA-model emulator
(tests/emu), against the golden fixtures generated from the reference's own source.  Same
assertions as the GPU parity tests (tests/test_gpu_parity.py), on the small cases."""
import numpy as np
import pytest
import torch

import dpc_b200.util.gauss_kernel as gk
import dpc_b200.util.point_cloud as pcm
from tests import cases
from tests.emu_support import emu  # noqa: F401


class Product:
    smoothing_kernel = staticmethod(gk.smoothing_kernel)
    pointcloud_project_fast = staticmethod(pcm.pointcloud_project_fast)


SMALL = cases.golden_names()


@pytest.mark.parametrize("name", SMALL)
def test_emulated_kernels_match_golden(emu, name):  # noqa: F811
    fx = cases.load_golden(name)
    outs, grads = cases.run_impl(Product, fx)
    cases.assert_parity(fx, outs, grads)


@pytest.mark.parametrize("ppt", [1, 2])
def test_splat_points_per_thread_variants(emu, ppt):  # noqa: F811
    emu.dpc_debug_set(0, ppt)
    emu.dpc_debug_set(1, ppt)
    try:
        for name in ("cfg1_drc_k11", "clustered_init", "trans_focal"):
            fx = cases.load_golden(name)
            outs, grads = cases.run_impl(Product, fx)
            cases.assert_parity(fx, outs, grads)
    finally:
        emu.dpc_debug_set(0, 4)
        emu.dpc_debug_set(1, 4)


def test_conv_xy_128_thread_variant(emu):  # noqa: F811
    emu.dpc_debug_set(2, 128)
    try:
        for name in ("v64_small", "v64_k11_max"):
            fx = cases.load_golden(name)
            outs, grads = cases.run_impl(Product, fx)
            cases.assert_parity(fx, outs, grads)
    finally:
        emu.dpc_debug_set(2, 256)


@pytest.mark.parametrize("v,k", [(128, 11), (32, 21)])
def test_other_grid_sizes_against_oracle(emu, v, k):  # noqa: F811
    """The 32^3 and 128^3 instantiations of the shape-specialised kernels (BASELINE config 5 shapes)."""
    import dpc_b200.util.gauss_kernel as gkm
    from dpc_b200.util.config import default_config
    from oracle import dpc_oracle as O
    cfg = default_config(vox_size=v, pc_gauss_kernel_size=k)
    g = torch.Generator().manual_seed(3)
    b = 1 if v == 128 else 4
    pc = torch.tanh(0.5 * torch.randn(b, 300, 3, generator=g)) / 2
    q = torch.randn(b, 4, generator=g)
    sc = torch.sigmoid(torch.randn(b, 1, generator=g))
    gt = (torch.rand(b, v, v, 1, generator=g) > 0.5).float()
    res = {}
    for name, mp, mg in (("emu", pcm, gkm), ("oracle", O, O)):
        a = [x.clone().requires_grad_(True) for x in (pc, q, sc)]
        out = mp.pointcloud_project_fast(cfg, a[0], a[1], None, None, mg.smoothing_kernel(cfg, torch.tensor(2.0)), a[2])
        (((gt - out["proj"]) ** 2).sum() / 2 / b).backward()
        res[name] = (out, [x.grad for x in a])
    (eo, eg), (oo, og) = res["emu"], res["oracle"]
    assert torch.equal(eo["tr_pc"], oo["tr_pc"])
    for key in ("proj", "voxels"):
        assert float((eo[key].detach() - oo[key].detach()).abs().max()) <= 1e-5
    for x, y in zip(eg, og):
        assert float((x - y).abs().max()) <= 1e-5 * max(1.0, float(y.abs().max()))


def test_persistent_prefetch_conv_xy(emu):  # noqa: F811
    """Persistent conv_xy variant (dpc_debug_set(5, 2)): with B=1 every CTA has one slice, so also run a
    batch of 8 (512 slices on 444 CTAs: some CTAs loop twice through the double-buffered prefetch)."""
    import dpc_b200.util.gauss_kernel as gkm
    from dpc_b200.util.config import default_config
    from oracle import dpc_oracle as O
    emu.dpc_debug_set(5, 2)
    try:
        fx = cases.load_golden("v64_small")
        outs, grads = cases.run_impl(Product, fx)
        cases.assert_parity(fx, outs, grads)
        cfg = default_config(vox_size=64, pc_gauss_kernel_size=11)
        g = torch.Generator().manual_seed(11)
        pc = torch.tanh(0.5 * torch.randn(8, 64, 3, generator=g)) / 2
        q = torch.randn(8, 4, generator=g)
        sc = torch.sigmoid(torch.randn(8, 1, generator=g))
        o1 = pcm.pointcloud_project_fast(cfg, pc, q, None, None, gkm.smoothing_kernel(cfg, torch.tensor(1.5)), sc)
        o2 = O.pointcloud_project_fast(cfg, pc, q, None, None, O.smoothing_kernel(cfg, torch.tensor(1.5)), sc)
        assert float((o1["voxels"] - o2["voxels"]).abs().max()) <= 1e-5
        assert float((o1["proj"] - o2["proj"]).abs().max()) <= 1e-5
    finally:
        emu.dpc_debug_set(5, 0)


def test_conv_z_cpasync_tile_load_variant(emu):  # noqa: F811
    emu.dpc_debug_set(6, 1)
    try:
        for name in ("v64_small", "cfg1_drc_k11"):
            fx = cases.load_golden(name)
            outs, grads = cases.run_impl(Product, fx)
            cases.assert_parity(fx, outs, grads)
    finally:
        emu.dpc_debug_set(6, 0)
