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


@pytest.mark.parametrize("ppt", [1, 2, 4])
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
        emu.dpc_debug_set(1, 1)


@pytest.mark.parametrize("knobs", [{19: 1}, {19: 2}, {0: 8, 20: 6}])
def test_pipelined_splat_kernels(emu, knobs):  # noqa: F811
    """The software-pipelined splat kernels (product default; the tests above run them with one tile per warp): knob 19 = 1 / 2
    sizes their grids for 1 / 2 warps per SM of the emulator's 4 SMs, so every warp walks 4 - 16 tiles of its sample, the
    last one ragged; {0: 8, 20: 6} = the tile-per-CTA kernels they replaced (still the path for rgb / no tr_pc)."""
    for k, v in knobs.items():
        emu.dpc_debug_set(k, v)
    try:
        names = ("cfg1_drc_k11", "trans_focal", "edge_points", "matrix_pose", "extra_upstream", "single_point")
        for name in (names if knobs.get(19) != 2 else names[:2]):
            if name not in SMALL:
                continue
            fx = cases.load_golden(name)
            outs, grads = cases.run_impl(Product, fx)
            cases.assert_parity(fx, outs, grads)
    finally:
        for k in knobs:
            emu.dpc_debug_set(k, {0: 4, 19: 0, 20: 0}[k])


def _grads_before_and_after_corrupting_tr_pc(fx):
    """ONE forward (its atomics are unordered, two forwards would not be bit-comparable), two backwards: the second
    sees a tr_pc that is not the forward's."""
    cfg = cases.make_cfg(fx["cfg_over"])
    pc = torch.from_numpy(fx["in_point_cloud"]).clone().requires_grad_(True)
    q = torch.from_numpy(fx["in_transform"]).clone().requires_grad_(True)
    sc = torch.from_numpy(fx["in_scaling_factor"]).clone().requires_grad_(True)
    kernel = gk.smoothing_kernel(cfg, float(fx["in_sigma"]))
    out = pcm.pointcloud_project_fast(cfg, pc, q, None, None, kernel, sc)
    up = torch.randn(out["proj"].shape, generator=torch.Generator().manual_seed(3))
    good = [x.clone() for x in torch.autograd.grad(out["proj"], (pc, q, sc), grad_outputs=up, retain_graph=True)]
    t = out["tr_pc"].data                # .data: no version bump, autograd hands the corrupted tensor to the backward
    t.copy_(torch.roll(t, 7, dims=1))
    t[:, ::5] = float("nan")
    t[:, 1::5] *= -1.0
    bad = [x.clone() for x in torch.autograd.grad(out["proj"], (pc, q, sc), grad_outputs=up)]
    return good, bad


@pytest.mark.parametrize("name,per_sm", [("cfg1_drc_k11", 1), ("clustered_init", 0), ("v64_small", 1)])
def test_backward_does_not_depend_on_the_tr_pc_hint(emu, name, per_sm):  # noqa: F811
    """dpc_project_params.tr_pc only AIMS the prefetch of the pipelined splat backward: a lane whose recomputed cell is
    not the one its corners were requested for fetches them again.  A wrong tr_pc (another forward's; here rolled by 7
    points, every fifth NaN, every fifth negated) therefore gives the same gradients: d_pc bit for bit, the per-sample
    sums (pose: one atomic per warp, unordered) to rounding."""
    if name not in SMALL:
        pytest.skip("fixture not present")
    fx = cases.load_golden(name)
    emu.dpc_debug_set(19, per_sm)
    try:
        good, bad = _grads_before_and_after_corrupting_tr_pc(fx)
    finally:
        emu.dpc_debug_set(19, 0)
    assert torch.equal(good[0], bad[0]), "d_pc changed with the tr_pc hint"
    for g, b in zip(good[1:], bad[1:]):
        assert float((g - b).abs().max()) <= 1e-6 * max(1.0, float(g.abs().max()))


KNOB_DEFAULTS = {10: 0, 11: 1, 13: 0, 14: 1, 15: 1}


@pytest.mark.parametrize("knob,value", [(11, 0), (10, 1), (15, 0)])
def test_splat_reduction_and_zeroing_variants(emu, knob, value):  # noqa: F811
    """Knob 11 = 0: 8-byte / scalar reductions and gathers instead of 16-byte ones; knob 10 = 1: zeroing kernel +
    transform ahead of the grid dependency; knob 15 = 0: x/y pass out of place, backward in the second grid."""
    default = KNOB_DEFAULTS[knob]
    emu.dpc_debug_set(knob, value)
    try:
        for name in ("cfg1_drc_k11", "clustered_init", "v64_small", "edge_points"):
            if name not in SMALL:
                continue
            fx = cases.load_golden(name)
            outs, grads = cases.run_impl(Product, fx)
            cases.assert_parity(fx, outs, grads)
    finally:
        emu.dpc_debug_set(knob, default)


@pytest.mark.parametrize("name", ["v64_small", "v64_k11_max", "cfg1_drc_k21_sigma3", "cfg1_drc_k11", "vox_z"])
def test_device_taps_kernels(emu, name):  # noqa: F811
    """On the CPU the taps are host tensors, so the tests above run the launch-parameter (uniform-register)
    smoothing kernels; knob 5 makes the library ignore the host copy and run the vector-register ones."""
    emu.dpc_debug_set(5, 1)
    try:
        fx = cases.load_golden(name)
        outs, grads = cases.run_impl(Product, fx)
        cases.assert_parity(fx, outs, grads)
    finally:
        emu.dpc_debug_set(5, 0)


@pytest.mark.parametrize("device_taps", [0, 1])
@pytest.mark.parametrize("v,k", [(128, 11), (32, 21)])
def test_other_grid_sizes_against_oracle(emu, v, k, device_taps):  # noqa: F811
    """The 32^3 and 128^3 instantiations of the shape-specialised kernels (BASELINE config 5 shapes)."""
    emu.dpc_debug_set(5, device_taps)
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
    emu.dpc_debug_set(5, 0)
    (eo, eg), (oo, og) = res["emu"], res["oracle"]
    assert torch.equal(eo["tr_pc"], oo["tr_pc"])
    for key in ("proj", "voxels"):
        assert float((eo[key].detach() - oo[key].detach()).abs().max()) <= 1e-5
    for x, y in zip(eg, og):
        assert float((x - y).abs().max()) <= 1e-5 * max(1.0, float(y.abs().max()))


def test_conv_z_cpasync_tile_load_variant(emu):  # noqa: F811
    emu.dpc_debug_set(6, 1)
    try:
        for name in ("v64_small", "cfg1_drc_k11"):
            fx = cases.load_golden(name)
            outs, grads = cases.run_impl(Product, fx)
            cases.assert_parity(fx, outs, grads)
    finally:
        emu.dpc_debug_set(6, 0)
