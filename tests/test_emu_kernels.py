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


SMALL = [n for n in cases.golden_names() if n not in ("cfg1_drc_k21_sigma3",)]


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
