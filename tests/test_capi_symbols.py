"""CPU: the C-ABI library builds for sm_100a, loads, and exports every function that
include/dpc_b200.h declares (no compute without a GPU); the host layer refuses CPU tensors."""
import os
import re

import pytest
import torch

from dpc_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "dpc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#ifdef DPC_EXPERIMENTS.*?#endif", "", text, flags=re.S)       # lab-only diagnostics
    return sorted(set(re.findall(r"\b(dpc_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_capi.LIB_PATH):
        _capi.build()
    return _capi.load_library(_capi.LIB_PATH)


def test_every_declared_symbol_is_exported(lib):
    names = header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_capi.EXPORTED_SYMBOLS) == names, "ctypes signatures out of sync with the header"


def test_is_the_cuda_build_with_sm100a_code(lib):
    assert lib.dpc_is_cuda_build() == 1


def test_product_build_has_no_experiment_knobs(lib):
    """The product library exports the ABI, not the lab: experiment keys are refused, lab-only symbols are absent."""
    assert lib.dpc_is_lab_build() == 0
    for key in (0, 1, 2, 4, 5, 10, 11, 13, 14, 15, 16, 18):
        assert lib.dpc_debug_set(key, 1) == -3
    for name in ("dpc_debug_mma_bench", "dpc_debug_phase_read", "dpc_debug_trace_read"):
        assert not hasattr(lib, name), name
    if os.path.isfile(_capi.LAB_LIB_PATH):
        lab = _capi.load_library(_capi.LAB_LIB_PATH)
        assert lab.dpc_is_lab_build() == 1 and hasattr(lab, "dpc_debug_mma_bench")
    assert lib.dpc_abi_version() == 3
    assert lib.dpc_error_string(-2).decode().startswith("unsupported shape")


def test_argument_validation_needs_no_gpu(lib):
    # NULL / shape errors are detected before anything touches the device
    assert lib.dpc_splat_fwd(None, None, 0, None, None, 1.0, 2.0, None, 1, 1, 8, 8, None, None, None, None, None, None) == -1
    p = _capi.ProjectParams(B=1, N=10, Vz=200, V=64, pose_kind=0, mode=0, K=21, Kz=21)
    import ctypes
    assert lib.dpc_project_fast_scratch_bytes(ctypes.byref(p)) == -1
    p.Vz = 64
    assert lib.dpc_project_fast_scratch_bytes(ctypes.byref(p)) == 2 * 64 ** 3 * 4 + 1024 + 256 + 256   # two grids + dL/dscale partials + counters + dL/dtr_pc of the 10 points (256-aligned)
    assert lib.dpc_project_fast_saved_bytes(ctypes.byref(p)) >= 2 * 64 ** 3 // 8


def test_product_refuses_cpu_tensors():
    from dpc_b200.util import point_cloud as pcm
    from dpc_b200.util.config import default_config
    cfg = default_config(vox_size=16)
    with pytest.raises(ValueError, match="CUDA"):
        pcm.pointcloud_project_fast(cfg, torch.zeros(1, 4, 3), torch.tensor([[1.0, 0, 0, 0]]), None, None)


def test_shape_errors_mirror_the_reference():
    from dpc_b200.util import point_cloud as pcm
    from dpc_b200.util.config import default_config
    cfg = default_config(vox_size=16)
    with pytest.raises(ValueError, match="quaternion"):
        pcm.pc_perspective_transform(cfg, torch.zeros(1, 4, 3), torch.zeros(1, 3))
    cfg2 = default_config(vox_size=16, pose_quaternion=False)
    with pytest.raises(ValueError):
        pcm.pc_perspective_transform(cfg2, torch.zeros(1, 4, 3), torch.zeros(1, 4, 4), torch.zeros(1, 3))
