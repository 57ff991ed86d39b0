"""TEST INFRASTRUCTURE: loads the CPU emulation build of the kernels (tests/emu) and points the
product's ctypes layer at it for the duration of a test, so the real host code + real kernel
sources run on CPU tensors.  The product itself never does this."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
EMU_LIB = os.path.join(EMU_DIR, "libdpc_b200_emu.so")
CSRC = os.path.join(os.path.dirname(HERE), "differentiable-point-clouds_b200", "csrc")


def _stale():
    if not os.path.isfile(EMU_LIB):
        return True
    t = os.path.getmtime(EMU_LIB)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs += [os.path.join(EMU_DIR, f) for f in ("cuda_emu.h", "cuda_emu.cpp")]
    return any(os.path.getmtime(s) > t for s in srcs)


def build_emu():
    if _stale():
        subprocess.run(["sh", os.path.join(EMU_DIR, "build_emu.sh")], check=True, capture_output=True)
    return EMU_LIB


@pytest.fixture()
def emu(monkeypatch):
    from dpc_b200 import _capi
    lib = _capi.load_library(build_emu())
    assert lib.dpc_is_cuda_build() == 0
    monkeypatch.setattr(_capi, "_LIB", lib)
    monkeypatch.setattr(_capi, "_REQUIRE_CUDA", False)
    return lib
