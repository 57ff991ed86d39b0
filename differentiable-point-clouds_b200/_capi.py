"""ctypes binding of the C-ABI in include/dpc_b200.h (libdpc_b200.so, hand-written sm_100a CUDA).

There is no CPU fallback: if the shared library is missing or a tensor is not on a CUDA device
the call raises.  (`_LIB` / `_REQUIRE_CUDA` are module attributes only so that tests/emu can
drive this same host code over the CPU emulation build of the kernels; nothing in the package
ever sets them.)
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC_DIR, "libdpc_b200.so")
# the LAB build of the same sources (-DDPC_EXPERIMENTS): experiment knobs (dpc_debug_set) and the experimental kernels.
# Loaded only when DPC_LAB=1 is set or a test asks for it (lab_lib()); the package itself always uses the product build.
LAB_LIB_PATH = os.path.join(CSRC_DIR, "libdpc_b200_lab.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

_LIB = None
_REQUIRE_CUDA = True

c_f = ctypes.c_float
c_i = ctypes.c_int
c_p = ctypes.c_void_p
c_i64 = ctypes.c_int64


class ProjectParams(ctypes.Structure):
    _fields_ = [("B", c_i), ("N", c_i), ("Vz", c_i), ("V", c_i), ("pose_kind", c_i), ("mode", c_i),
                ("K", c_i), ("Kz", c_i), ("focal_const", c_f), ("cam_dist", c_f), ("clip_eps", c_f),
                ("max_depth", c_f), ("flags", c_i), ("taps_xy_host", c_p), ("taps_z_host", c_p), ("tr_pc", c_p), ("sel", c_p), ("N_src", c_i)]


FLAG_SCRATCH_RAW_ZERO = 1
ABI_VERSION = 3


POSE_NONE, POSE_QUAT, POSE_MATRIX = -1, 0, 1
PROJ_NONE, PROJ_DRC, PROJ_MAX, PROJ_DRC_PROD = -1, 0, 1, 2
MAX_TAPS = 63
MAX_V = 128

_PP = ctypes.POINTER(ProjectParams)
_SIGNATURES = {
    "dpc_abi_version": (c_i, []),
    "dpc_error_string": (ctypes.c_char_p, [c_i]),
    "dpc_last_cuda_error": (c_i, []),
    "dpc_is_cuda_build": (c_i, []),
    "dpc_debug_set": (c_i, [c_i, c_i]),
    "dpc_debug_stage_ms": (c_i, [c_p]),
    "dpc_debug_ktrace_read": (c_i, [c_p]),
    "dpc_is_lab_build": (c_i, []),
    "dpc_proj_l2_loss_workspace_bytes": (c_i64, []),
    "dpc_proj_l2_loss": (c_i, [c_p, c_p, c_i64, c_f, c_p, c_p, c_p, c_i64, c_p]),
    "dpc_point_cloud_distance_workspace_bytes": (c_i64, [c_i, c_i, c_i]),
    "dpc_point_cloud_distance_f32": (c_i, [c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_i64, c_p]),
    "dpc_point_cloud_distance_f64": (c_i, [c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_i64, c_p]),
    "dpc_splat_fwd": (c_i, [c_p, c_p, c_i, c_p, c_p, c_f, c_f, c_p, c_i, c_i, c_i, c_i,
                            c_p, c_p, c_p, c_p, c_p, c_p]),
    "dpc_splat_bwd": (c_i, [c_p, c_p, c_i, c_p, c_p, c_f, c_f, c_p, c_i, c_i, c_i, c_i, c_i,
                            c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "dpc_conv_xy": (c_i, [c_p, c_p, c_p, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "dpc_conv_z_fwd": (c_i, [c_p, c_p, c_i, c_i, c_p, c_i, c_f, c_f, c_f, c_i, c_i, c_i, c_i,
                             c_p, c_p, c_p, c_p, c_p, c_p]),
    "dpc_conv_z_bwd": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_f, c_i, c_i, c_i, c_i,
                             c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "dpc_project_fast_scratch_bytes": (c_i64, [_PP]),
    "dpc_project_fast_saved_bytes": (c_i64, [_PP]),
    "dpc_project_fast_fwd": (c_i, [_PP, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                                   c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_p, c_i64, c_p]),
    "dpc_project_fast_bwd": (c_i, [_PP, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                                   c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_p, c_i64, c_p]),
    "dpc_project_rgb_fwd": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    "dpc_project_rgb_bwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p]),
    "dpc_tap_corr": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "dpc_dropout_indices": (c_i, [ctypes.c_uint64, ctypes.c_uint64, c_p, c_i, c_i, c_i, c_p, c_p]),
    "dpc_gather_points": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    "dpc_gather_points_bwd": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)
# additional exports of the lab build only (diagnostics of the experiments)
_LAB_SIGNATURES = {
    "dpc_debug_trace_read": (c_i, [c_p]),
    "dpc_debug_phase_read": (c_i, [c_p]),
    "dpc_debug_mma_bench": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p]),
    "dpc_debug_gather_bench": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p]),
}


def build(verbose=False, lab=None):
    """Compile csrc/dpc_capi.cu for sm_100a with nvcc (cross-compiles without a GPU): the product library and (lab=True,
    or lab=None for both) the lab build with the experiment knobs."""
    src = os.path.join(CSRC_DIR, "dpc_capi.cu")
    out = None
    for is_lab in ((False, True) if lab is None else (bool(lab),)):
        path = LAB_LIB_PATH if is_lab else LIB_PATH
        cmd = ["nvcc"] + NVCC_FLAGS + (["-DDPC_EXPERIMENTS"] if is_lab else []) + ["-o", path, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (res.stdout, res.stderr))
        if not is_lab:
            out = res.stderr if verbose else path
    return out if out is not None else LAB_LIB_PATH


def _declare(lib):
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


def load_library(path):
    lib = _declare(ctypes.CDLL(path))
    if lib.dpc_is_lab_build():
        for name, (restype, argtypes) in _LAB_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
    return lib


_LAB = None


def lab_lib():
    """The lab build (experiment knobs).  Test / experiment infrastructure: nothing in the package calls this."""
    global _LAB
    if _LAB is None:
        if not os.path.isfile(LAB_LIB_PATH):
            raise RuntimeError("dpc_b200: %s is missing (python __graft_entry__.py builds it)" % LAB_LIB_PATH)
        _LAB = load_library(LAB_LIB_PATH)
    return _LAB


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "dpc_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        lab = os.environ.get("DPC_LAB") == "1" or bool(os.environ.get("DPC_KNOBS"))
        _LIB = lab_lib() if lab else load_library(LIB_PATH)
        if _LIB.dpc_abi_version() != ABI_VERSION:
            raise RuntimeError("dpc_b200: %s has ABI version %d, this package needs %d -- rebuild it (python __graft_entry__.py)"
                               % (LIB_PATH, _LIB.dpc_abi_version(), ABI_VERSION))
        if os.environ.get("DPC_TC"):      # override of the smoothing-kernel family (dpc_debug_set key 8)
            _LIB.dpc_debug_set(8, int(os.environ["DPC_TC"]))
        for kv in filter(None, os.environ.get("DPC_KNOBS", "").split(",")):     # experiments (lab build): "10=0,11=1"
            k, v = kv.split("=")
            check(_LIB.dpc_debug_set(int(k), int(v)))
    return _LIB


class DpcError(RuntimeError):
    pass


def check(code):
    if code != 0:
        L = lib()
        msg = L.dpc_error_string(code).decode()
        if code == -4:
            msg += " [cudaError %d]" % L.dpc_last_cuda_error()
        if code in (-2, -3):
            raise ValueError("dpc_b200: " + msg)
        raise DpcError("dpc_b200: " + msg)


def ptr(t):
    """Raw device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    if _REQUIRE_CUDA and not t.is_cuda:
        raise ValueError("dpc_b200 kernels need CUDA tensors (got a %s tensor); there is no CPU path" % t.device)
    if not t.is_contiguous():
        raise ValueError("dpc_b200: tensor must be contiguous")
    return t.data_ptr()


def stream_of(t):
    if t.is_cuda:
        return torch.cuda.current_stream(t.device).cuda_stream
    return None


def on_tensor_device(fn):
    """Decorator for autograd Function forward / backward: run the body with the device of the first CUDA tensor argument
    current (the C-ABI launches on the process's current device; tensors on cuda:1 while cuda:0 is current would otherwise
    launch on the wrong device).  No-op for CPU tensors (the emulation build of the test-suite)."""
    import functools

    @functools.wraps(fn)
    def wrapped(ctx, *args):
        t = next((a for a in args if torch.is_tensor(a) and a.is_cuda), None)
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(ctx, *args)
        with torch.cuda.device(t.device):
            return fn(ctx, *args)
    return wrapped


class device_of:
    """Context manager: make `t`'s device the current CUDA device for the C-ABI calls inside (kernel launches, memsets
    and function attributes act on the process's CURRENT device, whatever device the pointers live on)."""

    def __init__(self, t):
        self._ctx = torch.cuda.device(t.device) if (t is not None and t.is_cuda) else None

    def __enter__(self):
        if self._ctx is not None:
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            return self._ctx.__exit__(*exc)
        return False


def f32c(t):
    """contiguous float32 view/copy (None passes through)."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
