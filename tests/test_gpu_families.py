"""GPU tests of the three smoothing-kernel families behind the same C-ABI (dpc_debug_set key 8):
  0 = CUDA-core FFMA2 / generic kernels, 1 = single-tile tcgen05 kernels, 2 = persistent tcgen05 pipelines (default).
Every family must meet the same parity bar (fixtures generated from the reference's source, the oracle at the full
benchmark shape) and the families must agree with each other and with an fp64 correlation on the standalone
entry points, for odd and even tap counts, every projection mode and gradient-sized magnitudes.
"""
import pytest
import torch
import torch.nn.functional as F

import dpc_b200.util.gauss_kernel as gk
import dpc_b200.util.point_cloud as pcm
from dpc_b200 import _capi
from tests import cases
from dpc_b200.util.config import default_config
from tests.test_gpu_parity import Product, _bench_inputs, _compare, _run_both

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FAMILIES = [0, 1, 2]


@pytest.fixture
def family(request):
    L = _capi.lib()
    L.dpc_debug_set(8, request.param)
    yield request.param
    L.dpc_debug_set(8, 2)


def _taps(K, sigma):
    x = torch.arange(K, dtype=torch.float64) - (K - 1) // 2
    w = torch.exp(-x * x / (2 * sigma * sigma))
    return (w / w.sum()).float()


def _ref_axis(v, taps, axis, pl):
    K = taps.numel()
    v = v.double().movedim(axis, -1)
    sh = v.shape
    x = F.pad(v.reshape(-1, 1, sh[-1]), (pl, K - 1 - pl))
    y = F.conv1d(x, taps.double().view(1, 1, K).to(x.device))
    return y.reshape(sh).movedim(-1, axis)


def _conv_xy(v, taps, clip_in=0, mask_out=None, mask_in=None):
    L, P = _capi.lib(), _capi.ptr
    K = taps.numel()
    B, Z, Y, X = v.shape
    out = torch.full_like(v, float("nan"))
    t = taps.to(DEV)
    _capi.check(L.dpc_conv_xy(P(v), P(out), P(t), K, (K - 1) // 2, P(t), K, (K - 1) // 2, B, Z, X, clip_in,
                              P(mask_out), P(mask_in), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out


def _conv_z(v, taps, scale=None, mode=-1):
    L, P = _capi.lib(), _capi.ptr
    K = taps.numel()
    B, Z, Y, X = v.shape
    out = torch.full_like(v, float("nan"))
    proj = torch.full((B, Y, X), float("nan"), device=DEV)
    mask2 = torch.zeros(B * Y * X * 2, dtype=torch.int32, device=DEV)
    t = taps.to(DEV)
    _capi.check(L.dpc_conv_z_fwd(P(v), P(t), K, (K - 1) // 2, P(scale), mode, 1e-5, 2.0, 10.0, 0, B, Z, X, P(out),
                                 P(mask2) if scale is not None else None, P(proj) if mode != -1 else None, None, None,
                                 torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out, proj, mask2


@pytest.mark.parametrize("family", FAMILIES, indirect=True)
@pytest.mark.parametrize("name", ["v64_small", "v64_extra_upstream", "v64_k11_max", "cfg1_drc_k21_sigma3", "even_kernel"])
def test_golden_fixture_per_family(family, name):
    fx = cases.load_golden(name)
    outs, grads = cases.run_impl(Product, fx, device=DEV)
    cases.assert_parity(fx, outs, grads)


@pytest.mark.parametrize("family", FAMILIES, indirect=True)
def test_full_shape_against_oracle_per_family(family):
    cfg = default_config(vox_size=64, pc_gauss_kernel_size=21)
    pc, q, sc, gt = _bench_inputs(4, 8000, 64, 0.5)
    _compare(_run_both(cfg, pc, q, sc, gt, 3.0, host_sigma=True))


@pytest.mark.parametrize("K,sigma", [(21, 3.0), (11, 1.5), (21, 0.2), (8, 2.0), (63, 9.0), (1, 1.0)])
def test_standalone_passes_agree_with_fp64(K, sigma):
    """6 samples = 192 depth-pass tiles: more than one tile per SM for the persistent pipelines."""
    L = _capi.lib()
    torch.manual_seed(K)
    v = torch.rand(6, 64, 64, 64, device=DEV)
    t = _taps(K, sigma)
    pl = (K - 1) // 2
    want_z = _ref_axis(v, t, 1, pl)
    want_xy = _ref_axis(_ref_axis(v, t, 3, pl), t, 2, pl)
    try:
        for fam in FAMILIES:
            L.dpc_debug_set(8, fam)
            got_z, _, _ = _conv_z(v, t)
            got_xy = _conv_xy(v, t)
            assert float((got_z.double() - want_z).abs().max()) <= 2e-6, ("z", fam)
            assert float((got_xy.double() - want_xy).abs().max()) <= 2e-6, ("xy", fam)
    finally:
        L.dpc_debug_set(8, 2)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_projection_modes_agree_across_families(mode):
    L = _capi.lib()
    torch.manual_seed(mode)
    v = torch.rand(6, 64, 64, 64, device=DEV)
    sc = torch.tensor([0.7, 1.9, 1.0, 0.3, 2.5, 1.2], device=DEV)
    t = _taps(21, 3.0)
    res = {}
    try:
        for fam in FAMILIES:
            L.dpc_debug_set(8, fam)
            res[fam] = _conv_z(v, t, scale=sc, mode=mode)
    finally:
        L.dpc_debug_set(8, 2)
    for fam in (1, 2):
        assert float((res[fam][0] - res[0][0]).abs().max()) <= 2e-6
        assert float((res[fam][1] - res[0][1]).abs().max()) <= 5e-6
        # the clip-pass bit planes may differ only where scale*smoothed sits within rounding of the clip bounds
        diff = (res[fam][2] ^ res[0][2])
        assert int((diff != 0).sum()) <= 64


def test_clip_masks_and_in_place_backward_pass():
    """forward xy pass writes the clip-pass bit plane; the backward xy pass runs in place and applies it."""
    L = _capi.lib()
    torch.manual_seed(7)
    raw = torch.rand(4, 64, 64, 64, device=DEV) * 1.6 - 0.3          # values on both sides of [0, 1]
    t = _taps(21, 3.0)
    g0 = torch.randn_like(raw) * 50.0
    outs = {}
    try:
        for fam in FAMILIES:
            L.dpc_debug_set(8, fam)
            mask = torch.zeros(raw.numel() // 32, dtype=torch.int32, device=DEV)
            sm = _conv_xy(raw, t, clip_in=1, mask_out=mask)
            g = g0.clone()
            P = _capi.ptr
            tt = t.to(DEV)
            _capi.check(L.dpc_conv_xy(P(g), P(g), P(tt), 21, 10, P(tt), 21, 10, 4, 64, 64, 0, None, P(mask),
                                      torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            outs[fam] = (sm, mask, g)
    finally:
        L.dpc_debug_set(8, 2)
    bits = ((raw >= 0) & (raw <= 1)).reshape(-1, 32).to(torch.int64)
    want_mask = (bits << torch.arange(32, device=DEV)).sum(1)
    want = _ref_axis(_ref_axis(raw.clamp(0, 1), t, 3, 10), t, 2, 10)
    want_g = _ref_axis(_ref_axis(g0, t, 3, 10), t, 2, 10) * ((raw >= 0) & (raw <= 1))   # symmetric taps: transpose = same pass
    for fam in FAMILIES:
        sm, mask, g = outs[fam]
        assert torch.equal(mask.to(torch.int64) & 0xffffffff, want_mask), fam
        assert float((sm.double() - want).abs().max()) <= 2e-6, fam
        assert float((g.double() - want_g).abs().max()) <= 2e-6 * 50 * 4, fam


def test_large_magnitudes_keep_relative_accuracy():
    L = _capi.lib()
    torch.manual_seed(3)
    g = torch.randn(2, 64, 64, 64, device=DEV) * 1e3
    t = _taps(21, 3.0)
    want = _ref_axis(_ref_axis(g, t, 3, 10), t, 2, 10)
    try:
        for fam in FAMILIES:
            L.dpc_debug_set(8, fam)
            got = _conv_xy(g, t)
            assert float((got.double() - want).abs().max()) <= 2e-6 * 1e3, fam
    finally:
        L.dpc_debug_set(8, 2)


def test_pipelines_run_when_tiles_are_fewer_than_sms_and_odd():
    """B = 1 (32 depth-pass tiles, 32 xy tiles) and B = 5 (160 tiles: one extra round for 12 CTAs)."""
    cfg = default_config(vox_size=64, pc_gauss_kernel_size=21)
    for b in (1, 5):
        pc, q, sc, gt = _bench_inputs(b, 2000, 64, 0.5, seed=99 + b)
        _compare(_run_both(cfg, pc, q, sc, gt, 3.0))
