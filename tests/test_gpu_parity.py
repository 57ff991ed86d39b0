"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the
reference-shaped Python API and directly through the C-ABI, against
  * the golden fixtures generated from the reference's own source (tests/golden),
  * the oracle on the same seeded inputs, at config 1 and at the full benchmark shape,
  * size-independent properties at the full shape.
Tolerances are the north star's: voxel indices and tr_pc bit-exact, grids / silhouettes /
gradients within 1e-5 abs.
"""
import ctypes

import numpy as np
import pytest
import torch

import dpc_b200.util.drc as drc_mod
import dpc_b200.util.gauss_kernel as gk
import dpc_b200.util.point_cloud as pcm
from dpc_b200 import _capi
from dpc_b200.util.config import default_config
from oracle import dpc_oracle as O
from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Product:
    smoothing_kernel = staticmethod(gk.smoothing_kernel)
    pointcloud_project_fast = staticmethod(pcm.pointcloud_project_fast)


def test_library_is_the_cuda_build():
    assert _capi.lib().dpc_is_cuda_build() == 1


@pytest.mark.parametrize("name", cases.golden_names())
def test_golden_fixture(name):
    fx = cases.load_golden(name)
    outs, grads = cases.run_impl(Product, fx, device=DEV)
    cases.assert_parity(fx, outs, grads)


@pytest.mark.parametrize("name", cases.golden_names())
def test_golden_fixture_host_sigma(name):
    """sigma as a host value: the taps also travel as launch parameters (uniform-register kernels)."""
    fx = cases.load_golden(name)
    outs, grads = cases.run_impl(Product, fx, device=DEV, host_sigma=True)
    cases.assert_parity(fx, outs, grads)


@pytest.mark.parametrize("name", ["cfg1_drc_k11", "v64_small", "edge_points", "vox_z", "clustered_init"])
def test_voxel_indices_bit_exact_through_c_abi(name):
    fx = cases.load_golden(name)
    cfg = cases.make_cfg(fx["cfg_over"])
    L = _capi.lib()
    pc = torch.from_numpy(fx["in_point_cloud"]).to(DEV)
    q = torch.from_numpy(fx["in_transform"]).to(DEV)
    b, n = pc.shape[:2]
    vz = cfg.vox_size_z if cfg.vox_size_z != -1 else cfg.vox_size
    idx = torch.empty(b, n, 3, dtype=torch.int32, device=DEV)
    valid = torch.empty(b, n, dtype=torch.uint8, device=DEV)
    tr = torch.empty_like(pc)
    vox = torch.zeros(b, vz, cfg.vox_size, cfg.vox_size, device=DEV)
    _capi.check(L.dpc_splat_fwd(pc.data_ptr(), q.data_ptr(), _capi.POSE_QUAT, None, None, float(cfg.focal_length),
                                float(cfg.camera_distance), None, b, n, vz, cfg.vox_size, tr.data_ptr(), vox.data_ptr(),
                                None, idx.data_ptr(), valid.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    v = fx["valid"]
    assert np.array_equal(valid.cpu().numpy().astype(bool), v)
    assert np.array_equal(idx.cpu().numpy()[v], fx["idx"][v])
    # mass conservation: every valid point deposits weights that sum to 1
    assert np.allclose(vox.sum(dim=(1, 2, 3)).cpu().numpy(), v.sum(1), atol=1e-3)


def _bench_inputs(b, n, v, spread, seed=1234):
    g = torch.Generator().manual_seed(seed)
    pc = torch.tanh(spread * torch.randn(b, n, 3, generator=g)) / 2
    q = torch.randn(b, 4, generator=torch.Generator().manual_seed(1236))
    sc = torch.sigmoid(torch.randn(b, 1, generator=torch.Generator().manual_seed(1237)))
    gt = (torch.rand(b, v, v, 1, generator=torch.Generator().manual_seed(1238)) > 0.5).float()
    return pc, q, sc, gt


def _run_both(cfg, pc, q, sc, gt, sigma, trans=None, host_sigma=False):
    """host_sigma: sigma reaches the product as a Python float -> taps also on the host -> the
    launch-parameter (uniform-register) smoothing kernels; otherwise the vector-register ones."""
    res = {}
    for name, mod_pc, mod_gk, dev in (("cuda", pcm, gk, DEV), ("oracle", O, O, "cpu")):
        leaves = [t.clone().to(dev).requires_grad_(True) for t in (pc, q, sc)]
        tr = trans.clone().to(dev).requires_grad_(True) if trans is not None else None
        ker = mod_gk.smoothing_kernel(cfg, sigma if (host_sigma and name == "cuda") else torch.tensor(sigma, device=dev))
        out = mod_pc.pointcloud_project_fast(cfg, leaves[0], leaves[1], tr, None, ker, leaves[2])
        loss = ((gt.to(dev) - out["proj"]) ** 2).sum() / 2 / pc.shape[0]
        loss.backward()
        res[name] = ({k: (None if v is None else v.detach().cpu()) for k, v in out.items()},
                     [t.grad.detach().cpu() for t in leaves] + ([tr.grad.detach().cpu()] if tr is not None else []))
    return res


def _compare(res, atol=1e-5):
    (co, cg), (oo, og) = res["cuda"], res["oracle"]
    assert torch.equal(co["tr_pc"], oo["tr_pc"]), "tr_pc must be bit-exact"
    for k in ("proj", "voxels", "drc_probs", "proj_depth"):
        if oo[k] is None:
            assert co[k] is None
            continue
        d = float((co[k] - oo[k]).abs().max())
        tol = atol * (max(1.0, float(oo[k].abs().max())) if k == "proj_depth" else 1.0)  # depth is O(10), see tests/cases.py
        assert d <= tol, "%s differs by %.3g" % (k, d)
    assert float((co["proj"] - oo["proj"]).abs().mean()) < 1e-5  # "silhouette L1 vs reference < 1e-5"
    for i, (a, b) in enumerate(zip(cg, og)):
        # relative to max(1, |g|_inf): these cases use small batches, and the loss is normalised by B, so gradients here are
        # up to 8x larger than at the headline shape (|d_pc| up to 25 at B=4, sigma=0.2: 2.3e-5 abs = 1e-6 relative).  The
        # ABSOLUTE 1e-5 gates (d_pc, d_scale; d_q with its fp32 bound) are applied at exactly the headline configuration in
        # tests/test_gpu_headline.py and printed by bench.py (`parity`).
        scale = max(1.0, float(b.abs().max()))
        d = float((a - b).abs().max())
        assert d <= atol * scale, "gradient %d differs by %.3g (scale %.3g)" % (i, d, scale)


@pytest.mark.parametrize("spread", [0.5, 0.025])
def test_config1_against_oracle(spread):
    """BASELINE config 1: B=2 N=1000 V=32 (numerics gate), spread and init-clustered clouds."""
    cfg = default_config(vox_size=32, pc_gauss_kernel_size=21)
    pc, q, sc, gt = _bench_inputs(2, 1000, 32, spread, seed=1234 if spread == 0.5 else 1235)
    _compare(_run_both(cfg, pc, q, sc, gt, 3.0))


@pytest.mark.parametrize("host_sigma", [False, True])
@pytest.mark.parametrize("sigma", [3.0, 0.2])
def test_full_benchmark_shape_against_oracle(sigma, host_sigma):
    """BASELINE config 2 shape (V=64, N=8000, K=21); B=4 keeps the CPU oracle to a few seconds."""
    cfg = default_config(vox_size=64, pc_gauss_kernel_size=21)
    pc, q, sc, gt = _bench_inputs(4, 8000, 64, 0.5)
    _compare(_run_both(cfg, pc, q, sc, gt, sigma, host_sigma=host_sigma))


def test_full_benchmark_shape_clustered_and_translation():
    cfg = default_config(vox_size=64, pc_gauss_kernel_size=21)
    pc, q, sc, gt = _bench_inputs(2, 8000, 64, 0.025, seed=1235)
    tr = 0.05 * torch.randn(2, 3, generator=torch.Generator().manual_seed(9))
    _compare(_run_both(cfg, pc, q, sc, gt, 3.0, trans=tr))
    _compare(_run_both(cfg, pc, q, sc, gt, 3.0, trans=tr, host_sigma=True))


@pytest.mark.parametrize("knob,value", [(2, 1), (4, 1), (10, 1), (10, 2), (10, 3), (11, 0), (13, 1), (14, 0), (15, 0),
                                        (0, 8), (20, 6), (20, 3), (19, 1), (19, 24)])
def test_splat_variants_full_shape(knob, value):
    """The experiment knobs of the fused path give the same results as the defaults, spread and clustered clouds:
    10 = 1 / 2 zeroing kernel + forward transform ahead of the grid dependency, 3 = the splat zeroes its grid itself behind a
    grid-wide barrier (default: cudaMemsetAsync); 11 = 0 8-byte /
    scalar reductions and gathers (default: 16-byte); 13 = 1 zeroing launch + dL/dscale atomics in the backward (default:
    folded partials); 14 = 0 wait-first backward splat; 15 = 0 x/y pass out of place, backward in the second grid;
    2 = 1 the producer warps of the x/y pipeline store the tiles; 4 = 1 the gathers of the splat backward inside the x/y pass
    of the backward (dpc_fused_bwd.cuh) + chain-rule kernel; 0 = 8 / 20 = 6 the tile-per-CTA forward / backward splat
    kernels instead of the software-pipelined ones, 20 = 3 predicated gathers in the tile-per-CTA backward; 19 = the
    number of warps per SM the pipelined kernels' grids are sized for (1: four tiles per warp at this shape, 24: one)."""
    from dpc_b200 import _capi
    L = _capi.lab_lib()             # the experiment knobs exist only in the lab build of the sources
    product, _capi._LIB = _capi._LIB, L
    default = {2: 0, 4: 0, 10: 0, 11: 1, 13: 0, 14: 1, 15: 1, 0: 4, 19: 0, 20: 0}[knob]
    cfg = default_config(vox_size=64, pc_gauss_kernel_size=21)
    _capi.check(L.dpc_debug_set(knob, value))
    try:
        for spread, seed in ((0.5, 1234), (0.025, 1235)):
            pc, q, sc, gt = _bench_inputs(2, 8000, 64, spread, seed=seed)
            _compare(_run_both(cfg, pc, q, sc, gt, 3.0))
    finally:
        L.dpc_debug_set(knob, default)
        _capi._LIB = product


@pytest.mark.parametrize("n,b,v", [(4001, 3, 32), (1000, 5, 32), (33, 2, 64)])
def test_pipelined_splat_ragged_multi_tile(n, b, v):
    """The software-pipelined splat kernels with several tiles per warp AND a ragged last tile (N % 32 != 0): the grid is
    sized for ONE warp per SM (lab knob 19 = 1), so at N = 4001, B = 3 every warp walks three tiles of its sample."""
    from dpc_b200 import _capi
    L = _capi.lab_lib()
    product, _capi._LIB = _capi._LIB, L
    cfg = default_config(vox_size=v, pc_gauss_kernel_size=11)
    _capi.check(L.dpc_debug_set(19, 1))
    try:
        g = torch.Generator().manual_seed(n)
        pc = torch.tanh(0.5 * torch.randn(b, n, 3, generator=g)) / 2
        q = torch.randn(b, 4, generator=g)
        sc = torch.sigmoid(torch.randn(b, 1, generator=g))
        gt = (torch.rand(b, v, v, 1, generator=g) > 0.5).float()
        tr = 0.05 * torch.randn(b, 3, generator=g)
        _compare(_run_both(cfg, pc, q, sc, gt, 3.0))
        _compare(_run_both(cfg, pc, q, sc, gt, 3.0, trans=tr))
    finally:
        L.dpc_debug_set(19, 0)
        _capi._LIB = product


def test_backward_does_not_depend_on_the_tr_pc_hint():
    """dpc_project_params.tr_pc only aims the prefetch of the pipelined splat backward (a lane whose recomputed cell is not
    the one its corners were requested for fetches them again): ONE forward at the headline shape, a backward, then tr_pc
    corrupted in place (rolled by 7 points, every fifth NaN, every fifth negated) and the backward again -- same d_pc bit
    for bit, per-sample sums (one atomic per warp, unordered) to rounding."""
    cfg = default_config(vox_size=64, pc_gauss_kernel_size=21)
    pc, q, sc, gt = _bench_inputs(8, 8000, 64, 0.5, seed=4321)
    leaves = [t.clone().to(DEV).requires_grad_(True) for t in (pc, q, sc)]
    out = pcm.pointcloud_project_fast(cfg, leaves[0], leaves[1], None, None, gk.smoothing_kernel(cfg, 3.0), leaves[2])
    up = (out["proj"].detach() - gt.to(DEV)) / 8
    good = [x.clone() for x in torch.autograd.grad(out["proj"], leaves, grad_outputs=up, retain_graph=True)]
    t = out["tr_pc"].data
    t.copy_(torch.roll(t, 7, dims=1))
    t[:, ::5] = float("nan")
    t[:, 1::5] *= -1.0
    bad = [x.clone() for x in torch.autograd.grad(out["proj"], leaves, grad_outputs=up)]
    assert torch.equal(good[0], bad[0]), "d_pc changed with the tr_pc hint"
    for g, b in zip(good[1:], bad[1:]):
        assert float((g - b).abs().max()) <= 1e-6 * max(1.0, float(g.abs().max()))


def test_max_projection_full_shape():
    cfg = default_config(vox_size=64, pc_gauss_kernel_size=21, ptn_max_projection=True)
    pc, q, sc, gt = _bench_inputs(2, 8000, 64, 0.5)
    _compare(_run_both(cfg, pc, q, sc, gt, 3.0))


@pytest.mark.parametrize("v,n,b", [(32, 2000, 8), (32, 32000, 4), (128, 2000, 2), (64, 16000, 2)])
def test_sweep_shapes_against_oracle(v, n, b):
    """BASELINE config 5 shapes (N in 2k..32k, V in 32/64/128) against the oracle."""
    cfg = default_config(vox_size=v, pc_gauss_kernel_size=21)
    g = torch.Generator().manual_seed(77)
    pc = torch.tanh(0.5 * torch.randn(b, n, 3, generator=g)) / 2
    q = torch.randn(b, 4, generator=g)
    sc = torch.sigmoid(torch.randn(b, 1, generator=g))
    gt = (torch.rand(b, v, v, 1, generator=g) > 0.5).float()
    _compare(_run_both(cfg, pc, q, sc, gt, 3.0))
    _compare(_run_both(cfg, pc, q, sc, gt, 3.0, host_sigma=True))


def test_transform_bit_exact_on_a_million_points():
    """The branch-free division of the camera transform (dpc_math.cuh: dpc_div) against IEEE division
    in the oracle, with translation and per-sample focal length: every bit of every coordinate."""
    cfg = default_config()
    g = torch.Generator().manual_seed(5)
    pc = torch.tanh(0.6 * torch.randn(32, 32768, 3, generator=g)) / 2
    q = torch.randn(32, 4, generator=g)
    tr = 0.05 * torch.randn(32, 3, generator=g)
    fl = 1.875 + 0.2 * torch.randn(32, 1, generator=g)
    a = pcm.pc_perspective_transform(cfg, pc.to(DEV), q.to(DEV), tr.to(DEV), fl.to(DEV)).cpu()
    assert torch.equal(a, O.pc_perspective_transform(cfg, pc, q, tr, fl))
    a = pcm.pc_perspective_transform(cfg, pc.to(DEV), q.to(DEV)).cpu()
    assert torch.equal(a, O.pc_perspective_transform(cfg, pc, q))


def test_properties_at_b32_n8000_v64():
    """Size-independent checks at BASELINE's full size (the oracle is too slow here to be the
    checker for every run): mass conservation, range, determinism of the index path, zero
    gradient for invalid points, batch independence."""
    cfg = default_config(vox_size=64, pc_gauss_kernel_size=21)
    pc, q, sc, gt = _bench_inputs(32, 8000, 64, 0.5)
    pc, q, sc, gt = pc.to(DEV), q.to(DEV), sc.to(DEV), gt.to(DEV)
    pcg = pc.clone().requires_grad_(True)
    ker = gk.smoothing_kernel(cfg, torch.tensor(3.0, device=DEV))
    out = pcm.pointcloud_project_fast(cfg, pcg, q, None, None, ker, sc)
    assert float(out["proj"].min()) >= 0.0 and float(out["proj"].max()) <= 1.0
    assert float(out["voxels"].min()) >= 0.0 and float(out["voxels"].max()) <= 1.0
    tr = out["tr_pc"].detach()
    valid = ((tr >= -0.5) & (tr <= 0.5)).all(-1)
    raw, _ = pcm.pointcloud2voxels3d_fast(cfg, tr, None)
    assert torch.allclose(raw.sum(dim=(1, 2, 3)), valid.sum(1).float(), atol=5e-2, rtol=1e-5)
    (((gt - out["proj"]) ** 2).sum() / 64).backward()
    assert torch.all(pcg.grad[~valid] == 0)
    assert torch.isfinite(pcg.grad).all()
    # a sample's result does not depend on its batch neighbours (sharding over ranks relies on it)
    out8 = pcm.pointcloud_project_fast(cfg, pc[8:16], q[8:16], None, None, ker, sc[8:16])
    assert torch.equal(out8["tr_pc"], tr[8:16])
    assert float((out8["proj"] - out["proj"][8:16]).abs().max()) <= 2e-6
    # termination probabilities of every ray sum to ~1
    s = out["drc_probs"].sum(0)
    assert float((s - 1).abs().max()) < 1e-3


def test_plus_half_dropped_and_nan_skipped():
    cfg = default_config(vox_size=16)
    q = torch.tensor([[1.0, 0, 0, 0]], device=DEV)
    pc = torch.tensor([[[0.5, 0.0, 0.0], [float("nan"), 0.0, 0.0], [0.0, 0.0, 0.0]]], device=DEV)
    out = pcm.pointcloud_project_fast(cfg, pc, q, None, None)
    ref = O.pointcloud_project_fast(cfg, pc.cpu(), q.cpu(), None, None)
    assert float((out["voxels"].cpu() - ref["voxels"]).abs().max()) <= 1e-6
    assert abs(float(out["voxels"].sum()) - 2.0) < 1e-5


def test_standalone_entry_points():
    cfg = default_config(vox_size=32, pc_gauss_kernel_size=11)
    g = torch.Generator().manual_seed(3)
    pc = (torch.rand(2, 500, 3, generator=g) - 0.5) * 1.1
    q = torch.randn(2, 4, generator=g)
    # pc_perspective_transform
    a = pcm.pc_perspective_transform(cfg, pc.to(DEV), q.to(DEV))
    assert torch.equal(a.cpu(), O.pc_perspective_transform(cfg, pc, q))
    # pointcloud2voxels3d_fast with gradient
    x = pc.clone().to(DEV).requires_grad_(True)
    xr = pc.clone().requires_grad_(True)
    w = torch.randn(2, 32, 32, 32, generator=g)
    v, none = pcm.pointcloud2voxels3d_fast(cfg, x, None)
    vr, _ = O.pointcloud2voxels3d_fast(cfg, xr, None)
    assert none is None
    assert float((v.cpu() - vr).abs().max()) <= 1e-5
    (v * w.to(DEV)).sum().backward()
    (vr * w).sum().backward()
    assert float((x.grad.cpu() - xr.grad).abs().max()) <= 1e-4
    # smoothen_voxels3d with gradient
    vox = torch.rand(2, 32, 32, 32, 1, generator=g)
    ker_c = gk.smoothing_kernel(cfg, torch.tensor(1.3, device=DEV))
    ker_o = O.smoothing_kernel(cfg, torch.tensor(1.3))
    for kc, ko in zip(ker_c, ker_o):
        assert kc.shape == ko.shape
        assert float((kc.cpu() - ko).abs().max()) <= 1e-7
    vc = vox.clone().to(DEV).requires_grad_(True)
    vo = vox.clone().requires_grad_(True)
    sc_ = pcm.smoothen_voxels3d(cfg, vc, ker_c)
    so_ = O.smoothen_voxels3d(cfg, vo, ker_o)
    assert float((sc_.cpu() - so_).abs().max()) <= 1e-5
    wv = torch.randn(2, 32, 32, 32, 1, generator=g)
    (sc_ * wv.to(DEV)).sum().backward()
    (so_ * wv).sum().backward()
    assert float((vc.grad.cpu() - vo.grad).abs().max()) <= 1e-5
    # drc_projection / depth with gradient
    pc_, pp = drc_mod.drc_projection(vc.detach().requires_grad_(True), cfg)
    po_, po = O.drc_projection(vo.detach(), cfg)
    assert float((pc_.cpu() - po_).abs().max()) <= 1e-5 and float((pp.cpu() - po).abs().max()) <= 1e-5
    d1 = drc_mod.drc_depth_projection(pp, cfg)
    assert float((d1.detach().cpu() - O.drc_depth_projection(po, cfg)).abs().max()) <= 1e-4


def test_point_dropout_gather():
    g = torch.Generator().manual_seed(4)
    pts = torch.randn(3, 100, 3, generator=g)
    rgb = torch.rand(3, 100, 3, generator=g)
    sel = torch.stack([torch.randperm(100, generator=g)[:37] for _ in range(3)])
    p = pts.clone().to(DEV).requires_grad_(True)
    op, orgb = pcm.pc_point_dropout(p, rgb.to(DEV), 0.37, selected_indices=sel)
    rp, rrgb = O.pc_point_dropout_with_indices(pts, rgb, sel)
    assert torch.equal(op.cpu(), rp) and torch.equal(orgb.cpu(), rrgb)
    op.sum().backward()
    assert float(p.grad.sum()) == 37 * 3 * 3
    assert pcm.num_points_after_dropout(8000, 0.07) == O.num_points_after_dropout(8000, 0.07) == 560
    op2, _ = pcm.pc_point_dropout(p, None, 0.5)
    assert op2.shape == (3, 50, 3)


def test_error_conventions():
    cfg = default_config(vox_size=16)
    with pytest.raises(ValueError):
        pcm.pointcloud_project_fast(cfg, torch.zeros(1, 4, 3, device=DEV), torch.zeros(1, 3, device=DEV), None, None)
    L = _capi.lib()
    p = _capi.ProjectParams(B=1, N=4, Vz=16, V=16, pose_kind=0, mode=0, K=0, Kz=0, focal_const=1.875,
                            cam_dist=2.0, clip_eps=1e-5, max_depth=10.0)
    assert L.dpc_project_fast_fwd(ctypes.byref(p), None, None, None, None, None, None, None, None, None, None, None,
                                  None, None, 0, None, 0, None) == -1
