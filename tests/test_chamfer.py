"""f-4: point_cloud_distance (util/point_cloud_distance.py:26-39).  CPU: the oracle against the reference's own
source run through the TF1 shim, a brute-force known answer, and the real kernel sources under the CPU emulation.
GPU: the CUDA kernels against the oracle, bit-exact indices and distances."""
import os

import numpy as np
import pytest
import torch

from oracle import dpc_oracle as O
from oracle import run_reference


def _sets(ns, nt, dtype, seed=0, dup=True):
    g = torch.Generator().manual_seed(seed)
    vs = (torch.tanh(0.5 * torch.randn(ns, 3, generator=g, dtype=torch.float64)) / 2).to(dtype)
    vt = (torch.tanh(0.5 * torch.randn(nt, 3, generator=g, dtype=torch.float64)) / 2).to(dtype)
    if dup and nt >= 8:
        vt[nt // 2] = vt[3]              # exact duplicates: the FIRST index has to win
        vt[nt - 1] = vt[3]
        vs[0] = vt[3]                    # distance exactly 0
    return vs, vt


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_matches_reference_source(dtype):
    if not run_reference.available():
        pytest.skip("reference not present")
    ref = run_reference.load()
    tf = ref.tf
    vs, vt = _sets(300, 700, dtype, seed=1)
    proj, md, idx = ref.point_cloud_distance.point_cloud_distance(
        tf.constant(vs.numpy(), dtype=dtype), tf.constant(vt.numpy(), dtype=dtype))   # float64 placeholders: eval_chamfer.py:50-51
    o_proj, o_md, o_idx = O.point_cloud_distance(vs, vt)
    assert np.array_equal(np.asarray(idx.numpy()), o_idx.numpy())
    assert np.array_equal(np.asarray(md.numpy()), o_md.numpy())
    assert np.array_equal(np.asarray(proj.numpy()), o_proj.numpy())


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chamfer")
GOLDEN_NAMES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_matches_golden_fixture(name):
    """Fixtures generated from the reference's own source (tests/golden/make_golden_chamfer.py): they travel, the
    reference does not."""
    fx = _load(name)
    proj, md, idx = O.point_cloud_distance(fx["vs"], fx["vt"])
    assert torch.equal(idx, fx["idx"]) and torch.equal(md, fx["min_dist"]) and torch.equal(proj, fx["proj"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_cuda_kernels_match_golden_fixture(name):
    import dpc_b200.util.point_cloud_distance as pcd
    fx = _load(name)
    proj, md, idx = pcd.point_cloud_distance(fx["vs"].cuda(), fx["vt"].cuda())
    assert torch.equal(idx.cpu(), fx["idx"]) and torch.equal(md.cpu(), fx["min_dist"]) and torch.equal(proj.cpu(), fx["proj"])


def test_oracle_known_answers():
    vt = torch.tensor([[0., 0., 0.], [1., 0., 0.], [0., 2., 0.], [1., 0., 0.]], dtype=torch.float64)
    vs = torch.tensor([[0.9, 0., 0.], [0., 1.1, 0.], [0.5, 0., 0.], [3., 4., 0.]], dtype=torch.float64)
    proj, md, idx = O.point_cloud_distance(vs, vt)
    assert idx.tolist() == [1, 2, 0, 2]          # duplicate target 3 never wins; the tie at 0.5 goes to index 0
    assert torch.allclose(md, torch.tensor([0.1, 0.9, 0.5, np.sqrt(13.0)], dtype=torch.float64), atol=1e-15)
    assert torch.equal(proj, vt[idx.long()])


def _check(impl, vs, vt):
    proj, md, idx = impl(vs, vt)
    o_proj, o_md, o_idx = O.point_cloud_distance(vs, vt)
    assert idx.dtype == torch.int32 and md.dtype == vs.dtype and proj.dtype == vs.dtype
    assert torch.equal(idx.cpu(), o_idx), "nearest-neighbour indices must be bit-exact"
    assert torch.equal(md.cpu(), o_md), "distances must be bit-exact"
    assert torch.equal(proj.cpu(), o_proj)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("ns,nt", [(1, 1), (5, 3000), (300, 1025), (700, 333)])
def test_emulated_kernels_match_oracle(ns, nt, dtype):
    from tests.emu_support import build_emu
    from dpc_b200 import _capi
    import dpc_b200.util.point_cloud_distance as pcd
    lib = _capi.load_library(build_emu())
    old = (_capi._LIB, _capi._REQUIRE_CUDA)
    _capi._LIB, _capi._REQUIRE_CUDA = lib, False
    try:
        _check(pcd.point_cloud_distance, *_sets(ns, nt, dtype, seed=ns + nt))
    finally:
        _capi._LIB, _capi._REQUIRE_CUDA = old


def test_argument_errors():
    import dpc_b200.util.point_cloud_distance as pcd
    with pytest.raises(ValueError):
        pcd.point_cloud_distance(torch.zeros(4, 2), torch.zeros(4, 3))
    with pytest.raises(ValueError):
        pcd.point_cloud_distance(torch.zeros(4, 3), torch.zeros(4, 3, dtype=torch.float64))
    with pytest.raises(ValueError):
        pcd.point_cloud_distance(torch.zeros(0, 3), torch.zeros(4, 3))
    with pytest.raises((ValueError, RuntimeError)):     # CPU tensors: there is no CPU path
        pcd.point_cloud_distance(torch.zeros(4, 3), torch.zeros(4, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("ns,nt", [(1, 1), (17, 5), (8000, 10000), (2500, 40000), (10000, 8000)])
def test_cuda_kernels_match_oracle(ns, nt, dtype):
    import dpc_b200.util.point_cloud_distance as pcd
    vs, vt = _sets(ns, nt, dtype, seed=ns + 7 * nt)
    _check(lambda a, b: pcd.point_cloud_distance(a.cuda(), b.cuda()), vs, vt)


@pytest.mark.gpu
def test_cuda_chamfer_properties_at_evaluation_size():
    """Size-independent properties at the evaluation's size (8000 predicted vs 100k ground-truth points, fp64):
    a set against itself is at distance 0 with idx = identity (first duplicate aside); distances are symmetric under
    swapping the roles for mutual nearest neighbours; the result does not depend on how the targets are split."""
    import dpc_b200.util.point_cloud_distance as pcd
    vs, vt = _sets(8000, 100000, torch.float64, seed=5, dup=False)
    vs, vt = vs.cuda(), vt.cuda()
    proj, md, idx = pcd.point_cloud_distance(vs, vs)
    assert float(md.abs().max()) == 0.0 and torch.equal(idx.long(), torch.arange(8000, device="cuda"))
    proj, md, idx = pcd.point_cloud_distance(vs, vt)
    assert torch.equal(proj, vt[idx.long()])
    assert torch.equal(md, torch.sqrt(((proj - vs) ** 2)[:, 0] + ((proj - vs) ** 2)[:, 1] + ((proj - vs) ** 2)[:, 2]))
    _, md_b, idx_b = pcd.point_cloud_distance(vt, vs)
    mutual = idx_b[idx.long()].long() == torch.arange(8000, device="cuda")
    assert int(mutual.sum()) > 0
    assert torch.equal(md[mutual], md_b[idx.long()][mutual])
    # halves of the target set, folded by hand, give the same answer
    h = 50000
    _, m0, i0 = pcd.point_cloud_distance(vs, vt[:h].contiguous())
    _, m1, i1 = pcd.point_cloud_distance(vs, vt[h:].contiguous())
    pick1 = m1 < m0
    assert torch.equal(torch.where(pick1, m1, m0), md)
    assert torch.equal(torch.where(pick1, i1 + h, i0), idx)
