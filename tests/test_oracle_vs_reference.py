"""CPU, build container only: the oracle restatement against the reference's own source run
live over the TF1 shim, on fresh random inputs (the committed fixtures cover the GPU box,
where /root/reference does not exist)."""
import numpy as np
import pytest
import torch

from oracle import dpc_oracle as O
from oracle import run_reference
from dpc_b200.util.config import default_config
from tests import cases

pytestmark = pytest.mark.skipif(not run_reference.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return run_reference.load()


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_inputs_bit_exact_transform_and_close_everything(ref, seed):
    tf = ref.tf
    cfg = default_config(vox_size=16, pc_gauss_kernel_size=7)
    g = torch.Generator().manual_seed(seed)
    pc = torch.tanh(0.7 * torch.randn(2, 257, 3, generator=g)) / 2
    q = torch.randn(2, 4, generator=g)
    sc = torch.sigmoid(torch.randn(2, 1, generator=g))
    tr = 0.1 * torch.randn(2, 3, generator=g)
    sigma = torch.tensor(1.1)
    k_ref = ref.gauss_kernel.smoothing_kernel(cfg, tf.Tensor(sigma))
    o_ref = ref.point_cloud.pointcloud_project_fast(cfg, tf.Tensor(pc), tf.Tensor(q), tf.Tensor(tr), None, k_ref, tf.Tensor(sc))
    o = O.pointcloud_project_fast(cfg, pc, q, tr, None, O.smoothing_kernel(cfg, sigma), sc)
    assert cases.nan_equal_bits(o["tr_pc"].numpy(), o_ref["tr_pc"].t.numpy())
    for k in ("proj", "voxels", "drc_probs", "proj_depth"):
        assert cases.max_abs_diff(o[k].numpy(), o_ref[k].t.numpy()) <= 1e-6, k


def test_python_float_sigma_taps(ref):
    cfg = default_config(pc_gauss_kernel_size=21)
    for s in (3.0, 0.2, 1.7):
        a = ref.gauss_kernel.smoothing_kernel(cfg, s)
        b = O.smoothing_kernel(cfg, s)
        for x, y in zip(a, b):
            assert np.array_equal(x.t.numpy(), y.numpy())


def test_standalone_entry_points(ref):
    tf = ref.tf
    cfg = default_config(vox_size=16)
    g = torch.Generator().manual_seed(5)
    pc = torch.rand(2, 100, 3, generator=g) - 0.5
    v_ref, _ = ref.point_cloud.pointcloud2voxels3d_fast(cfg, tf.Tensor(pc), None)
    v, _ = O.pointcloud2voxels3d_fast(cfg, pc, None)
    assert cases.max_abs_diff(v.numpy(), v_ref.t.numpy()) <= 1e-6
    vox = torch.rand(2, 16, 16, 16, 1, generator=g)
    p_ref, probs_ref = ref.drc.drc_projection(tf.Tensor(vox), cfg)
    p, probs = O.drc_projection(vox, cfg)
    assert cases.max_abs_diff(p.numpy(), p_ref.t.numpy()) <= 1e-6
    assert cases.max_abs_diff(probs.numpy(), probs_ref.t.numpy()) <= 1e-6
    d_ref = ref.drc.drc_depth_projection(probs_ref, cfg)
    assert cases.max_abs_diff(O.drc_depth_projection(probs, cfg).numpy(), d_ref.t.numpy()) <= 1e-6
