"""CPU: the oracle restatement reproduces the fixtures generated from the reference's own
source (tests/golden/make_golden.py) -- bit-exact tr_pc / indices / taps, 1e-6 elsewhere."""
import numpy as np
import pytest
import torch

from oracle import dpc_oracle
from tests import cases


@pytest.mark.parametrize("name", cases.golden_names())
def test_oracle_matches_golden(name):
    fx = cases.load_golden(name)
    outs, grads = cases.run_impl(dpc_oracle, fx)
    assert cases.nan_equal_bits(outs["tr_pc"].numpy(), fx["out_tr_pc"]), "tr_pc must be bit-exact"
    cfg = cases.make_cfg(fx["cfg_over"])
    valid, idx, _ = dpc_oracle.voxel_indices(cfg, outs["tr_pc"])
    assert np.array_equal(valid.numpy(), fx["valid"])
    assert np.array_equal(idx.numpy()[fx["valid"]], fx["idx"][fx["valid"]]), "voxel indices must be bit-exact"
    for k in cases.OUTPUT_KEYS:
        ref = fx.get("out_" + k)
        if ref is None:
            assert outs[k] is None, k
            continue
        assert outs[k] is not None, k
        assert tuple(outs[k].shape) == ref.shape, k
        assert cases.max_abs_diff(outs[k].numpy(), ref) <= 1e-6, k
    for k, v in fx.items():
        if k.startswith("grad_"):
            g = grads[k[5:]].numpy()
            scale = max(1.0, float(np.nanmax(np.abs(v))))
            assert cases.max_abs_diff(g, v) <= 1e-6 * scale, k


@pytest.mark.parametrize("name", [n for n in cases.golden_names() if n not in ("no_kernel_no_scale",)])
def test_oracle_taps_bit_exact(name):
    fx = cases.load_golden(name)
    cfg = cases.make_cfg(fx["cfg_over"])
    ker = dpc_oracle.smoothing_kernel(cfg, torch.tensor(float(fx["in_sigma"]), dtype=torch.float32))
    for i in range(3):
        assert ker[i].shape == fx["kernel%d" % i].shape
        assert np.array_equal(ker[i].numpy(), fx["kernel%d" % i])
