"""N1: dL/dsigma.  The smoothing taps are a differentiable function of sigma (gauss_kernel.py:5-11); when sigma requires
a gradient the product returns dL/d(taps) of every pass (dpc_tap_corr) and torch autograd carries it to sigma.  Checked
against the oracle's autograd (the reference's own op graph) and a central finite difference of the oracle in float64."""
import pytest
import torch

import dpc_b200.util.gauss_kernel as gk
import dpc_b200.util.point_cloud as pcm
from dpc_b200.util.config import default_config
from oracle import dpc_oracle as O
from tests.emu_support import emu  # noqa: F401


def _case(device, b, n, v, ksize, sigma0, vz=-1, rtol=2e-4):
    over = dict(vox_size=v, pc_gauss_kernel_size=ksize)
    if vz != -1:
        over["vox_size_z"] = vz
    cfg = default_config(**over)
    g = torch.Generator().manual_seed(31)
    pc = torch.tanh(0.5 * torch.randn(b, n, 3, generator=g)) / 2
    q = torch.randn(b, 4, generator=g)
    sc = torch.sigmoid(torch.randn(b, 1, generator=g))
    gt = (torch.rand(b, v, v, 1, generator=g) > 0.5).float()

    def run(mod_pc, mod_gk, dev, dtype=torch.float32, sig=sigma0):
        sigma = torch.tensor(sig, dtype=dtype, device=dev, requires_grad=True)
        leaves = [t.clone().to(device=dev, dtype=dtype).requires_grad_(True) for t in (pc, q, sc)]
        ker = mod_gk.smoothing_kernel(cfg, sigma)
        if dtype != torch.float32:
            ker = [k.to(dtype) for k in ker]
        out = mod_pc.pointcloud_project_fast(cfg, leaves[0], leaves[1], None, None, ker, leaves[2])
        loss = ((gt.to(device=dev, dtype=dtype) - out["proj"]) ** 2).sum() / 2 / b
        grads = torch.autograd.grad(loss, [sigma] + leaves)
        return float(loss), [x.detach().cpu().double() for x in grads]

    loss_c, gc = run(pcm, gk, device)
    loss_o, go = run(O, O, "cpu")
    assert abs(loss_c - loss_o) <= 1e-5 * max(1.0, abs(loss_o))
    ds_c, ds_o = float(gc[0]), float(go[0])
    assert abs(ds_c - ds_o) <= rtol * max(1.0, abs(ds_o)), (ds_c, ds_o)
    for a, r in zip(gc[1:], go[1:]):            # the other gradients are unchanged by the route
        assert float((a - r).abs().max()) <= 1e-5 * max(1.0, float(r.abs().max()))
    # finite difference of the float64 oracle (sigma enters only through the taps)
    h = 1e-4
    lp, _ = run(O, O, "cpu", torch.float64, sigma0 + h)
    lm, _ = run(O, O, "cpu", torch.float64, sigma0 - h)
    fd = (lp - lm) / (2 * h)
    assert abs(ds_c - fd) <= 5e-3 * max(1.0, abs(fd)), (ds_c, fd)
    return ds_c


def test_sigma_gradient_emulated(emu):  # noqa: F811
    assert _case("cpu", b=2, n=300, v=16, ksize=5, sigma0=0.9) != 0.0
    _case("cpu", b=1, n=200, v=16, ksize=7, sigma0=1.2, vz=8)      # depth taps with their own sigma / length


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 8000, 64, 21, 3.0), (2, 1000, 32, 11, 1.0), (2, 2000, 64, 21, 0.6)])
def test_sigma_gradient_gpu(shape):
    b, n, v, k, s = shape
    assert _case("cuda:0", b, n, v, k, s) != 0.0
