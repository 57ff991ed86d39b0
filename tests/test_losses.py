"""util/losses.proj_l2_loss = tf.nn.l2_loss(gt - pred) / num_samples (models/model_pc.py:414-415): value and gradient in
one kernel.  CPU: the real kernel source under the emulation against plain torch; GPU: the CUDA kernel."""
import pytest
import torch


def _reference(gt, pred, n):
    pred = pred.clone().requires_grad_(True)
    loss = ((gt - pred) ** 2).sum() / 2 / n
    loss.backward()
    return float(loss), pred.grad


def _check(device):
    from dpc_b200.util.losses import proj_l2_loss
    g = torch.Generator().manual_seed(5)
    for shape, n in (((32, 64, 64, 1), 32), ((3, 5, 7, 1), 3), ((2, 33), 2), ((1, 1), 1)):
        gt = (torch.rand(*shape, generator=g) > 0.5).float()
        pred = torch.rand(*shape, generator=g)
        want, want_g = _reference(gt, pred, n)
        p = pred.to(device).requires_grad_(True)
        loss = proj_l2_loss(gt.to(device), p, n)
        (2.0 * loss).backward()           # upstream gradient other than one
        assert abs(float(loss) - want) <= 1e-5 * max(1.0, abs(want)), (shape, float(loss), want)
        assert float((p.grad.cpu() - 2.0 * want_g).abs().max()) <= 1e-6
        again = proj_l2_loss(gt.to(device), p.detach(), n)     # the workspace is reusable and the sum deterministic
        assert float(again) == float(loss)
        # gradient-only form of the C-ABI entry point (loss = NULL, no workspace)
        from dpc_b200 import _capi
        pd, gd = pred.to(device).contiguous(), gt.to(device).contiguous()
        gp = torch.empty_like(pd)
        _capi.check(_capi.lib().dpc_proj_l2_loss(_capi.ptr(pd), _capi.ptr(gd), pd.numel(), 1.0 / n, None, _capi.ptr(gp), None, 0,
                                                 _capi.stream_of(pd)))
        assert float((gp.cpu() - want_g).abs().max()) <= 1e-6


def test_emulated_loss_kernel():
    from tests.emu_support import build_emu
    from dpc_b200 import _capi
    import dpc_b200.util.losses as losses
    lib = _capi.load_library(build_emu())
    old = (_capi._LIB, _capi._REQUIRE_CUDA)
    _capi._LIB, _capi._REQUIRE_CUDA = lib, False
    losses._WORK.clear()
    try:
        _check(torch.device("cpu"))
    finally:
        _capi._LIB, _capi._REQUIRE_CUDA = old
        losses._WORK.clear()


@pytest.mark.gpu
def test_cuda_loss_kernel():
    _check(torch.device("cuda:0"))
