"""Drop-in for the reference's `dpc/util/gauss_kernel.py` (runtime Gaussian taps).

  /root/reference/dpc/util/gauss_kernel.py:5   gauss_kernel_1d
  /root/reference/dpc/util/gauss_kernel.py:27  separable_kernels
  /root/reference/dpc/util/gauss_kernel.py:35  smoothing_kernel

sigma is a run-time value (a function of the global step, model_pc.py:35-40), so the taps are a
device buffer handed to the kernels, never a compile-time constant.  They are a few dozen
floats; building them is plain torch on whatever device sigma lives on.  When sigma is known on
the HOST (a Python float or a CPU tensor -- the reference's schedule is a function of the step
count) the taps are built on the CPU, exactly as the oracle builds them, and the kernel object
keeps that host copy next to the device copy: the fused path then passes the taps to the
smoothing kernels as launch parameters (uniform registers; see csrc/dpc_smooth_fast.cuh).
"""
import math

import torch

from .point_cloud import SeparableKernel


def gauss_kernel_1d(l, sig):
    """Gaussian taps of length l: x = range(-l//2 + 1, l//2 + 1) with Python's precedence
    (l=21 -> -10..10, l=10 -> -4..5), exp(-x^2 / (2 sig^2)), normalised to sum 1."""
    lo, hi = (-l) // 2 + 1.0, l // 2 + 1.0
    if torch.is_tensor(sig):
        xx = torch.arange(lo, hi, dtype=torch.float32, device=sig.device)
        sig = sig.to(torch.float32)
    else:
        xx = torch.arange(lo, hi, dtype=torch.float32)
    k = torch.exp(-xx ** 2 / (2.0 * sig ** 2))
    return k / k.sum()


def _attach_taps(out, taps_xy, taps_z):
    out.taps_xy = taps_xy.contiguous()
    out.taps_z = out.taps_xy if taps_z is taps_xy else taps_z.contiguous()
    if out.taps_xy.device.type == "cpu":
        out.host_taps_xy = out.taps_xy.detach()
        out.host_taps_z = out.taps_z.detach()
    return out


def separable_kernels(kernel):
    size = kernel.shape[0]
    out = SeparableKernel([kernel.reshape(1, 1, size, 1, 1), kernel.reshape(1, size, 1, 1, 1),
                           kernel.reshape(size, 1, 1, 1, 1)])
    return _attach_taps(out, kernel, kernel)


def smoothing_kernel(cfg, sigma):
    """[k1 (1,1,K,1,1), k2 (1,K,1,1,1), k3 (Kz,1,1,1,1)]; with cfg.vox_size_z != -1 the depth filter
    uses sigma*Vz/V and floor(K*Vz/V) taps bumped to odd."""
    fsz = int(cfg.pc_gauss_kernel_size)
    k1d = gauss_kernel_1d(fsz, sigma)
    if int(cfg.vox_size_z) != -1:
        ratio = cfg.vox_size_z / cfg.vox_size
        fsz_z = int(math.floor(fsz * ratio))
        if fsz_z % 2 == 0:
            fsz_z += 1
        kz = gauss_kernel_1d(fsz_z, sigma * ratio)
        out = SeparableKernel([k1d.reshape(1, 1, fsz, 1, 1), k1d.reshape(1, fsz, 1, 1, 1),
                               kz.reshape(fsz_z, 1, 1, 1, 1)])
        return _attach_taps(out, k1d, kz)
    if not cfg.pc_separable_gauss_filter:
        # the reference reaches an unbound local here (gauss_kernel.py:51-54)
        raise NotImplementedError("pc_separable_gauss_filter=false has no kernel in the reference either")
    return separable_kernels(k1d)


def gauss_smoothen_image(cfg, img, sigma_rel):
    """[B,H,W,C] image smoothed by the separable Gaussian of pc_gauss_kernel_size taps, every channel on its own,
    zero 'SAME' padding: two tf.nn.depthwise_conv2d calls, along W then along H (gauss_kernel.py:14-24).  Library
    convolution (the loss side of the train step is library code)."""
    import torch.nn.functional as F
    fsz = int(cfg.pc_gauss_kernel_size)
    sig = sigma_rel.to(img.device) if torch.is_tensor(sigma_rel) else sigma_rel
    k = gauss_kernel_1d(fsz, sig).to(device=img.device, dtype=img.dtype)
    c = img.shape[-1]
    x = img.permute(0, 3, 1, 2)
    lo, hi = (fsz - 1) // 2, (fsz - 1) - (fsz - 1) // 2
    x = F.conv2d(F.pad(x, [lo, hi, 0, 0]), k.reshape(1, 1, 1, fsz).expand(c, 1, 1, fsz).contiguous(), groups=c)
    x = F.conv2d(F.pad(x, [0, 0, lo, hi]), k.reshape(1, 1, fsz, 1).expand(c, 1, fsz, 1).contiguous(), groups=c)
    return x.permute(0, 2, 3, 1).contiguous()
