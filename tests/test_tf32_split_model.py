"""CPU model of the 3xTF32 split the tensor-core smoothing kernels use (csrc/dpc_smooth_tc.cuh: dpc_tc_hi, the lo plane,
the Toeplitz operand split the same way, D = a_lo*t_hi + a_hi*t_lo + a_hi*t_hi accumulated in fp32, the tensor core
ignoring the low 13 mantissa bits of every operand).  Pins the claim of DESIGN section 4 -- fp32-level accuracy, far inside
the 1e-5 gate -- on the CPU; the kernels themselves are checked on the GPU by tests/test_gpu_families.py."""
import numpy as np
import pytest


def tf32_hi(v):
    """round to nearest (ties away) on the 13 dropped mantissa bits: (bits + 0x1000) & 0xffffe000, as dpc_tc_hi"""
    u = v.astype(np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_trunc(v):
    """what the tensor core reads of an fp32 operand register: the low 13 mantissa bits are ignored"""
    return (v.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def gauss_taps(k, sigma_rel, v=64):
    x = np.arange((-k) // 2 + 1.0, k // 2 + 1.0)
    sig = sigma_rel
    t = np.exp(-(x ** 2) / (2.0 * sig ** 2))
    return (t / t.sum()).astype(np.float32)


@pytest.mark.parametrize("k,sigma", [(21, 3.0), (11, 1.0), (21, 0.2), (63, 10.0), (1, 1.0)])
def test_three_term_split_matches_fp64_correlation(k, sigma):
    rng = np.random.default_rng(k)
    rows = rng.random((256, 64), dtype=np.float32)            # clipped occupancies in [0, 1]
    rows[::7] = 0.0
    rows[3] = 1.0
    taps = gauss_taps(k, sigma)
    pl = (k - 1) // 2
    # Toeplitz operand T[n][kk] = tap(kk - n + pl)
    T = np.zeros((64, 64), dtype=np.float32)
    for n in range(64):
        for kk in range(64):
            j = kk - n + pl
            if 0 <= j < k:
                T[n, kk] = taps[j]
    want = rows.astype(np.float64) @ T.astype(np.float64).T
    a_hi = tf32_hi(rows)
    a_lo = tf32_trunc((rows - a_hi).astype(np.float32))       # exact difference, then what the MMA sees of it
    t_hi = tf32_hi(T)
    t_lo = tf32_trunc(tf32_hi((T - t_hi).astype(np.float32)))
    assert np.array_equal(tf32_trunc(a_hi), a_hi) and np.array_equal(tf32_trunc(t_hi), t_hi)   # hi planes are exact tf32 values
    # fp32 accumulation of the three GEMMs in the kernel's order (products of tf32 values are exact in fp32-width pairs;
    # the accumulation order inside the tensor core is unspecified -- model it with fp32 partial sums per k-step of 8)
    acc = np.zeros_like(want, dtype=np.float32)
    for A, Bm in ((a_lo, t_hi), (a_hi, t_lo), (a_hi, t_hi)):
        for k0 in range(0, 64, 8):
            acc = (acc + (A[:, k0:k0 + 8].astype(np.float64) @ Bm[:, k0:k0 + 8].astype(np.float64).T).astype(np.float32)).astype(np.float32)
    err = np.abs(acc.astype(np.float64) - want).max()
    assert err <= 5e-7, err
    # and the single-pass tf32 product alone would NOT be good enough for the 1e-5 gate on sums of 21 taps
    naive = (a_hi.astype(np.float64) @ t_hi.astype(np.float64).T)
    if k >= 11 and sigma >= 1.0:
        assert np.abs(naive - want).max() > 2e-5
