// Tensor-core (tcgen05 + TMEM) smoothing / projection kernels for the 64^3 grid.
//
// Why tensor cores on a "bandwidth-bound" path: at K = 21 the separable Gaussian of
// smoothen_voxels3d (point_cloud.py:139-145) costs 63 FMA per voxel per direction.  On the CUDA
// cores that is >= 29 us of pure FFMA2 issue per B=32 forward+backward step against ~35 us of HBM
// time, and the FFMA2 kernels (dpc_smooth_fast.cuh) sit at ~45 % FMA-pipe utilisation -- the
// FP32 pipe, not HBM, bounds them.  A 1-D correlation along an axis of length 64 is the product
// with a 64 x 64 banded Toeplitz matrix of the taps, so each pass becomes
//     D[128 rows x 64] = A[128 x 64] * T[64 x 64]
// on the 5th-generation tensor cores (tcgen05.mma kind::tf32, M=128 N=64 K=8, accumulator in
// TMEM).  fp32 accuracy is kept with the 3xTF32 split: a = a_hi + a_lo, t = t_hi + t_lo (each half
// exactly representable in tf32), D = a_lo*t_hi + a_hi*t_lo + a_hi*t_hi accumulated in fp32; the
// dropped a_lo*t_lo term is < 2^-22 relative.  That leaves per voxel and pass ~3 ALU instructions
// (split), two shared-memory stores and 1/32 of a TMEM load instead of 10.5 FFMA2, and the
// kernels become memory-bound as the byte count says they should be.
//
// Layout trick: with M = 128 the accumulator row m lives in TMEM lane m, and tcgen05.ld
// 32x32b gives THREAD m all 64 columns of row m.  So
//   * depth pass: row = ray (two image rows x 64 x), column = depth level: after the MMA every
//     thread holds the whole smoothed ray in registers -- exactly what the scale / clip / DRC
//     scan of pointcloud_project_fast needs (point_cloud.py:249-267, drc.py:47-123);
//   * x pass: row = (slice, y), column = x.  Thread (slice, y') then writes its 64 values as
//     COLUMN y' of the second operand A2[(slice, x)][y'] -- 32 lanes write 128 contiguous bytes, no
//     bank conflicts -- which transposes the slice for free; the y pass returns row = (slice, x),
//     column = y, and the global store is coalesced again (lanes = consecutive x).
//
// Operands are written to shared memory by the CUDA cores (they have to pass through registers for
// the split anyway) in the canonical K-major SWIZZLE_128B layout: rows of 128 bytes (32 tf32),
// 16-byte chunk c of row r stored at chunk c ^ (r & 7), 8-row groups 1024 bytes apart (SBO); the
// 64-long K extent is two such blocks ("halves").
//
// Not emulated: tests/emu covers the FFMA2 kernels (same maths); these kernels are checked on the
// GPU against the oracle, the fixtures and the FFMA2 kernels.
#pragma once
#include "dpc_common.cuh"
#include "dpc_smooth.cuh"
#include "dpc_smooth_fast.cuh"

#ifndef DPC_EMU

#define DPC_TC_THREADS 128
#define DPC_TC_A_LO 32768u        // byte offsets from the 1024-aligned smem base
#define DPC_TC_T_HI 65536u
#define DPC_TC_T_LO 81920u
#define DPC_TC_A_HALF 16384u      // 128 rows x 128 B
#define DPC_TC_T_HALF 8192u       //  64 rows x 128 B
#define DPC_TC_SMEM_BYTES (98304 + 1024)
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1<<4), a/b_format TF32 (2<<7, 2<<10),
// A and B K-major, N>>3 = 8 at bit 17, M>>4 = 8 at bit 24
#define DPC_TC_IDESC 0x08100910u

// dpc_debug_set key 8: 0 = CUDA-core (FFMA2 / generic) kernels only, 1 = single-tile tensor-core kernels (this file;
// kept as the simple reference form of the maths), 2 = persistent tensor-core pipelines (dpc_smooth_tcp.cuh; default)
static int dpc_tc_enable = 2;

DPC_DEV uint32_t dpc_tc_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(16 B)<<16 | SBO(1024 B)<<32 |
// version 1 <<46 | SWIZZLE_128B (2) << 61
DPC_DEV uint64_t dpc_tc_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

DPC_DEV void dpc_tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(DPC_TC_IDESC), "r"(accum) : "memory");
}
DPC_DEV void dpc_tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dpc_tc_s32(bar)) : "memory");
}
// one lane of a converged warp (elect.sync): with a warp-uniform branch around it the compiler keeps the MMA
// descriptors in uniform registers and emits the tcgen05.mma back to back; a divergent `lane == 0` branch costs
// ~70 cycles of issue per MMA instead (R2UR / ELECT / BRA.U.ANY per instruction; measured, scripts/mma_bench.py)
DPC_DEV uint32_t dpc_elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred));
  return pred;
}
DPC_DEV void dpc_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DPC_DEV void dpc_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
DPC_DEV void dpc_tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// whole warp: allocate `ncols` TMEM columns (power of two >= 32), address written to *slot
DPC_DEV void dpc_tc_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dpc_tc_s32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
DPC_DEV void dpc_tc_dealloc(uint32_t tmem, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols) : "memory");
}

// this thread's TMEM lane, 32 consecutive columns -> r[0..31]
DPC_DEV void dpc_tc_ld32(uint32_t taddr, float* r) {
  uint32_t u[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
        "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
        "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(u[i]);
}

// v = hi + lo: hi = v rounded to tf32 (10 explicit mantissa bits; round to nearest, ties away, done on the
// integer pipe -- cvt.rna.tf32.f32 is emulated with five instructions on sm_100a), lo = v - hi exactly (fp32).
// The tensor core ignores the low 13 mantissa bits of lo, an error <= 2^-21 |v|.  No inf/nan handling:
// an infinite voxel would poison its whole row through 0 * inf anyway.
DPC_DEV float dpc_tc_hi(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }

// this thread's TMEM lane, 32 consecutive columns <- r[0..31]
DPC_DEV void dpc_tc_st32(uint32_t taddr, const float* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16,"
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
      ::"r"(taddr),
        "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])), "r"(__float_as_uint(r[3])),
        "r"(__float_as_uint(r[4])), "r"(__float_as_uint(r[5])), "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7])),
        "r"(__float_as_uint(r[8])), "r"(__float_as_uint(r[9])), "r"(__float_as_uint(r[10])), "r"(__float_as_uint(r[11])),
        "r"(__float_as_uint(r[12])), "r"(__float_as_uint(r[13])), "r"(__float_as_uint(r[14])), "r"(__float_as_uint(r[15])),
        "r"(__float_as_uint(r[16])), "r"(__float_as_uint(r[17])), "r"(__float_as_uint(r[18])), "r"(__float_as_uint(r[19])),
        "r"(__float_as_uint(r[20])), "r"(__float_as_uint(r[21])), "r"(__float_as_uint(r[22])), "r"(__float_as_uint(r[23])),
        "r"(__float_as_uint(r[24])), "r"(__float_as_uint(r[25])), "r"(__float_as_uint(r[26])), "r"(__float_as_uint(r[27])),
        "r"(__float_as_uint(r[28])), "r"(__float_as_uint(r[29])), "r"(__float_as_uint(r[30])), "r"(__float_as_uint(r[31]))
      : "memory");
}
DPC_DEV void dpc_tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A operand from TMEM (row m in lane m, K element k in column a_tmem + k, 32 bits each), B from shared memory
DPC_DEV void dpc_tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(DPC_TC_IDESC), "r"(accum) : "memory");
}
// 32 values of row-half `h` of this thread's operand row -> hi / lo planes in TMEM (columns a_hi + 32h .., a_lo + 32h ..)
DPC_DEV void dpc_tc_split_st32(uint32_t a_hi_t, uint32_t a_lo_t, const float* v) {
  float hi[32], lo[32];
  const float2 m1 = make_float2(-1.0f, -1.0f);
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    hi[j] = dpc_tc_hi(v[j]); hi[j + 1] = dpc_tc_hi(v[j + 1]);
    const float2 l = __ffma2_rn(make_float2(hi[j], hi[j + 1]), m1, make_float2(v[j], v[j + 1]));
    lo[j] = l.x; lo[j + 1] = l.y;
  }
  dpc_tc_st32(a_hi_t, hi);
  dpc_tc_st32(a_lo_t, lo);
}
// One elected thread: D = A * T with A = (a_hi, a_lo) planes of 64 TMEM columns each, T = (hi, lo) in shared memory
DPC_DEV void dpc_tc_issue_ts(uint32_t a_hi_t, uint32_t a_lo_t, uint32_t t_base, uint32_t d_tmem, uint64_t* bar) {
  const uint64_t dT = dpc_tc_desc(t_base);
  uint32_t accum = 0;
#pragma unroll
  for (int term = 0; term < 3; ++term) {          // a_lo*t_hi, a_hi*t_lo, a_hi*t_hi
    const uint32_t at = (term == 0) ? a_lo_t : a_hi_t;
    const uint32_t to = (term == 1) ? 2u * DPC_TC_T_HALF : 0u;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const uint64_t bd = dT + (uint64_t)((to + (uint32_t)(kk >> 2) * DPC_TC_T_HALF + (uint32_t)(kk & 3) * 32u) >> 4);
      dpc_tc_mma_ts(d_tmem, at + (uint32_t)(8 * kk), bd, accum);
      accum = 1;
    }
  }
  dpc_tc_commit(bar);
}

// four consecutive K elements (k0 = 32*half + 4*chunk) of operand row `row`, hi and lo planes
DPC_DEV void dpc_tc_store4(unsigned char* sm, int row, int half, int chunk, float a, float b, float c, float d) {
  float4 hi, lo;
  hi.x = dpc_tc_hi(a); hi.y = dpc_tc_hi(b); hi.z = dpc_tc_hi(c); hi.w = dpc_tc_hi(d);
  const float2 m1 = make_float2(-1.0f, -1.0f);
  const float2 l0 = __ffma2_rn(make_float2(hi.x, hi.y), m1, make_float2(a, b));     // v - hi, exact
  const float2 l1 = __ffma2_rn(make_float2(hi.z, hi.w), m1, make_float2(c, d));
  lo = make_float4(l0.x, l0.y, l1.x, l1.y);
  const uint32_t off = (uint32_t)half * DPC_TC_A_HALF + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
  *reinterpret_cast<float4*>(sm + off) = hi;
  *reinterpret_cast<float4*>(sm + DPC_TC_A_LO + off) = lo;
}

// One elected thread: D[128 x 64] (TMEM columns d_tmem .. d_tmem+63) = A * T, three tf32 products, then
// commit -> one arrival on `bar` once every MMA has completed (and has finished reading shared memory).
// a_base: A_hi plane (A_lo 32 KiB above it); t_base: T_hi plane (T_lo 16 KiB above it); shared-window addresses.
DPC_DEV void dpc_tc_issue2(uint32_t a_base, uint32_t t_base, uint32_t d_tmem, uint64_t* bar) {
  const uint64_t dA = dpc_tc_desc(a_base), dT = dpc_tc_desc(t_base);
  uint32_t accum = 0;
#pragma unroll
  for (int term = 0; term < 3; ++term) {          // small terms first: a_lo*t_hi, a_hi*t_lo, a_hi*t_hi
    const uint32_t ao = (term == 0) ? DPC_TC_A_LO : 0u;
    const uint32_t to = (term == 1) ? 2u * DPC_TC_T_HALF : 0u;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {               // K = 64 in steps of 8 tf32 (32 bytes inside the 128-byte swizzle row)
      const uint64_t ad = dA + (uint64_t)((ao + (uint32_t)(kk >> 2) * DPC_TC_A_HALF + (uint32_t)(kk & 3) * 32u) >> 4);
      const uint64_t bd = dT + (uint64_t)((to + (uint32_t)(kk >> 2) * DPC_TC_T_HALF + (uint32_t)(kk & 3) * 32u) >> 4);
      dpc_tc_mma(d_tmem, ad, bd, accum);
      accum = 1;
    }
  }
  dpc_tc_commit(bar);
}
DPC_DEV void dpc_tc_issue(uint32_t sbase, uint32_t d_tmem, uint64_t* bar) { dpc_tc_issue2(sbase, sbase + DPC_TC_T_HI, d_tmem, bar); }

// Toeplitz operand of a correlation with zero padding: out[n] = sum_j tap(j) in[n + j - pl]  =>
// T[n][k] = tap(k - n + pl); rows n (the MMA's N), K-major.  tap(j) = taps[rev ? K-1-j : j] inside 0..K-1
// (NULL taps = identity).  All 128 threads; contains a __syncthreads.
DPC_DEV void dpc_tc_build_toeplitz_from(unsigned char* tbase, const float* tp_hi, const float* tp_lo, int pl);
DPC_DEV void dpc_tc_build_toeplitz(unsigned char* tbase, float* tp_hi, float* tp_lo, const float* taps, int K, int pl, int rev) {
  for (int i = threadIdx.x; i < 192; i += (int)blockDim.x) {
    const int j = i - 64;
    const float t = (j >= 0 && j < K) ? dpc_tap(taps, K, j, rev) : 0.0f;
    const float th = dpc_tc_hi(t);
    tp_hi[i] = th;
    tp_lo[i] = dpc_tc_hi(t - th);
  }
  __syncthreads();
  dpc_tc_build_toeplitz_from(tbase, tp_hi, tp_lo, pl);
}
// second half: the padded tap rows (hi / lo, 192 entries, tap j at 64 + j) -> the operand in the MMA's smem layout
DPC_DEV void dpc_tc_build_toeplitz_from(unsigned char* tbase, const float* tp_hi, const float* tp_lo, int pl) {
  if (threadIdx.x < 128) {
    const int n = threadIdx.x >> 1, h = threadIdx.x & 1;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int i0 = 32 * h + 4 * c - n + pl + 64;    // in [1, 189]
      const uint32_t off = (uint32_t)h * DPC_TC_T_HALF + (uint32_t)n * 128u + (uint32_t)((c ^ (n & 7)) << 4);
      *reinterpret_cast<float4*>(tbase + off) = make_float4(tp_hi[i0], tp_hi[i0 + 1], tp_hi[i0 + 2], tp_hi[i0 + 3]);
      *reinterpret_cast<float4*>(tbase + 2u * DPC_TC_T_HALF + off) = make_float4(tp_lo[i0], tp_lo[i0 + 1], tp_lo[i0 + 2], tp_lo[i0 + 3]);
    }
  }
}

// Common prologue (runs before griddepcontrol.wait, i.e. under the previous kernel's tail): TMEM
// allocation, mbarrier, Toeplitz operand.  Returns the TMEM base address.
DPC_DEV uint32_t dpc_tc_prologue(unsigned char* sm, uint64_t* bar, uint32_t* slot, float* tp_hi, float* tp_lo,
                                 const float* taps, int K, int pl, int rev, uint32_t ncols) {
  if ((threadIdx.x >> 5) == 0) dpc_tc_alloc(slot, ncols);
  if (threadIdx.x == 0) dpc_mbar_init(bar, 1);
  dpc_tc_build_toeplitz(sm + DPC_TC_T_HI, tp_hi, tp_lo, taps, K, pl, rev);
  dpc_fence_proxy_async();
  dpc_tc_fence_before();
  __syncthreads();
  dpc_tc_fence_after();
  return *slot;
}

DPC_DEV void dpc_tc_epilogue(uint32_t tmem, uint32_t ncols) {
  dpc_tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) dpc_tc_dealloc(tmem, ncols);
}

// operands written -> visible to the tensor core -> one thread issues the MMAs -> everyone waits for them
DPC_DEV void dpc_tc_run(uint32_t sbase, uint32_t d_tmem, uint64_t* bar, unsigned phase) {
  dpc_fence_proxy_async();
  dpc_tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) {
    if (dpc_elect_one()) {
      dpc_tc_fence_after();
      dpc_tc_issue(sbase, d_tmem, bar);
    }
    __syncwarp();
  }
  dpc_mbar_wait(bar, phase);
  dpc_tc_fence_after();
}

#define DPC_TC_SMEM_SETUP()                                                         \
  extern __shared__ __align__(16) unsigned char dpc_tc_dsm[];                       \
  __shared__ __align__(8) uint64_t bar;                                             \
  __shared__ uint32_t tmem_slot;                                                    \
  __shared__ __align__(16) float tp_hi[192];                                        \
  __shared__ __align__(16) float tp_lo[192];                                        \
  const uint32_t sraw = dpc_tc_s32(dpc_tc_dsm);                                     \
  const uint32_t sbase = (sraw + 1023u) & ~1023u;                                   \
  unsigned char* sm = dpc_tc_dsm + (sbase - sraw)

// ------------------------------------------------------------------------------ depth pass, forward
// CTA = 128 rays (image rows y0, y0+1) x 64 depth levels; thread = ray.  MODE = DPC_PROJ_*, HAS_S = occupancy
// scale + clip stage present: compile-time, so the 64 unrolled levels carry no option tests.
template <int MODE, bool HAS_S>
__global__ void __launch_bounds__(DPC_TC_THREADS, 2) dpc_tc_conv_z_fwd_kernel(const __grid_constant__ DpcConvZArgs a) {
  constexpr int V = 64, Vz = 64;
  constexpr bool CLAMPU = (MODE == DPC_PROJ_DRC);
  DPC_TC_SMEM_SETUP();
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t tmem = dpc_tc_prologue(sm, &bar, &tmem_slot, tp_hi, tp_lo, a.taps, a.K, a.pl, a.rev, 64);
  dpc_grid_dep_sync();
  const int b = blockIdx.y, y = blockIdx.x * 2 + (tid >> 6), x = tid & 63;
  const size_t ray = ((size_t)b * Vz * V + y) * V + x;       // voxel index of level 0 of this ray
  {
    const float* src = a.in + ray;
    float v[Vz];
#pragma unroll
    for (int z = 0; z < Vz; ++z) v[z] = src[(size_t)z * V * V];
#pragma unroll
    for (int q = 0; q < 16; ++q) dpc_tc_store4(sm, tid, q >> 3, q & 7, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  const float s = HAS_S ? a.scale[b] : 1.0f;
  const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
  dpc_tc_run(sbase, tmem, &bar, 0);
  float r[Vz];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  dpc_tc_ld32(taddr, r);
  dpc_tc_ld32(taddr + 32, r + 32);
  dpc_tc_wait_ld();
  // scale, clip (+ pass bits), voxels out, ray scan in four depth quarters (short dependency chains;
  // same association as the FFMA2 kernel: proj = S0 + T0 (S1 + T1 (S2 + T2 S3)))
  float T[4] = {1.f, 1.f, 1.f, 1.f}, S[4] = {0.f, 0.f, 0.f, 0.f}, mx = -INFINITY;
  uint32_t m0 = 0u, m1 = 0u;
  float* vout = a.vox_out + ray;
#pragma unroll
  for (int z = 0; z < Vz; ++z) {
    float v = r[z];
    if (HAS_S) {
      const float t = __fmul_rn(v, s);
      v = __saturatef(t);                        // clip to [0,1]; the clip passes the gradient iff it left t unchanged
      if (v == t) { if (z < 32) m0 |= 1u << z; else m1 |= 1u << (z - 32); }
    }
    vout[(size_t)z * V * V] = v;
    if (MODE == DPC_PROJ_MAX) {
      mx = fmaxf(mx, v);
    } else if (MODE != DPC_PROJ_NONE) {
      const float u = CLAMPU ? fminf(fmaxf(v, D.lo), D.hi) : v;
      float p = u * T[z >> 4];
      T[z >> 4] -= p;
      if (z == 0) p *= D.c0;
      S[z >> 4] += p;
    }
  }
  if (HAS_S && a.mask2_out) *reinterpret_cast<uint2*>(a.mask2_out + (((size_t)b * V + y) * V + x) * 2) = make_uint2(m0, m1);
  if (MODE != DPC_PROJ_NONE) {
    float out = mx;
    if (MODE != DPC_PROJ_MAX) out = fmaf(T[0], fmaf(T[1], fmaf(T[2], S[3], S[2]), S[1]), S[0]);
    const int yo = a.flip_y ? (V - 1 - y) : y;
    a.proj[((size_t)b * V + yo) * V + x] = out;
  }
  dpc_tc_epilogue(tmem, 64);
}

// ------------------------------------------------------------------------------ depth pass, backward (lean)
// Training configuration: DRC silhouette gradient only, occupancy scale present (see
// dpc_conv_z64_bwd_kernel for the quotient form of the DRC gradient).  Thread = ray: the gradient of all
// 64 levels is formed in registers, becomes the A operand, and the transposed depth correlation is one GEMM.
__global__ void __launch_bounds__(DPC_TC_THREADS, 2) dpc_tc_conv_z_bwd_lean_kernel(const __grid_constant__ DpcConvZBwdArgs a) {
  constexpr int V = 64, Vz = 64;
  DPC_TC_SMEM_SETUP();
  __shared__ float red[DPC_TC_THREADS / 32];
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t tmem = dpc_tc_prologue(sm, &bar, &tmem_slot, tp_hi, tp_lo, a.taps, a.K, a.pl, a.rev, 64);
  dpc_grid_dep_sync();
  const int b = blockIdx.y, y = blockIdx.x * 2 + (tid >> 6), x = tid & 63;
  const size_t ray = ((size_t)b * Vz * V + y) * V + x;
  const float s = a.scale[b];
  const float inv_s = (s != 0.0f) ? 1.0f / s : 0.0f;
  const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
  float ds;
  {
    const float* src = a.vox + ray;
    float v[Vz];
#pragma unroll
    for (int z = 0; z < Vz; ++z) v[z] = src[(size_t)z * V * V];
    const int yo = a.flip_y ? (V - 1 - y) : y;
    const float gp = a.g_proj[((size_t)b * V + yo) * V + x];
    const uint2 mw = *reinterpret_cast<const uint2*>(a.mask2 + (((size_t)b * V + y) * V + x) * 2);
    float T0 = 1.0f, T1 = 1.0f, T2 = 1.0f, T3 = 1.0f;
#pragma unroll
    for (int z = 0; z < Vz; z += 4) {
      T0 *= 1.0f - fminf(fmaxf(v[z + 0], D.lo), D.hi);
      T1 *= 1.0f - fminf(fmaxf(v[z + 1], D.lo), D.hi);
      T2 *= 1.0f - fminf(fmaxf(v[z + 2], D.lo), D.hi);
      T3 *= 1.0f - fminf(fmaxf(v[z + 3], D.lo), D.hi);
    }
    const float gT = gp * ((T0 * T1) * (T2 * T3));
    float dsv = 0.0f;
#pragma unroll
    for (int z = 0; z < Vz; ++z) {
      const float vv = v[z];
      const float u = fminf(fmaxf(vv, D.lo), D.hi);
      float dv = __fdividef(gT, 1.0f - u);
      if (z == 0) dv = fmaf(gp, D.c0 - 1.0f, dv);
      const uint32_t bit = ((z < 32 ? mw.x : mw.y) >> (z & 31)) & 1u;
      if ((u != vv) || !bit) dv = 0.0f;
      dsv = fmaf(dv, vv, dsv);
      v[z] = dv * s;
    }
    ds = dsv * inv_s;
#pragma unroll
    for (int q = 0; q < 16; ++q) dpc_tc_store4(sm, tid, q >> 3, q & 7, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  dpc_tc_run(sbase, tmem, &bar, 0);
  float r[Vz];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  dpc_tc_ld32(taddr, r);
  dpc_tc_ld32(taddr + 32, r + 32);
  dpc_tc_wait_ld();
  float* dout = a.d_in + ray;
#pragma unroll
  for (int z = 0; z < Vz; ++z) dout[(size_t)z * V * V] = r[z];
  if (a.d_scale) {
    const float w = dpc_warp_sum(ds);
    if ((tid & 31) == 0) red[warp] = w;
  }
  dpc_tc_epilogue(tmem, 64);      // contains the __syncthreads that publishes red[]
  if (a.d_scale && tid == 0) atomicAdd(a.d_scale + b, (red[0] + red[1]) + (red[2] + red[3]));
}

// ------------------------------------------------------------------------------ x and y passes
// CTA = two depth slices (128 rows of 64 x).  [clip, pass bits] -> x pass -> y pass -> [* saved mask].
// Both passes use the same Toeplitz operand (taps_x == taps_y, as smoothing_kernel builds them).
template <bool CLIP, bool MOUT, bool MIN>
__global__ void __launch_bounds__(DPC_TC_THREADS, 2) dpc_tc_conv_xy_kernel(const __grid_constant__ DpcConvXY64Args a, int K, int pl) {
  constexpr int V = 64;
  DPC_TC_SMEM_SETUP();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t tmem = dpc_tc_prologue(sm, &bar, &tmem_slot, tp_hi, tp_lo, a.taps_x, K, pl, a.rev, 128);
  dpc_grid_dep_sync();
  const size_t base = (size_t)blockIdx.x * (2 * V * V);     // first voxel of this CTA's two slices
  // saved clip mask of the OUTPUT voxels this thread will write (backward): thread = (slice, x), lane = x & 31,
  // word of row y = (base >> 5) + slice*128 + 2y + (x >> 5); lane l fetches rows l and l + 32.
  uint32_t mw0 = 0xffffffffu, mw1 = 0xffffffffu;
  if (MIN) {
    const uint32_t* mrow = a.mask_in + (base >> 5) + (size_t)(tid >> 6) * 128 + (warp & 1);
    mw0 = mrow[2 * lane];
    mw1 = mrow[2 * (lane + 32)];
  }
  // ---- two slices -> A1[(slice, y)][x]: coalesced float4 loads, clip, pass bits, split
  {
    const float4* src = reinterpret_cast<const float4*>(a.in + base);
    float4 v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = src[tid + DPC_TC_THREADS * k];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int i = tid + DPC_TC_THREADS * k;       // float4 index within the two slices
      if (MOUT) {
        unsigned nib = ((v[k].x >= 0.0f && v[k].x <= 1.0f) ? 1u : 0u) | ((v[k].y >= 0.0f && v[k].y <= 1.0f) ? 2u : 0u) |
                       ((v[k].z >= 0.0f && v[k].z <= 1.0f) ? 4u : 0u) | ((v[k].w >= 0.0f && v[k].w <= 1.0f) ? 8u : 0u);
        unsigned word = nib << (4 * (tid & 7));
        word |= __shfl_xor_sync(DPC_FULL, word, 1);
        word |= __shfl_xor_sync(DPC_FULL, word, 2);
        word |= __shfl_xor_sync(DPC_FULL, word, 4);
        if ((tid & 7) == 0) a.mask_out[(base >> 5) + (i >> 3)] = word;
      }
      if (CLIP) { v[k].x = dpc_clip01(v[k].x); v[k].y = dpc_clip01(v[k].y); v[k].z = dpc_clip01(v[k].z); v[k].w = dpc_clip01(v[k].w); }
      const int row = i >> 4, c16 = i & 15;         // 16 float4 per row of 64
      dpc_tc_store4(sm, row, c16 >> 3, c16 & 7, v[k].x, v[k].y, v[k].z, v[k].w);
    }
  }
  dpc_tc_run(sbase, tmem, &bar, 0);
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  // ---- x-smoothed row (slice, y' = tid & 63) -> column y' of A2[(slice, x)][y']
  {
    float r[V];
    dpc_tc_ld32(taddr, r);
    dpc_tc_ld32(taddr + 32, r + 32);
    dpc_tc_wait_ld();
    const int yp = tid & 63, sl = tid >> 6;
    const uint32_t colb = (uint32_t)(yp >> 5) * DPC_TC_A_HALF + (uint32_t)(sl * 64) * 128u + (uint32_t)((yp & 3) * 4);
    const int c = (yp & 31) >> 2;
#pragma unroll
    for (int x = 0; x < V; ++x) {
      const float hi = dpc_tc_hi(r[x]), lo = r[x] - hi;
      const uint32_t off = colb + (uint32_t)x * 128u + (uint32_t)((c ^ (x & 7)) << 4);
      *reinterpret_cast<float*>(sm + off) = hi;
      *reinterpret_cast<float*>(sm + DPC_TC_A_LO + off) = lo;
    }
  }
  dpc_tc_run(sbase, tmem + 64, &bar, 1);
  // ---- thread = (slice, x): 64 output rows y, lanes = consecutive x
  {
    float r[V];
    dpc_tc_ld32(taddr + 64, r);
    dpc_tc_ld32(taddr + 96, r + 32);
    dpc_tc_wait_ld();
    float* dst = a.out + base + (size_t)(tid >> 6) * V * V + (tid & 63);
#pragma unroll
    for (int y = 0; y < V; ++y) {
      float v = r[y];
      if (MIN) {
        const uint32_t w = __shfl_sync(DPC_FULL, (y < 32) ? mw0 : mw1, y & 31);
        if (!((w >> lane) & 1u)) v = 0.0f;
      }
      dst[y * V] = v;
    }
  }
  dpc_tc_epilogue(tmem, 128);
}

#include "dpc_smooth_tcp.cuh"

// ------------------------------------------------------------------------------ dispatch
// ---- 2-D tensor map of a [B,64,64,64] grid for the depth pass: inner dimension = the 4096 (y, x) positions of a
// level, outer = the B*64 (b, level) rows; box = {128 positions, 64 levels} = one depth-pass tile.  The encoder is a
// driver entry point fetched through the runtime (no link against libcuda).
typedef CUresult (*DpcEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DpcEncodeTiledFn dpc_tc_encoder() {
  static DpcEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<DpcEncodeTiledFn>(p);
  }
  return fn;
}
// effective kernel family: the pipelines need the tensor-map encoder of the driver; without it the single-tile kernels run
static int dpc_tc_level() {
  if (dpc_tc_enable >= 2) return dpc_tc_encoder() ? 2 : 1;
  return dpc_tc_enable > 0 ? 1 : 0;
}
static int dpc_tc_make_zmap(CUtensorMap* m, const float* grid, int B) {
  DpcEncodeTiledFn fn = dpc_tc_encoder();
  if (!fn) return DPC_ERR_CUDA;
  const cuuint64_t dims[2] = {4096, (cuuint64_t)B * 64};
  const cuuint64_t strides[1] = {4096 * sizeof(float)};
  const cuuint32_t box[2] = {128, 64};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(grid), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DPC_OK : DPC_ERR_CUDA;
}
// 2-D view of nslices depth slices for the x / y passes: [nslices*64 rows (slice, y)][64 x], box {32 x, 128 rows},
// 128-byte swizzle (16-byte chunk c of row r lands at chunk c ^ (r & 7))
static int dpc_tc_make_xymap(CUtensorMap* m, const float* grid, int64_t nslices) {
  DpcEncodeTiledFn fn = dpc_tc_encoder();
  if (!fn) return DPC_ERR_CUDA;
  const cuuint64_t dims[2] = {64, (cuuint64_t)nslices * 64};
  const cuuint64_t strides[1] = {64 * sizeof(float)};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(grid), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DPC_OK : DPC_ERR_CUDA;
}
static int dpc_tc_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
    else n = 148;
  }
  return n;
}
// Host copy of the taps for the NEXT pipeline launch on this thread (set by the C-ABI layer right before it calls a
// launcher below, consumed and cleared there); NULL = the kernel reads the device taps.
static thread_local const float* dpc_tcp_host_taps_next = nullptr;
static DPC_KNOB_T dpc_tcp_ns3 = 0;         // experiment knob 2 = 2: 3-slot staging ring in the x/y pipeline (what a co-resident CTA would need)
static DPC_KNOB_T dpc_tcp_pdrain = 0;      // experiment knob 2: the producer warps of the x/y pipeline store the tiles
static inline DpcTcpTaps dpc_tcp_take_host_taps(int K) {
  DpcTcpTaps ht;
  const float* h = dpc_tcp_host_taps_next;
  dpc_tcp_host_taps_next = nullptr;
  ht.valid = (h != nullptr && K >= 1 && K <= DPC_MAX_TAPS && !dpc_ignore_host_taps) ? 1 : 0;
  for (int i = 0; i <= DPC_MAX_TAPS; ++i) ht.t[i] = (ht.valid && i < K) ? h[i] : 0.0f;
  return ht;
}

static inline bool dpc_tc_conv_xy_supported(int V, int Kx, int plx, int Ky, int ply, const float* tx, const float* ty,
                                            int64_t nslices, const float* zero_ptr) {
  return dpc_tc_level() && V == 64 && Kx == Ky && plx == ply && tx == ty && Kx >= 1 && Kx <= DPC_MAX_TAPS &&
         (nslices % 2) == 0 && !zero_ptr;
}
static inline int dpc_tc_conv_xy_launch(const float* in, float* out, const float* taps, int K, int pl, int64_t nslices,
                                        int clip_in, uint32_t* mask_out, const uint32_t* mask_in, int rev, void* stream) {
  if ((((uintptr_t)in) & 15u) || (((uintptr_t)out) & 3u)) return DPC_ERR_ARG;
  DpcConvXY64Args a;
  a.in = in; a.out = out; a.taps_x = taps; a.taps_y = taps; a.clip_in = clip_in; a.mask_out = mask_out; a.mask_in = mask_in;
  a.nslices = (int)nslices; a.rev = rev; a.zero_ptr = nullptr; a.dbg = 0;
  for (int i = 0; i < 24; ++i) { a.ht.px[i] = make_float2(0.f, 0.f); a.ht.dy[i] = make_float2(0.f, 0.f); }
#define DPC_TC_XY_GO(C, MO, MI) do { \
    if (cudaFuncSetAttribute(dpc_tc_conv_xy_kernel<C, MO, MI>, cudaFuncAttributeMaxDynamicSharedMemorySize, DPC_TC_SMEM_BYTES) != cudaSuccess) \
      return DPC_ERR_CUDA; \
    DPC_LAUNCH((dpc_tc_conv_xy_kernel<C, MO, MI>), dim3((unsigned)(nslices / 2)), dim3(DPC_TC_THREADS), DPC_TC_SMEM_BYTES, stream, a, K, pl); \
    return DPC_OK; } while (0)
  const int sel = (clip_in ? 4 : 0) | (mask_out ? 2 : 0) | (mask_in ? 1 : 0);
  if (dpc_tc_level() == 2) {
    const int ntiles = (int)(nslices / 2), grid = ntiles < dpc_tc_sm_count() ? ntiles : dpc_tc_sm_count();
    CUtensorMap xymap;
    if (dpc_tc_make_xymap(&xymap, in, nslices) != DPC_OK) return DPC_ERR_CUDA;
    const DpcTcpTaps ht = dpc_tcp_take_host_taps(taps ? K : 0);
    const int xy_ns = dpc_tcp_ns3 ? 3 : DPC_TCP_NS;
#define DPC_TCP_XY_GO1(C, MO, MI, PD) do { \
    if (cudaFuncSetAttribute(dpc_tcp_conv_xy_kernel<C, MO, MI, PD>, cudaFuncAttributeMaxDynamicSharedMemorySize, DPC_TCP_SMEM_BYTES) != cudaSuccess) \
      return DPC_ERR_CUDA; \
    DPC_LAUNCH((dpc_tcp_conv_xy_kernel<C, MO, MI, PD>), dim3(grid), dim3(DPC_TCP_THREADS), (size_t)(98304 + xy_ns * 32768), stream, a, xymap, K, pl, ntiles, ht, xy_ns); \
    return DPC_OK; } while (0)
#ifdef DPC_EXPERIMENTS
#define DPC_TCP_XY_GO(C, MO, MI) do { if (dpc_tcp_pdrain) DPC_TCP_XY_GO1(C, MO, MI, true); else DPC_TCP_XY_GO1(C, MO, MI, false); } while (0)
#else
#define DPC_TCP_XY_GO(C, MO, MI) DPC_TCP_XY_GO1(C, MO, MI, false)
#endif
    switch (sel) {
      case 0: DPC_TCP_XY_GO(false, false, false);
      case 1: DPC_TCP_XY_GO(false, false, true);
      case 2: DPC_TCP_XY_GO(false, true, false);
      case 3: DPC_TCP_XY_GO(false, true, true);
      case 4: DPC_TCP_XY_GO(true, false, false);
      case 5: DPC_TCP_XY_GO(true, false, true);
      case 6: DPC_TCP_XY_GO(true, true, false);
      default: DPC_TCP_XY_GO(true, true, true);
    }
#undef DPC_TCP_XY_GO
#undef DPC_TCP_XY_GO1
  }
  switch (sel) {
    case 0: DPC_TC_XY_GO(false, false, false);
    case 1: DPC_TC_XY_GO(false, false, true);
    case 2: DPC_TC_XY_GO(false, true, false);
    case 3: DPC_TC_XY_GO(false, true, true);
    case 4: DPC_TC_XY_GO(true, false, false);
    case 5: DPC_TC_XY_GO(true, false, true);
    case 6: DPC_TC_XY_GO(true, true, false);
    default: DPC_TC_XY_GO(true, true, true);
  }
#undef DPC_TC_XY_GO
}

static inline bool dpc_tc_conv_z_supported(int V, int Vz, int Kz, bool extras) {
  return dpc_tc_level() && V == 64 && Vz == 64 && Kz >= 1 && Kz <= DPC_MAX_TAPS && !extras;
}
template <int MODE, bool HAS_S>
static inline int dpc_tc_conv_z_fwd_go(const DpcConvZArgs& a, void* stream) {
  if (cudaFuncSetAttribute(dpc_tc_conv_z_fwd_kernel<MODE, HAS_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, DPC_TC_SMEM_BYTES) != cudaSuccess)
    return DPC_ERR_CUDA;
  DPC_LAUNCH((dpc_tc_conv_z_fwd_kernel<MODE, HAS_S>), dim3(32, a.B), dim3(DPC_TC_THREADS), DPC_TC_SMEM_BYTES, stream, a);
  return DPC_OK;
}
template <int MODE, bool HAS_S>
static inline int dpc_tcp_conv_z_fwd_go(const DpcConvZArgs& a, void* stream) {
  if ((((uintptr_t)a.in) & 15u) != 0) return DPC_ERR_ARG;
  CUtensorMap zmap;
  if (dpc_tc_make_zmap(&zmap, a.in, a.B) != DPC_OK) return DPC_ERR_CUDA;
  if (cudaFuncSetAttribute(dpc_tcp_conv_z_fwd_kernel<MODE, HAS_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, DPC_TCP_SMEM_BYTES) != cudaSuccess)
    return DPC_ERR_CUDA;
  const int ntiles = 32 * a.B, grid = ntiles < dpc_tc_sm_count() ? ntiles : dpc_tc_sm_count();
  const DpcTcpTaps ht = dpc_tcp_take_host_taps(a.taps ? a.K : 0);
  DPC_LAUNCH((dpc_tcp_conv_z_fwd_kernel<MODE, HAS_S>), dim3(grid), dim3(DPC_TCP_THREADS), DPC_TCP_SMEM_BYTES, stream, a, zmap, ntiles, ht);
  return DPC_OK;
}
static inline int dpc_tc_conv_z_fwd_launch(const DpcConvZArgs& a, void* stream) {
  const bool hs = a.scale != nullptr;
  if (dpc_tc_level() == 2) {
    switch (a.mode) {
      case DPC_PROJ_DRC: return hs ? dpc_tcp_conv_z_fwd_go<DPC_PROJ_DRC, true>(a, stream) : dpc_tcp_conv_z_fwd_go<DPC_PROJ_DRC, false>(a, stream);
      case DPC_PROJ_MAX: return hs ? dpc_tcp_conv_z_fwd_go<DPC_PROJ_MAX, true>(a, stream) : dpc_tcp_conv_z_fwd_go<DPC_PROJ_MAX, false>(a, stream);
      case DPC_PROJ_DRC_PROD: return hs ? dpc_tcp_conv_z_fwd_go<DPC_PROJ_DRC_PROD, true>(a, stream) : dpc_tcp_conv_z_fwd_go<DPC_PROJ_DRC_PROD, false>(a, stream);
      default: return hs ? dpc_tcp_conv_z_fwd_go<DPC_PROJ_NONE, true>(a, stream) : dpc_tcp_conv_z_fwd_go<DPC_PROJ_NONE, false>(a, stream);
    }
  }
  switch (a.mode) {
    case DPC_PROJ_DRC: return hs ? dpc_tc_conv_z_fwd_go<DPC_PROJ_DRC, true>(a, stream) : dpc_tc_conv_z_fwd_go<DPC_PROJ_DRC, false>(a, stream);
    case DPC_PROJ_MAX: return hs ? dpc_tc_conv_z_fwd_go<DPC_PROJ_MAX, true>(a, stream) : dpc_tc_conv_z_fwd_go<DPC_PROJ_MAX, false>(a, stream);
    case DPC_PROJ_DRC_PROD: return hs ? dpc_tc_conv_z_fwd_go<DPC_PROJ_DRC_PROD, true>(a, stream) : dpc_tc_conv_z_fwd_go<DPC_PROJ_DRC_PROD, false>(a, stream);
    default: return hs ? dpc_tc_conv_z_fwd_go<DPC_PROJ_NONE, true>(a, stream) : dpc_tc_conv_z_fwd_go<DPC_PROJ_NONE, false>(a, stream);
  }
}
static inline int dpc_tc_conv_z_bwd_lean_launch(const DpcConvZBwdArgs& a, void* stream) {
  if (dpc_tc_level() == 2) {
    if (cudaFuncSetAttribute(dpc_tcp_conv_z_bwd_lean_kernel<DPC_PROJ_DRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, DPC_TCP_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(dpc_tcp_conv_z_bwd_lean_kernel<DPC_PROJ_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, DPC_TCP_SMEM_BYTES) != cudaSuccess)
      return DPC_ERR_CUDA;
    if ((((uintptr_t)a.vox) & 15u) != 0) return DPC_ERR_ARG;
    CUtensorMap zmap;
    if (dpc_tc_make_zmap(&zmap, a.vox, a.B) != DPC_OK) return DPC_ERR_CUDA;
    const int ntiles = 32 * a.B, grid = ntiles < dpc_tc_sm_count() ? ntiles : dpc_tc_sm_count();
    const DpcTcpTaps ht = dpc_tcp_take_host_taps(a.taps ? a.K : 0);
    if (a.mode == DPC_PROJ_MAX) { DPC_LAUNCH(dpc_tcp_conv_z_bwd_lean_kernel<DPC_PROJ_MAX>, dim3(grid), dim3(DPC_TCP_THREADS), DPC_TCP_SMEM_BYTES, stream, a, zmap, ntiles, ht); }
    else { DPC_LAUNCH(dpc_tcp_conv_z_bwd_lean_kernel<DPC_PROJ_DRC>, dim3(grid), dim3(DPC_TCP_THREADS), DPC_TCP_SMEM_BYTES, stream, a, zmap, ntiles, ht); }
    return DPC_OK;
  }
  if (cudaFuncSetAttribute(dpc_tc_conv_z_bwd_lean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DPC_TC_SMEM_BYTES) != cudaSuccess)
    return DPC_ERR_CUDA;
  DPC_LAUNCH(dpc_tc_conv_z_bwd_lean_kernel, dim3(32, a.B), dim3(DPC_TC_THREADS), DPC_TC_SMEM_BYTES, stream, a);
  return DPC_OK;
}

#endif  // !DPC_EMU
