// Shape-specialised smoothing / projection kernels for the 64^3 grid (the benchmark and both
// experiment configs: vox_size 64, pc_gauss_kernel_size 21 or 11).
//
// At K = 21 the separable Gaussian costs 63 FMA per voxel per direction -- 1.06 G FMA per B=32
// forward+backward -- which makes these kernels FP32-pipe-bound before they are HBM-bound.  They
// are therefore built around Blackwell's packed FFMA2 (fma.rn.f32x2: two fp32 FMAs per issued
// instruction):
//   * along x, two adjacent INPUTS are paired with two adjacent TAPS:  acc2 += (in[m],in[m+1]) *
//     (t[a],t[a+1]); the output is acc2.lo + acc2.hi.  Row-major smem, aligned pairs, no shuffles.
//   * along y and along depth, two adjacent-x OUTPUTS share a tap:  acc2 += (in[.][x],in[.][x+1])
//     * (t,t), 8 outputs per thread held in registers, each input row read once.
// Every loop is fully unrolled so all register arrays are statically indexed.
#pragma once
#include "dpc_common.cuh"
#include "dpc_smooth.cuh"

#define DPC_F64_V 64
#define DPC_F64_S 68          // smem row stride in floats: 68/4 = 17 (odd) -> 8 consecutive rows hit 8 distinct 16-byte bank groups

DPC_DEV float2 dpc_f2(float a, float b) { return make_float2(a, b); }

// acc[o] += sum_j tt[j] * in[r0 + o + j - PL][pair], rows outside [0, nrows) read as zero.
// base points at the pair's element in row 0; stride in floats.
template <int K, int R>
DPC_DEV void dpc_col_conv_pairs(const float* base, int stride, int r0, int nrows, const float2* tt, float2* acc) {
  constexpr int PL = (K - 1) / 2;
#pragma unroll
  for (int i = 0; i < R + K - 1; ++i) {
    const int row = r0 - PL + i;
    float2 w = dpc_f2(0.0f, 0.0f);
    if (row >= 0 && row < nrows) w = *reinterpret_cast<const float2*>(base + (size_t)row * stride);
#pragma unroll
    for (int o = 0; o < R; ++o) {
      if (i - o >= 0 && i - o < K) acc[o] = dpc_ffma2(w, tt[i - o], acc[o]);
    }
  }
}

// ------------------------------------------------------------------------------ conv_xy, V = 64
struct DpcConvXY64Args {
  const float* in; float* out; const float* taps_x; const float* taps_y;
  int clip_in; uint32_t* mask_out; const uint32_t* mask_in;
};

template <int K>
#ifndef DPC_EMU
__global__ void __launch_bounds__(256, 2)
#else
static void
#endif
dpc_conv_xy64_kernel(DpcConvXY64Args a) {
  constexpr int V = DPC_F64_V, S = DPC_F64_S, PL = (K - 1) / 2;
  constexpr int WL = ((PL + 3) / 4) * 4;            // window starts WL floats left of the first output
  constexpr int NW4 = (WL + 16 + WL) / 4;           // float4 groups in the x window
  static_assert((K & 1) == 1 && K <= 21, "odd K <= 21");
  __shared__ __align__(16) float A[V * S];
  __shared__ __align__(16) float M[V * S];
  __shared__ float tx[K + 3], ty[K + 3];
  const int tid = threadIdx.x;
  const size_t slice = (size_t)blockIdx.x * (V * V);
  if (tid < K) { tx[tid + 1] = a.taps_x[tid]; ty[tid] = a.taps_y[tid]; }
  if (tid == 0) { tx[0] = 0.0f; tx[K + 1] = 0.0f; }

  // ---- phase 0: slice -> smem (float4, coalesced), clip, clip-mask bits
  {
    const float4* src = reinterpret_cast<const float4*>(a.in + slice);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = tid + 256 * k;            // float4 index in the slice: row = i/16, col4 = i%16
      float4 v = src[i];
      if (a.mask_out) {
        unsigned nib = ((v.x >= 0.0f && v.x <= 1.0f) ? 1u : 0u) | ((v.y >= 0.0f && v.y <= 1.0f) ? 2u : 0u) |
                       ((v.z >= 0.0f && v.z <= 1.0f) ? 4u : 0u) | ((v.w >= 0.0f && v.w <= 1.0f) ? 8u : 0u);
        unsigned word = nib << (4 * (tid & 7));
        word |= __shfl_xor_sync(DPC_FULL, word, 1);
        word |= __shfl_xor_sync(DPC_FULL, word, 2);
        word |= __shfl_xor_sync(DPC_FULL, word, 4);
        if ((tid & 7) == 0) a.mask_out[(slice >> 5) + (i >> 3)] = word;
      }
      if (a.clip_in) { v.x = dpc_clip01(v.x); v.y = dpc_clip01(v.y); v.z = dpc_clip01(v.z); v.w = dpc_clip01(v.w); }
      *reinterpret_cast<float4*>(&A[(i >> 4) * S + (i & 15) * 4]) = v;
    }
  }
  __syncthreads();

  // ---- phase 1: x correlation.  Thread = (row y, run r of 16 outputs); a warp = 32 rows, one r.
  {
    const int y = tid & 63, r = tid >> 6;
    const int x0 = r * 16;
    float2 tp[K + 1];                      // tp[a+1] = (t[a], t[a+1]), a = -1..K-1, zero outside
#pragma unroll
    for (int q = 0; q < K + 1; ++q) tp[q] = dpc_f2(tx[q], tx[q + 1]);
    float2 acc[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = dpc_f2(0.0f, 0.0f);
    const float* rowp = A + y * S;
#pragma unroll
    for (int g = 0; g < NW4; ++g) {
      const int xs = x0 - WL + 4 * g;       // warp-uniform: whole float4 in or out of the row
      float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (xs >= 0 && xs < V) w4 = *reinterpret_cast<const float4*>(rowp + xs);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 w = h ? dpc_f2(w4.z, w4.w) : dpc_f2(w4.x, w4.y);
        // window pair index i = 4g + 2h; for output o the first tap of the pair is a = i - o - WL + PL
#pragma unroll
        for (int o = 0; o < 16; ++o) {
          if (4 * g + 2 * h - o - WL + PL >= -1 && 4 * g + 2 * h - o - WL + PL <= K - 1)
            acc[o] = dpc_ffma2(w, tp[4 * g + 2 * h - o - WL + PL + 1], acc[o]);
        }
      }
    }
    float* dst = M + y * S + x0;
#pragma unroll
    for (int o = 0; o < 16; o += 4) {
      *reinterpret_cast<float4*>(dst + o) = make_float4(acc[o].x + acc[o].y, acc[o + 1].x + acc[o + 1].y,
                                                         acc[o + 2].x + acc[o + 2].y, acc[o + 3].x + acc[o + 3].y);
    }
  }
  __syncthreads();

  // ---- phase 2: y correlation.  Thread = (x pair, run of 8 rows); a warp = one run, 32 x pairs.
  {
    const int xp = tid & 31, y0 = (tid >> 5) * 8;
    float2 tt[K];
#pragma unroll
    for (int j = 0; j < K; ++j) tt[j] = dpc_f2(ty[j], ty[j]);
    float2 acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = dpc_f2(0.0f, 0.0f);
    dpc_col_conv_pairs<K, 8>(M + 2 * xp, S, y0, V, tt, acc);
    float* dst = a.out + slice + (size_t)y0 * V + 2 * xp;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float2 v = acc[o];
      if (a.mask_in) {
        const size_t e = slice + (size_t)(y0 + o) * V + 2 * xp;
        const uint32_t wbits = a.mask_in[e >> 5] >> (e & 31);
        if (!(wbits & 1u)) v.x = 0.0f;
        if (!(wbits & 2u)) v.y = 0.0f;
      }
      *reinterpret_cast<float2*>(dst + (size_t)o * V) = v;
    }
  }
}

// ------------------------------------------------------------------------------ conv_z, V = Vz = 64
// CTA = 4 image rows x all 64 depth levels (64 KiB tile, TMA bulk loads), 256 threads = 8 warps.
// Warp w works on image row w>>1 and on depth half w&1 (levels 32h .. 32h+31): the correlation has
// no carry along depth, and the ray scan is affine in its carry (p_i = u_i T_i with T a running
// product), so each half scans from T = 1 and the two halves are combined through smem:
//   proj = S_lo + T_lo * S_hi,   max = max(max_lo, max_hi).
// That doubles the resident warps per tile (24 per SM) against one warp per row.
// Not handled here (the generic kernel takes over): drc_probs / proj_depth outputs or gradients.
#define DPC_ZF_TY 4
#define DPC_ZF_THREADS 256

template <int K>
#ifndef DPC_EMU
__global__ void __launch_bounds__(DPC_ZF_THREADS)
#else
static void
#endif
dpc_conv_z64_fwd_kernel(DpcConvZArgs a) {
  constexpr int V = DPC_F64_V, Vz = DPC_F64_V, TY = DPC_ZF_TY, RW = TY * V;
  DPC_DYN_SMEM(float, tile);                // [Vz][TY][V]
  __shared__ __align__(8) uint64_t bar;
  __shared__ float tz[K];
  __shared__ __align__(8) float comb[TY][V][2];   // (T, S) or (max, -) of the low depth half per ray
  const int tid = threadIdx.x;
  const int b = blockIdx.y, y0 = blockIdx.x * TY;
  if (tid < K) tz[tid] = a.taps[tid];
  // ---- tile load: 64 bulk copies (one per depth level, TY*V*4 = 1 KiB each) through the TMA
  // engine, completion on one mbarrier; no register staging.
  const float* src = a.in + ((size_t)b * Vz * V + y0) * V;
  if (tid == 0) dpc_mbar_init(&bar, 1);
  __syncthreads();
  if (tid == 0) {
#ifndef DPC_EMU
    unsigned bb = (unsigned)__cvta_generic_to_shared(&bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"((unsigned)(Vz * RW * 4)) : "memory");
    for (int z = 0; z < Vz; ++z) {
      unsigned d = (unsigned)__cvta_generic_to_shared(tile + z * RW);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(d), "l"(src + (size_t)z * V * V), "r"((unsigned)(RW * 4)), "r"(bb) : "memory");
    }
#else
    for (int z = 0; z < Vz; ++z) memcpy(tile + z * RW, src + (size_t)z * V * V, RW * 4);
#endif
  }
  dpc_mbar_wait(&bar, 0);
  __syncthreads();

  const int w = tid >> 5, xp = tid & 31;
  const int ty = w >> 1, h = w & 1, y = y0 + ty;
  float2 tt[K];
#pragma unroll
  for (int j = 0; j < K; ++j) tt[j] = dpc_f2(tz[j], tz[j]);
  const bool has_s = a.scale != nullptr;
  const float s = has_s ? a.scale[b] : 1.0f;
  const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
  float2 T = dpc_f2(1.f, 1.f), S = dpc_f2(0.f, 0.f), mx = dpc_f2(-INFINITY, -INFINITY);
  uint32_t m0 = 0u, m1 = 0u;   // clip-pass bits of the two rays for this depth half
  float* vout = a.vox_out + ((size_t)b * Vz * V + y) * V + 2 * xp;
  const float* col = tile + ty * V + 2 * xp;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const int zc = (4 * h + c) * 8;
    float2 acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = dpc_f2(0.0f, 0.0f);
    dpc_col_conv_pairs<K, 8>(col, RW, zc, Vz, tt, acc);
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float2 v = acc[o];
      if (has_s) {
        const float t0 = __fmul_rn(v.x, s), t1 = __fmul_rn(v.y, s);
        v = dpc_f2(dpc_clip01(t0), dpc_clip01(t1));
        // the clip passes the gradient iff 0 <= t <= 1, i.e. iff it left t unchanged
        if (v.x == t0) m0 |= 1u << (c * 8 + o);
        if (v.y == t1) m1 |= 1u << (c * 8 + o);
      }
      *reinterpret_cast<float2*>(vout + (size_t)(zc + o) * V * V) = v;
      if (a.mode == DPC_PROJ_MAX) {
        mx = dpc_f2(fmaxf(mx.x, v.x), fmaxf(mx.y, v.y));
      } else if (a.mode != DPC_PROJ_NONE) {
        const float u0 = D.clampu ? fminf(fmaxf(v.x, D.lo), D.hi) : v.x;
        const float u1 = D.clampu ? fminf(fmaxf(v.y, D.lo), D.hi) : v.y;
        float p0 = u0 * T.x, p1 = u1 * T.y;           // p_i = u_i T_i (local T)
        T = dpc_f2(T.x - p0, T.y - p1);                // T_{i+1} = T_i (1 - u_i)
        if (o == 0 && zc == 0) { p0 *= D.c0; p1 *= D.c0; }   // the reference's e^eps on the first event
        S = dpc_f2(S.x + p0, S.y + p1);
      }
    }
  }
  if (a.mask2_out && has_s) {
    uint32_t* mp = a.mask2_out + (((size_t)b * V + y) * V + 2 * xp) * 2 + h;
    mp[0] = m0; mp[2] = m1;
  }
  if (a.mode == DPC_PROJ_NONE) return;
  // combine the two depth halves
  if (h == 0) {
    *reinterpret_cast<float2*>(&comb[ty][2 * xp][0]) = (a.mode == DPC_PROJ_MAX) ? dpc_f2(mx.x, 0.f) : dpc_f2(T.x, S.x);
    *reinterpret_cast<float2*>(&comb[ty][2 * xp + 1][0]) = (a.mode == DPC_PROJ_MAX) ? dpc_f2(mx.y, 0.f) : dpc_f2(T.y, S.y);
  }
  __syncthreads();
  if (h == 1) {
    const float2 c0v = *reinterpret_cast<const float2*>(&comb[ty][2 * xp][0]);
    const float2 c1v = *reinterpret_cast<const float2*>(&comb[ty][2 * xp + 1][0]);
    float2 out;
    if (a.mode == DPC_PROJ_MAX) out = dpc_f2(fmaxf(c0v.x, mx.x), fmaxf(c1v.x, mx.y));
    else out = dpc_f2(fmaf(c0v.x, S.x, c0v.y), fmaf(c1v.x, S.y, c1v.y));
    const int yo = a.flip_y ? (V - 1 - y) : y;
    *reinterpret_cast<float2*>(a.proj + ((size_t)b * V + yo) * V + 2 * xp) = out;
  }
}

template <int K>
#ifndef DPC_EMU
__global__ void __launch_bounds__(DPC_ZF_THREADS)
#else
static void
#endif
dpc_conv_z64_bwd_kernel(DpcConvZBwdArgs a) {
  constexpr int V = DPC_F64_V, Vz = DPC_F64_V, TY = DPC_ZF_TY, RW = TY * V;
  DPC_DYN_SMEM(float, tile);                // [Vz][TY][V]: T_k, then dL/d(smoothed)
  __shared__ float tz[K];
  __shared__ float red[DPC_ZF_THREADS / 32];
  const int tid = threadIdx.x;
  const int b = blockIdx.y, y0 = blockIdx.x * TY;
  if (tid < K) tz[tid] = a.taps[tid];
  const bool has_s = a.scale != nullptr;
  const float s = has_s ? a.scale[b] : 1.0f;
  const float inv_s = (s != 0.0f) ? 1.0f / s : 0.0f;
  const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
  float ds = 0.0f;
  // ---- phase 1: one thread per ray.  dL/dvoxel from the projection, then back through
  // clip(. * scale); result (dL/d smoothed) left in the tile.
  {
    const int ty = tid >> 6, x = tid & 63, y = y0 + ty;
    const int yo = a.flip_y ? (V - 1 - y) : y;
    const float* vin = a.vox + ((size_t)b * Vz * V + y) * V + x;
    const float* gv = a.g_vox ? a.g_vox + ((size_t)b * Vz * V + y) * V + x : nullptr;
    float* col = tile + ty * V + x;
    const float gp = a.g_proj ? a.g_proj[((size_t)b * V + yo) * V + x] : 0.0f;
    uint32_t mlo = 0xffffffffu, mhi = 0xffffffffu;
    if (a.mask2 && has_s) {
      const uint2 mw = *reinterpret_cast<const uint2*>(a.mask2 + (((size_t)b * V + y) * V + x) * 2);
      mlo = mw.x; mhi = mw.y;
    }
    float share = 0.0f, mx = -INFINITY;
    if (a.mode == DPC_PROJ_MAX) {
      for (int z = 0; z < Vz; ++z) mx = fmaxf(mx, vin[(size_t)z * V * V]);
      int cnt = 0;
      for (int z = 0; z < Vz; ++z) cnt += (vin[(size_t)z * V * V] == mx) ? 1 : 0;
      share = gp / (float)cnt;     // TF _MaxGrad: ties share the gradient equally
    } else if (a.mode != DPC_PROJ_NONE) {
      // prefix products T_k = prod_{j<k} (1-u_j) into the tile
      float T = 1.0f;
#pragma unroll 8
      for (int z = 0; z < Vz; ++z) {
        const float v = vin[(size_t)z * V * V];
        const float u = D.clampu ? fminf(fmaxf(v, D.lo), D.hi) : v;
        col[z * RW] = T;
        T = T - u * T;
      }
    }
    // reverse sweep.  With only the silhouette gradient g:  dL/du_k = g T_k (c_k - 1 + R_k),
    // R_k = prod_{j>k} (1-u_j)  (for k > 0 that is g * prod_{j != k} (1-u_j)).
    float R = 1.0f;
#pragma unroll 8
    for (int z = Vz - 1; z >= 0; --z) {
      const float v = vin[(size_t)z * V * V];      // second read of the column: L1/L2 hit
      float dv = 0.0f;
      if (a.mode == DPC_PROJ_MAX) {
        dv = (v == mx) ? share : 0.0f;
      } else if (a.mode != DPC_PROJ_NONE) {
        const float u = D.clampu ? fminf(fmaxf(v, D.lo), D.hi) : v;
        const float ck = (z == 0) ? (D.c0 - 1.0f) : 0.0f;
        dv = gp * col[z * RW] * (ck + R);
        R = R - u * R;
        if (D.clampu && !(v >= D.lo && v <= D.hi)) dv = 0.0f;   // clip_by_value passes lo <= v <= hi
      }
      if (gv) dv += gv[(size_t)z * V * V];
      if (has_s) {
        const uint32_t wbits = (z < 32) ? mlo : mhi;
        if (!((wbits >> (z & 31)) & 1u)) dv = 0.0f;
        ds = fmaf(dv, v * inv_s, ds);      // smoothed value = voxels / scale where the clip passed
        dv *= s;
      }
      col[z * RW] = dv;
    }
  }
  __syncthreads();
  // ---- phase 2: transposed depth correlation (reversed taps), warp = (row, depth half)
  {
    const int w = tid >> 5, xp = tid & 31;
    const int ty = w >> 1, h = w & 1, y = y0 + ty;
    float2 tt[K];
#pragma unroll
    for (int j = 0; j < K; ++j) tt[j] = dpc_f2(tz[j], tz[j]);
    const float* col = tile + ty * V + 2 * xp;
    float* dout = a.d_in + ((size_t)b * Vz * V + y) * V + 2 * xp;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int zc = (4 * h + c) * 8;
      float2 acc[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = dpc_f2(0.0f, 0.0f);
      dpc_col_conv_pairs<K, 8>(col, RW, zc, Vz, tt, acc);
#pragma unroll
      for (int o = 0; o < 8; ++o) *reinterpret_cast<float2*>(dout + (size_t)(zc + o) * V * V) = acc[o];
    }
  }
  if (a.d_scale) {
    const float v = dpc_warp_sum(ds);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      float t = 0.0f;
      for (int i = 0; i < DPC_ZF_THREADS / 32; ++i) t += red[i];
      atomicAdd(a.d_scale + b, t);
    }
  }
}

// ------------------------------------------------------------------------------ dispatch
static inline bool dpc_fast_k(int K) { return K == 21 || K == 11; }

static inline bool dpc_conv_xy_fast_supported(int V, int Kx, int plx, int Ky, int ply) {
  return V == 64 && Kx == Ky && dpc_fast_k(Kx) && plx == (Kx - 1) / 2 && ply == (Ky - 1) / 2;
}

static inline int dpc_conv_xy_fast_launch(const float* in, float* out, const float* taps_x, const float* taps_y, int K,
                                          int B, int Vz, int V, int clip_in, uint32_t* mask_out, const uint32_t* mask_in,
                                          void* stream) {
  (void)V;
  if ((((uintptr_t)in) & 15u) || (((uintptr_t)out) & 7u)) return DPC_ERR_ARG;
  DpcConvXY64Args a;
  a.in = in; a.out = out; a.taps_x = taps_x; a.taps_y = taps_y; a.clip_in = clip_in; a.mask_out = mask_out; a.mask_in = mask_in;
  if (K == 21) { DPC_LAUNCH(dpc_conv_xy64_kernel<21>, dim3(B * Vz), dim3(256), 0, stream, a); }
  else { DPC_LAUNCH(dpc_conv_xy64_kernel<11>, dim3(B * Vz), dim3(256), 0, stream, a); }
  return DPC_OK;
}

// `extras`: drc_probs / proj_depth outputs (forward) or their gradients (backward) requested
static inline bool dpc_conv_z_fast_supported(int V, int Vz, int Kz, int plz, bool extras) {
  return V == 64 && Vz == 64 && dpc_fast_k(Kz) && plz == (Kz - 1) / 2 && !extras;
}

static inline int dpc_conv_z_fwd_fast_launch(const float* in, const float* taps_z, int Kz, const float* scale, int mode,
                                             float eps, float cam_dist, float max_depth, int flip_y, int B, int Vz, int V,
                                             float* vox_out, uint32_t* mask2_out, float* proj, float* probs, float* depth,
                                             void* stream) {
  DpcConvZArgs a;
  a.in = in; a.taps = taps_z; a.K = Kz; a.pl = (Kz - 1) / 2; a.scale = scale; a.mode = mode; a.eps = eps;
  a.cam_dist = cam_dist; a.max_depth = max_depth; a.flip_y = flip_y; a.B = B; a.Vz = Vz; a.V = V; a.TY = DPC_ZF_TY;
  a.vox_out = vox_out; a.mask2_out = mask2_out; a.proj = proj; a.probs = probs; a.depth = depth;
  const size_t smem = (size_t)Vz * DPC_ZF_TY * V * sizeof(float);
  dim3 grid(V / DPC_ZF_TY, B), block(DPC_ZF_THREADS);
#ifndef DPC_EMU
  cudaError_t e = (Kz == 21)
      ? cudaFuncSetAttribute(dpc_conv_z64_fwd_kernel<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
      : cudaFuncSetAttribute(dpc_conv_z64_fwd_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return DPC_ERR_CUDA;
#endif
  if (Kz == 21) { DPC_LAUNCH(dpc_conv_z64_fwd_kernel<21>, grid, block, smem, stream, a); }
  else { DPC_LAUNCH(dpc_conv_z64_fwd_kernel<11>, grid, block, smem, stream, a); }
  return DPC_OK;
}

static inline int dpc_conv_z_bwd_fast_launch(const float* vox, const uint32_t* mask2, const float* scale,
                                             const float* taps_rev, int Kz, int mode, float eps, float cam_dist,
                                             float max_depth, int flip_y, int B, int Vz, int V, const float* g_proj,
                                             const float* g_vox, const float* g_probs, const float* g_depth, float* d_in,
                                             float* d_scale, void* stream) {
  DpcConvZBwdArgs a;
  a.vox = vox; a.mask2 = mask2; a.scale = scale; a.taps = taps_rev; a.K = Kz; a.pl = (Kz - 1) / 2;
  a.mode = mode; a.eps = eps; a.cam_dist = cam_dist; a.max_depth = max_depth; a.flip_y = flip_y;
  a.B = B; a.Vz = Vz; a.V = V; a.TY = DPC_ZF_TY;
  a.g_proj = g_proj; a.g_vox = g_vox; a.g_probs = g_probs; a.g_depth = g_depth; a.d_in = d_in; a.d_scale = d_scale;
  const size_t smem = (size_t)Vz * DPC_ZF_TY * V * sizeof(float);
  dim3 grid(V / DPC_ZF_TY, B), block(DPC_ZF_THREADS);
#ifndef DPC_EMU
  cudaError_t e = (Kz == 21)
      ? cudaFuncSetAttribute(dpc_conv_z64_bwd_kernel<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
      : cudaFuncSetAttribute(dpc_conv_z64_bwd_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return DPC_ERR_CUDA;
#endif
  if (Kz == 21) { DPC_LAUNCH(dpc_conv_z64_bwd_kernel<21>, grid, block, smem, stream, a); }
  else { DPC_LAUNCH(dpc_conv_z64_bwd_kernel<11>, grid, block, smem, stream, a); }
  return DPC_OK;
}
