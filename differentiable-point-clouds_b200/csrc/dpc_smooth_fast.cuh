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
#define DPC_ZF_TY 4           // image rows per CTA: tile = 64 z x 4 rows x 64 x fp32 = 64 KiB, 128 threads

template <int K>
#ifndef DPC_EMU
__global__ void __launch_bounds__(32 * DPC_ZF_TY)
#else
static void
#endif
dpc_conv_z64_fwd_kernel(DpcConvZArgs a) {
  constexpr int V = DPC_F64_V, Vz = DPC_F64_V, TY = DPC_ZF_TY, RW = TY * V;
  DPC_DYN_SMEM(float, tile);                // [Vz][TY][V]
  __shared__ __align__(8) uint64_t bar;
  __shared__ float tz[K];
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

  const int ty = tid >> 5, xp = tid & 31, y = y0 + ty;
  float2 tt[K];
#pragma unroll
  for (int j = 0; j < K; ++j) tt[j] = dpc_f2(tz[j], tz[j]);
  const bool has_s = a.scale != nullptr;
  const float s = has_s ? a.scale[b] : 1.0f;
  const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
  const int yo = a.flip_y ? (V - 1 - y) : y;
  const size_t ray = ((size_t)b * V + yo) * V + 2 * xp;
  const size_t plane = (size_t)a.B * V * V;
  float2 T = dpc_f2(1.f, 1.f), proj = dpc_f2(0.f, 0.f), dep = dpc_f2(0.f, 0.f), mx = dpc_f2(-INFINITY, -INFINITY);
  uint32_t m0 = 0u, m1 = 0u, m0lo = 0u, m1lo = 0u;   // clip-pass bits of the two rays (current word / low word)
  float* vout = a.vox_out + ((size_t)b * Vz * V + y) * V + 2 * xp;
  const float* col = tile + ty * V + 2 * xp;
#pragma unroll 1
  for (int c = 0; c < Vz / 8; ++c) {
    float2 acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = dpc_f2(0.0f, 0.0f);
    dpc_col_conv_pairs<K, 8>(col, RW, c * 8, Vz, tt, acc);
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int z = c * 8 + o;
      float2 v = acc[o];
      if (has_s) {
        const float t0 = __fmul_rn(v.x, s), t1 = __fmul_rn(v.y, s);
        if (t0 >= 0.0f && t0 <= 1.0f) m0 |= 1u << (z & 31);
        if (t1 >= 0.0f && t1 <= 1.0f) m1 |= 1u << (z & 31);
        v = dpc_f2(dpc_clip01(t0), dpc_clip01(t1));
      }
      *reinterpret_cast<float2*>(vout + (size_t)z * V * V) = v;
      if (a.mode == DPC_PROJ_MAX) {
        mx = dpc_f2(fmaxf(mx.x, v.x), fmaxf(mx.y, v.y));
      } else if (a.mode != DPC_PROJ_NONE) {
        const float u0 = D.clampu ? fminf(fmaxf(v.x, D.lo), D.hi) : v.x;
        const float u1 = D.clampu ? fminf(fmaxf(v.y, D.lo), D.hi) : v.y;
        const float c0 = (z == 0) ? D.c0 : 1.0f;
        const float2 p = dpc_f2(c0 * u0 * T.x, c0 * u1 * T.y);
        T = dpc_f2(T.x * (1.0f - u0), T.y * (1.0f - u1));
        proj = dpc_f2(proj.x + p.x, proj.y + p.y);
        if (a.probs) *reinterpret_cast<float2*>(a.probs + (size_t)z * plane + ray) = p;
        if (a.depth) { const float ps = dpc_psi(z, Vz, a.cam_dist); dep = dpc_f2(fmaf(p.x, ps, dep.x), fmaf(p.y, ps, dep.y)); }
      }
    }
    if (c == 3) { m0lo = m0; m1lo = m1; m0 = 0u; m1 = 0u; }   // depth levels 0..31 done
  }
  if (a.mask2_out) {
    uint4 mw; mw.x = m0lo; mw.y = m0; mw.z = m1lo; mw.w = m1;
    *reinterpret_cast<uint4*>(a.mask2_out + (((size_t)b * V + y) * V + 2 * xp) * 2) = mw;
  }
  if (a.mode == DPC_PROJ_MAX) {
    *reinterpret_cast<float2*>(a.proj + ray) = mx;
  } else if (a.mode != DPC_PROJ_NONE) {
    const float2 pZ = dpc_f2(D.cZ * T.x, D.cZ * T.y);
    *reinterpret_cast<float2*>(a.proj + ray) = proj;
    if (a.probs) *reinterpret_cast<float2*>(a.probs + (size_t)Vz * plane + ray) = pZ;
    if (a.depth) *reinterpret_cast<float2*>(a.depth + ray) = dpc_f2(fmaf(pZ.x, a.max_depth, dep.x), fmaf(pZ.y, a.max_depth, dep.y));
  }
}

// one ray of the backward: reverse-sweep step (see dpc_conv_z_bwd_kernel for the derivation)
DPC_DEV float dpc_drc_bwd_step(const DpcDrc& D, float v, float Tk, float G, int z, float& Q) {
  const float u = D.clampu ? fminf(fmaxf(v, D.lo), D.hi) : v;
  const float Gc = G * (z == 0 ? D.c0 : 1.0f);
  float du = Tk * (Gc - Q);
  Q = fmaf(1.0f - u, Q, Gc * u);
  if (D.clampu && !(v >= D.lo && v <= D.hi)) du = 0.0f;
  return du;
}

template <int K>
#ifndef DPC_EMU
__global__ void __launch_bounds__(32 * DPC_ZF_TY)
#else
static void
#endif
dpc_conv_z64_bwd_kernel(DpcConvZBwdArgs a) {
  constexpr int V = DPC_F64_V, Vz = DPC_F64_V, TY = DPC_ZF_TY, RW = TY * V;
  DPC_DYN_SMEM(float, tile);                // [Vz][TY][V]: T_k, then dL/d(smoothed)
  __shared__ float tz[K];
  __shared__ float red[DPC_ZF_TY];
  const int tid = threadIdx.x;
  const int b = blockIdx.y, y0 = blockIdx.x * TY;
  if (tid < K) tz[tid] = a.taps[tid];
  __syncthreads();
  const int ty = tid >> 5, xp = tid & 31, y = y0 + ty;
  const bool has_s = a.scale != nullptr;
  const float s = has_s ? a.scale[b] : 1.0f;
  const float inv_s = (s != 0.0f) ? 1.0f / s : 0.0f;
  const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
  const int yo = a.flip_y ? (V - 1 - y) : y;
  const size_t ray = ((size_t)b * V + yo) * V + 2 * xp;
  const size_t plane = (size_t)a.B * V * V;
  const float* vin = a.vox + ((size_t)b * Vz * V + y) * V + 2 * xp;
  const float* gv = a.g_vox ? a.g_vox + ((size_t)b * Vz * V + y) * V + 2 * xp : nullptr;
  float* col = tile + ty * V + 2 * xp;
  const float2 gp = a.g_proj ? *reinterpret_cast<const float2*>(a.g_proj + ray) : dpc_f2(0.f, 0.f);
  const float2 gd = a.g_depth ? *reinterpret_cast<const float2*>(a.g_depth + ray) : dpc_f2(0.f, 0.f);

  if (a.mode == DPC_PROJ_MAX) {
    float2 mx = dpc_f2(-INFINITY, -INFINITY);
    for (int z = 0; z < Vz; ++z) {
      const float2 v = *reinterpret_cast<const float2*>(vin + (size_t)z * V * V);
      mx = dpc_f2(fmaxf(mx.x, v.x), fmaxf(mx.y, v.y));
    }
    int c0 = 0, c1 = 0;
    for (int z = 0; z < Vz; ++z) {
      const float2 v = *reinterpret_cast<const float2*>(vin + (size_t)z * V * V);
      c0 += (v.x == mx.x); c1 += (v.y == mx.y);
    }
    const float s0 = gp.x / (float)c0, s1 = gp.y / (float)c1;
    for (int z = 0; z < Vz; ++z) {
      const float2 v = *reinterpret_cast<const float2*>(vin + (size_t)z * V * V);
      *reinterpret_cast<float2*>(col + z * RW) = dpc_f2(v.x == mx.x ? s0 : 0.0f, v.y == mx.y ? s1 : 0.0f);
    }
  } else if (a.mode == DPC_PROJ_NONE) {
    for (int z = 0; z < Vz; ++z) *reinterpret_cast<float2*>(col + z * RW) = dpc_f2(0.f, 0.f);
  } else {
    float2 T = dpc_f2(1.f, 1.f);
#pragma unroll 8
    for (int z = 0; z < Vz; ++z) {
      const float2 v = *reinterpret_cast<const float2*>(vin + (size_t)z * V * V);
      const float u0 = D.clampu ? fminf(fmaxf(v.x, D.lo), D.hi) : v.x;
      const float u1 = D.clampu ? fminf(fmaxf(v.y, D.lo), D.hi) : v.y;
      *reinterpret_cast<float2*>(col + z * RW) = T;
      T = dpc_f2(T.x * (1.0f - u0), T.y * (1.0f - u1));
    }
    float2 gZ = dpc_f2(gd.x * a.max_depth, gd.y * a.max_depth);
    if (a.g_probs) {
      const float2 t = *reinterpret_cast<const float2*>(a.g_probs + (size_t)Vz * plane + ray);
      gZ = dpc_f2(gZ.x + t.x, gZ.y + t.y);
    }
    float Q0 = gZ.x * D.cZ, Q1 = gZ.y * D.cZ;
#pragma unroll 8
    for (int z = Vz - 1; z >= 0; --z) {
      const float2 v = *reinterpret_cast<const float2*>(vin + (size_t)z * V * V);   // second read: L1/L2 hit
      const float2 Tk = *reinterpret_cast<const float2*>(col + z * RW);
      float G0 = gp.x, G1 = gp.y;
      if (a.g_depth) { const float ps = dpc_psi(z, Vz, a.cam_dist); G0 = fmaf(gd.x, ps, G0); G1 = fmaf(gd.y, ps, G1); }
      if (a.g_probs) {
        const float2 t = *reinterpret_cast<const float2*>(a.g_probs + (size_t)z * plane + ray);
        G0 += t.x; G1 += t.y;
      }
      const float d0 = dpc_drc_bwd_step(D, v.x, Tk.x, G0, z, Q0);
      const float d1 = dpc_drc_bwd_step(D, v.y, Tk.y, G1, z, Q1);
      *reinterpret_cast<float2*>(col + z * RW) = dpc_f2(d0, d1);
    }
  }
  // + direct gradient on voxels, back through clip(. * scale)
  float ds = 0.0f;
  if (gv || has_s) {
    uint4 mw = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (a.mask2 && has_s) mw = *reinterpret_cast<const uint4*>(a.mask2 + (((size_t)b * V + y) * V + 2 * xp) * 2);
#pragma unroll 8
    for (int z = 0; z < Vz; ++z) {
      float2 dv = *reinterpret_cast<const float2*>(col + z * RW);
      if (gv) { const float2 t = *reinterpret_cast<const float2*>(gv + (size_t)z * V * V); dv = dpc_f2(dv.x + t.x, dv.y + t.y); }
      if (has_s) {
        const uint32_t w0 = (z < 32) ? mw.x : mw.y, w1 = (z < 32) ? mw.z : mw.w;
        if (!((w0 >> (z & 31)) & 1u)) dv.x = 0.0f;
        if (!((w1 >> (z & 31)) & 1u)) dv.y = 0.0f;
        const float2 v = *reinterpret_cast<const float2*>(vin + (size_t)z * V * V);
        ds = fmaf(dv.x, v.x * inv_s, ds);
        ds = fmaf(dv.y, v.y * inv_s, ds);
        dv = dpc_f2(dv.x * s, dv.y * s);
      }
      *reinterpret_cast<float2*>(col + z * RW) = dv;
    }
  }
  // transposed depth correlation (reversed taps) straight to global
  float2 tt[K];
#pragma unroll
  for (int j = 0; j < K; ++j) tt[j] = dpc_f2(tz[j], tz[j]);
  float* dout = a.d_in + ((size_t)b * Vz * V + y) * V + 2 * xp;
#pragma unroll 1
  for (int c = 0; c < Vz / 8; ++c) {
    float2 acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = dpc_f2(0.0f, 0.0f);
    dpc_col_conv_pairs<K, 8>(col, RW, c * 8, Vz, tt, acc);
#pragma unroll
    for (int o = 0; o < 8; ++o) *reinterpret_cast<float2*>(dout + (size_t)(c * 8 + o) * V * V) = acc[o];
  }
  if (a.d_scale) {
    const float v = dpc_warp_sum(ds);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      float t = 0.0f;
      for (int i = 0; i < TY; ++i) t += red[i];
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

static inline bool dpc_conv_z_fast_supported(int V, int Vz, int Kz, int plz) {
  return V == 64 && Vz == 64 && dpc_fast_k(Kz) && plz == (Kz - 1) / 2;
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
  dim3 grid(V / DPC_ZF_TY, B), block(32 * DPC_ZF_TY);
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
  dim3 grid(V / DPC_ZF_TY, B), block(32 * DPC_ZF_TY);
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
