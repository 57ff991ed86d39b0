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

// ------------------------------------------------------------------------------ taps as launch parameters
// The filter taps are warp-uniform FFMA2 operands.  Held in vector registers they cost 44 (x pass)
// / 42 (y, depth pass) registers per thread and cap the kernels at 3-4 CTAs per SM, which is what
// bounds them (gpurun rounds 19-23: every stage of a CTA is latency-exposed; FMA pipe 46 % busy).
// When the CALLER knows the taps on the host -- the reference derives sigma from the global step
// (model_pc.py:35-40), so a training loop does -- they travel inside the kernel's launch parameters
// instead: constant bank 0, fetched by LDCU into UNIFORM registers, which FFMA2 accepts as its
// second operand.  The kernels then need 40-45 registers and 5-6 CTAs fit per SM.  (Copying
// device taps into a __constant__ array per call was measured too: the copy costs more stream time
// than the kernels gain.)  Without host taps the vector-register kernels run.
struct DpcTapsXY {
  float2 px[24];      // q = 0..K: (t[q-1], t[q]) of the x taps (zero outside 0..K-1)
  float2 dy[24];      // j = 0..K-1: (t[j], t[j]) of the y taps
};

static inline float dpc_host_tap(const float* t, int K, int i, int rev) {
  return (i >= 0 && i < K) ? t[rev ? (K - 1 - i) : i] : 0.0f;
}
static inline void dpc_build_taps_pairs(const float* t, int K, int rev, float2* px) {
  for (int q = 0; q < 24; ++q) px[q] = make_float2(dpc_host_tap(t, K, q - 1, rev), dpc_host_tap(t, K, q, rev));
}
static inline void dpc_build_taps_dup(const float* t, int K, int rev, float2* d) {
  for (int j = 0; j < 24; ++j) { const float v = dpc_host_tap(t, K, j, rev); d[j] = make_float2(v, v); }
}

// ------------------------------------------------------------------------------ conv_xy, V = 64
// One CTA per depth slice (2048 CTAs at B=32, three resident per SM, so the load phase of one
// overlaps the arithmetic of the others).  A persistent, TMA-prefetching variant was measured and
// was SLOWER (barrier stalls; profiles/r01_d_*), so the plain form stays.
struct DpcConvXY64Args {
  const float* in; float* out; const float* taps_x; const float* taps_y;
  int clip_in; uint32_t* mask_out; const uint32_t* mask_in; int nslices;
  int rev; float* zero_ptr;
  int dbg;   // diagnostics only: 1 = memory path only (no correlation), 2 = arithmetic only (no global load/store)
  DpcTapsXY ht;   // TS kernels: the taps (already reversed when rev), built on the host
};

// V in {32, 64, 128}: a CTA owns one "unit" of contiguous voxels: four 32x32 slices, one 64x64 slice or
// one 128x128 slice.  The x pass works on AR rows at a time (NP passes; two at V=128 so that the
// input rows need only half a slice of smem), the y pass on the whole unit.
// NT = threads per CTA: 256 (one task per thread and phase) or 128 (two tasks per thread and
// phase, twice as many independent CTAs resident per SM to overlap load / barrier bubbles).
// TS = taps taken from the launch parameters a.ht (uniform-register operands) instead of being
// held in 44 vector registers: the kernel is then compiled for TS CTAs per SM (5 at V <= 64) instead of 3.
template <int V, int K, int NT, int TS>
#ifndef DPC_EMU
__global__ void __launch_bounds__(NT, (TS ? TS * 256 : 768) / NT)
#else
static void
#endif
dpc_conv_xy_fast_kernel(const DPC_GRID_CONSTANT DpcConvXY64Args a) {
  constexpr int S = V + 4, PL = (K - 1) / 2;
  constexpr int SPC = (V == 32) ? 4 : 1;            // slices per unit
  constexpr int UNIT = SPC * V * V;                 // voxels per CTA: 4096, 4096, 16384
  constexpr int MR = SPC * V;                       // rows of the unit: 128, 64, 128
  constexpr int AR = (V == 128) ? 64 : MR;          // rows per x pass
  constexpr int NP = MR / AR;
  constexpr int WL = ((PL + 3) / 4) * 4;            // window starts WL floats left of the first output
  constexpr int NW4 = (WL + 16 + WL) / 4;           // float4 groups in the x window
  constexpr int XR = V / 16;                        // x runs of 16 outputs per row
  constexpr int XT = AR * XR;                       // x tasks per pass: 256, 256, 512
  constexpr int YT = (V / 2) * (V / 8);             // y tasks per slice: 64, 256, 1024
  static_assert((K & 1) == 1 && K <= 21, "odd K <= 21");
  static_assert(V == 32 || V == 64 || V == 128, "V");
  static_assert((S / 4) % 2 == 1, "row pitch must be an odd number of 16-byte units");
  DPC_DYN_SMEM(float, sm);
  float* A = sm;                                     // [AR][S]
  float* M = sm + AR * S;                            // [MR][S]
  __shared__ __align__(8) float txe[24];            // E[i] = t[i-1], i = 0..K+1 (zero outside)
  __shared__ __align__(8) float txo[24];            // O[i] = E[i+1]
  __shared__ __align__(8) float2 tyd[24];           // y taps, each duplicated into a float2 (FFMA2 operand)
  const int tid = threadIdx.x;
  const size_t slice = (size_t)blockIdx.x * UNIT;   // first voxel of this CTA's unit
  if (!TS && tid < 24) {
    const int a_e = tid - 1, a_o = tid;              // tap indices behind E[tid], O[tid]
    txe[tid] = (a_e >= 0 && a_e < K) ? dpc_tap(a.taps_x, K, a_e, a.rev) : 0.0f;
    txo[tid] = (a_o >= 0 && a_o < K) ? dpc_tap(a.taps_x, K, a_o, a.rev) : 0.0f;
    const float tyv = (tid < K) ? dpc_tap(a.taps_y, K, tid, a.rev) : 0.0f;
    tyd[tid] = dpc_f2(tyv, tyv);
  }
  dpc_grid_dep_sync();
  if (a.dbg == 3) return;

#pragma unroll 1
  for (int ps = 0; ps < NP; ++ps) {
  const size_t poff = (size_t)ps * AR * V;          // first voxel of this pass within the unit
  // ---- phase 0: AR rows -> smem (float4, coalesced), clip, clip-mask bits
  {
    const float4* src = reinterpret_cast<const float4*>(a.in + slice + poff);
    constexpr int NL = AR * V / 4 / NT;
    float4 v[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) v[k] = (a.dbg == 2) ? make_float4(0.5f, 0.25f, 0.f, 1.f) : src[tid + NT * k];   // all loads in flight
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      const int i = tid + NT * k;             // float4 index within the pass
      if (a.mask_out) {
        unsigned nib = ((v[k].x >= 0.0f && v[k].x <= 1.0f) ? 1u : 0u) | ((v[k].y >= 0.0f && v[k].y <= 1.0f) ? 2u : 0u) |
                       ((v[k].z >= 0.0f && v[k].z <= 1.0f) ? 4u : 0u) | ((v[k].w >= 0.0f && v[k].w <= 1.0f) ? 8u : 0u);
        unsigned word = nib << (4 * (tid & 7));
        word |= __shfl_xor_sync(DPC_FULL, word, 1);
        word |= __shfl_xor_sync(DPC_FULL, word, 2);
        word |= __shfl_xor_sync(DPC_FULL, word, 4);
        if ((tid & 7) == 0) a.mask_out[((slice + poff) >> 5) + (i >> 3)] = word;
      }
      if (a.clip_in) { v[k].x = dpc_clip01(v[k].x); v[k].y = dpc_clip01(v[k].y); v[k].z = dpc_clip01(v[k].z); v[k].w = dpc_clip01(v[k].w); }
      const int e = 4 * i, row = e / V, c0 = e % V;     // row within the pass
      *reinterpret_cast<float4*>(&A[row * S + c0]) = v[k];
      if (a.zero_ptr) reinterpret_cast<float4*>(a.zero_ptr + slice + poff)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();

  // ---- phase 1: x correlation.  Task = (row of the pass, run r of 16 outputs); a warp = 32 rows, one r.
  if (a.dbg == 1 || a.dbg == 4) {
    for (int i = tid; i < AR * S; i += NT) M[ps * AR * S + i] = A[i];
  } else
#pragma unroll 1
  for (int task = tid; task < XT; task += NT) {
    const int y = task % AR, r = task / AR;
    const int x0 = r * 16;
    // tap pairs (t[a], t[a+1]), a = -1..K-1, index q = a+1: aligned float2 in E for even q, in O for odd q
    float2 tp[TS ? 1 : K + 1];
    if (!TS) {
#pragma unroll
      for (int q = 0; q < K + 1; ++q)
        tp[TS ? 0 : q] = (q & 1) ? *reinterpret_cast<const float2*>(txo + (q - 1)) : *reinterpret_cast<const float2*>(txe + q);
    }
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
          if (4 * g + 2 * h - o - WL + PL >= -1 && 4 * g + 2 * h - o - WL + PL <= K - 1) {
            constexpr int dummy = 0; (void)dummy;
            const int q = 4 * g + 2 * h - o - WL + PL + 1;   // compile-time after unrolling
            const float2 tq = TS ? a.ht.px[q] : tp[TS ? 0 : q];
            acc[o] = dpc_ffma2(w, tq, acc[o]);
          }
        }
      }
    }
    float* dst = M + (ps * AR + y) * S + x0;
#pragma unroll
    for (int o = 0; o < 16; o += 4) {
      *reinterpret_cast<float4*>(dst + o) = make_float4(acc[o].x + acc[o].y, acc[o + 1].x + acc[o + 1].y,
                                                         acc[o + 2].x + acc[o + 2].y, acc[o + 3].x + acc[o + 3].y);
    }
  }
  __syncthreads();
  }  // x passes

  // ---- phase 2: y correlation.  Task = (slice, x pair, run of 8 rows).
#pragma unroll 1
  for (int task = tid; task < SPC * YT; task += NT) {
    const int sl = task / YT, t = task % YT;
    const int xp = t % (V / 2), y0 = (t / (V / 2)) * 8;
    float2 acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = dpc_f2(0.0f, 0.0f);
    const size_t base = slice + (size_t)sl * V * V;
    uint32_t mw[8];                       // clip-mask words of the 8 output rows, fetched before the arithmetic
    if (a.mask_in) {
#pragma unroll
      for (int o = 0; o < 8; ++o) mw[o] = a.mask_in[(base + (size_t)(y0 + o) * V + 2 * xp) >> 5];
    }
    if (a.dbg == 1 || a.dbg == 5) {
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = *reinterpret_cast<const float2*>(M + (sl * V + y0 + o) * S + 2 * xp);
    } else {
      dpc_col_conv_pairs<K, 8>(M + sl * V * S + 2 * xp, S, y0, V, TS ? a.ht.dy : tyd, acc);
    }
    float* dst = a.out + base + (size_t)y0 * V + 2 * xp;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float2 v = acc[o];
      if (a.mask_in) {
        const uint32_t wbits = mw[o] >> ((2 * xp) & 31);      // V is a multiple of 32: the row offset drops out
        if (!(wbits & 1u)) v.x = 0.0f;
        if (!(wbits & 2u)) v.y = 0.0f;
      }
      if (a.dbg != 2 || v.x == 123456.0f) *reinterpret_cast<float2*>(dst + (size_t)o * V) = v;
    }
  }
}

// Rows of `nrows` x `row_bytes` from global (row pitch src_pitch floats) into smem (row pitch
// dst_pitch floats) through the TMA engine, completion on `bar`.  Called by ALL lanes of warp 0:
// lane 0 posts the expected byte count, then every lane issues its share of the bulk copies.
DPC_DEV void dpc_warp_bulk_rows(float* dst, int dst_pitch, const float* src, size_t src_pitch, int nrows,
                                unsigned row_bytes, uint64_t* bar) {
  const int lane = threadIdx.x & 31;
#ifndef DPC_EMU
  const unsigned bb = (unsigned)__cvta_generic_to_shared(bar);
  if (lane == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"((unsigned)nrows * row_bytes) : "memory");
  __syncwarp();
  for (int r = lane; r < nrows; r += 32) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + (size_t)r * dst_pitch);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src + (size_t)r * src_pitch), "r"(row_bytes), "r"(bb) : "memory");
  }
#else
  for (int r = lane; r < nrows; r += 32) memcpy(dst + (size_t)r * dst_pitch, src + (size_t)r * src_pitch, row_bytes);
  __syncwarp();
  if (lane == 0) dpc_emu::mbar_complete(bar);
#endif
}

// Alternative tile load: the whole CTA copies nrows x 512 B with 16-byte cp.async (LDGSTS), no TMA op
// per row.  Returns after the data is visible to every thread.
static DPC_KNOB_T dpc_z_tile_cpasync = 0;   // experiment knob (dpc_debug_set key 6)
DPC_DEV void dpc_cta_cpasync_rows(float* dst, const float* src, size_t src_pitch, int nrows) {
  // rows of 128 floats = 32 chunks of 16 bytes; thread t copies chunk (t & 31) of rows (t >> 5), +8, ...
  const int ch = threadIdx.x & 31;
  for (int r = threadIdx.x >> 5; r < nrows; r += (int)(blockDim.x >> 5)) {
#ifndef DPC_EMU
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + (size_t)r * 128 + ch * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + (size_t)r * src_pitch + ch * 4) : "memory");
#else
    memcpy(dst + (size_t)r * 128 + ch * 4, src + (size_t)r * src_pitch + ch * 4, 16);
#endif
  }
#ifndef DPC_EMU
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
  __syncthreads();
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

// Forward, V = Vz in {32, 64, 128}.  A CTA owns 128 rays (TY = 128/V image rows) over all depth
// levels: tile [Vz][128] fp32 (16 / 32 / 64 KiB) loaded by Vz TMA bulk copies of 512 B.  256 threads =
// 4 depth segments x 64 ray pairs; a thread runs CPS = Vz/32 register chunks of 8 outputs.  The ray
// scan starts from T = 1 in every segment and the segments are combined through smem:
//   proj = S0 + T0 (S1 + T1 (S2 + T2 S3)),  max = max of the segment maxima.
// At 64^3 that is 1024 CTAs, four resident per SM.
template <int V, int K, int MINB, bool CT>
#ifndef DPC_EMU
__global__ void __launch_bounds__(DPC_ZF_THREADS, MINB)
#else
static void
#endif
dpc_conv_z_fast_fwd_kernel(const DPC_GRID_CONSTANT DpcConvZArgs a) {
  constexpr int Vz = V, TY = 128 / V, RW = 128, CPS = Vz / 32, NW = Vz / 32;
  static_assert(V == 32 || V == 64 || V == 128, "V");
  DPC_DYN_SMEM(float, tile);                // [Vz][TY][V] = [Vz][128]
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) float2 tzd[K];          // taps, each duplicated into a float2 (FFMA2 operand)
  __shared__ __align__(8) float comb[4][128][2];  // per depth segment and ray: (T, S) or (max, -)
  const int tid = threadIdx.x;
  const int b = blockIdx.y, y0 = blockIdx.x * TY;
  if (!CT && tid < K) { const float t = dpc_tap(a.taps, K, tid, a.rev); tzd[tid] = dpc_f2(t, t); }
  const float* src = a.in + ((size_t)b * Vz * V + y0) * V;
  if (tid == 0) dpc_mbar_init(&bar, 1);
  dpc_grid_dep_sync();
  __syncthreads();
  if (a.TY < 0) {
    dpc_cta_cpasync_rows(tile, src, (size_t)V * V, Vz);
  } else {
    if (tid < 32) dpc_warp_bulk_rows(tile, RW, src, (size_t)V * V, Vz, RW * 4, &bar);
    dpc_mbar_wait(&bar, 0);
    __syncthreads();
  }

  // depth segment and ray pair of this thread.  At V = 64 consecutive warps take the four segments of
  // one image row (measured ~3 us faster per launch than consecutive warps taking the two rows of one segment).
  const int seg = (V == 64) ? ((tid >> 5) & 3) : (tid >> 6);
  const int pidx = (V == 64) ? (((tid >> 7) << 5) | (tid & 31)) : (tid & 63);
  const int ty = pidx / (V / 2), xp = pidx % (V / 2), y = y0 + ty;
  float2 ttr[CT ? 1 : K];                         // taps in registers: measured faster than reading smem per FFMA2
  if (!CT) {
#pragma unroll
    for (int j = 0; j < K; ++j) ttr[CT ? 0 : j] = tzd[j];
  }
  const float2* tt = CT ? a.ht.dz : ttr;          // CT: launch-parameter taps -> uniform-register operands
  const bool has_s = a.scale != nullptr;
  const float s = has_s ? a.scale[b] : 1.0f;
  const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
  float2 T = dpc_f2(1.f, 1.f), S = dpc_f2(0.f, 0.f), mx = dpc_f2(-INFINITY, -INFINITY);
  uint32_t m0 = 0u, m1 = 0u;   // clip-pass bits of the two rays for this depth segment (8*CPS bits each)
  float* vout = a.vox_out + ((size_t)b * Vz * V + y) * V + 2 * xp;
  const float* col = tile + 2 * pidx;
#pragma unroll 1
  for (int c = 0; c < CPS; ++c) {
    const int zc = (CPS * seg + c) * 8;
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
    // per ray NW 32-bit words of clip-pass bits, bit z = depth level z; this thread owns CPS bytes of each ray
    uint8_t* mp = reinterpret_cast<uint8_t*>(a.mask2_out + (((size_t)b * V + y) * V + 2 * xp) * NW) + seg * CPS;
    if (CPS == 1) { mp[0] = (uint8_t)m0; mp[4 * NW] = (uint8_t)m1; }
    else if (CPS == 2) { *reinterpret_cast<uint16_t*>(mp) = (uint16_t)m0; *reinterpret_cast<uint16_t*>(mp + 4 * NW) = (uint16_t)m1; }
    else { *reinterpret_cast<uint32_t*>(mp) = m0; *reinterpret_cast<uint32_t*>(mp + 4 * NW) = m1; }
  }
  if (a.mode == DPC_PROJ_NONE) return;
  *reinterpret_cast<float2*>(&comb[seg][2 * pidx][0]) = (a.mode == DPC_PROJ_MAX) ? dpc_f2(mx.x, 0.f) : dpc_f2(T.x, S.x);
  *reinterpret_cast<float2*>(&comb[seg][2 * pidx + 1][0]) = (a.mode == DPC_PROJ_MAX) ? dpc_f2(mx.y, 0.f) : dpc_f2(T.y, S.y);
  __syncthreads();
  if (seg == 0) {
    float2 out;
    if (a.mode == DPC_PROJ_MAX) {
      out = mx;
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        out.x = fmaxf(out.x, comb[q][2 * pidx][0]);
        out.y = fmaxf(out.y, comb[q][2 * pidx + 1][0]);
      }
    } else {
      float r0 = comb[3][2 * pidx][1], r1 = comb[3][2 * pidx + 1][1];
#pragma unroll
      for (int q = 2; q >= 0; --q) {
        r0 = fmaf(comb[q][2 * pidx][0], r0, comb[q][2 * pidx][1]);
        r1 = fmaf(comb[q][2 * pidx + 1][0], r1, comb[q][2 * pidx + 1][1]);
      }
      out = dpc_f2(r0, r1);
    }
    const int yo = a.flip_y ? (V - 1 - y) : y;
    *reinterpret_cast<float2*>(a.proj + ((size_t)b * V + yo) * V + 2 * xp) = out;
  }
}

// LEAN = the training configuration (DRC projection, occupancy scale present, no gradient arriving
// directly at `voxels`): phase 1 is compiled without any of the per-level mode / option tests.
template <int K, bool LEAN>
#ifndef DPC_EMU
__global__ void __launch_bounds__(DPC_ZF_THREADS, 3)
#else
static void
#endif
dpc_conv_z64_bwd_kernel(DpcConvZBwdArgs a) {
  constexpr int V = DPC_F64_V, Vz = DPC_F64_V, TY = DPC_ZF_TY, RW = TY * V;
  DPC_DYN_SMEM(float, tile);                // [Vz][TY][V]: forward voxels, overwritten by dL/d(smoothed)
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) float2 tzd[K];
  __shared__ float red[DPC_ZF_THREADS / 32];
  const int tid = threadIdx.x;
  const int b = blockIdx.y, y0 = blockIdx.x * TY;
  if (tid < K) { const float t = dpc_tap(a.taps, K, tid, a.rev); tzd[tid] = dpc_f2(t, t); }
  // the forward's voxels of these 4 image rows, all depth levels: TMA bulk copies into the tile
  if (tid == 0) dpc_mbar_init(&bar, 1);
  dpc_grid_dep_sync();
  __syncthreads();
  if (tid < 32) dpc_warp_bulk_rows(tile, RW, a.vox + ((size_t)b * Vz * V + y0) * V, (size_t)V * V, Vz, RW * 4, &bar);
  const bool has_s = a.scale != nullptr;
  const float s = has_s ? a.scale[b] : 1.0f;
  const float inv_s = (s != 0.0f) ? 1.0f / s : 0.0f;
  const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
  float ds = 0.0f;
  dpc_mbar_wait(&bar, 0);
  // ---- phase 1: one thread per ray, in place in the tile (a thread touches only its own column).
  // Silhouette gradient g:  dL/du_k = g * prod_{j != k} (1-u_j) = g * T_Z / (1-u_k)  for k > 0 and
  // g * (c_0 - 1 + T_Z / (1-u_0)) for k = 0, T_Z = prod_j (1-u_j).  1-u >= eps (the clip), and where
  // T_Z underflows the true value is below 1e-35, so the quotient form is exact to ~1e-7 relative
  // or negligible in absolute terms; no second scan, no carried state.
  if (LEAN) {
    const int ty = tid >> 6, x = tid & 63, y = y0 + ty;
    const int yo = a.flip_y ? (V - 1 - y) : y;
    float* col = tile + ty * V + x;
    const float gp = a.g_proj[((size_t)b * V + yo) * V + x];
    const uint2 mw = *reinterpret_cast<const uint2*>(a.mask2 + (((size_t)b * V + y) * V + x) * 2);
    float T0 = 1.0f, T1 = 1.0f, T2 = 1.0f, T3 = 1.0f;   // four partial products: shorter dependency chains
#pragma unroll 4
    for (int z = 0; z < Vz; z += 4) {
      T0 *= 1.0f - fminf(fmaxf(col[(z + 0) * RW], D.lo), D.hi);
      T1 *= 1.0f - fminf(fmaxf(col[(z + 1) * RW], D.lo), D.hi);
      T2 *= 1.0f - fminf(fmaxf(col[(z + 2) * RW], D.lo), D.hi);
      T3 *= 1.0f - fminf(fmaxf(col[(z + 3) * RW], D.lo), D.hi);
    }
    const float gT = gp * ((T0 * T1) * (T2 * T3));
    float dsv = 0.0f;    // sum of dv * voxel; the 1/scale factor is applied once at the end
#pragma unroll
    for (int hw = 0; hw < 2; ++hw) {
      uint32_t wbits = hw ? mw.y : mw.x;
#pragma unroll 8
      for (int zz = 0; zz < 32; ++zz) {
        const int z = hw * 32 + zz;
        const float v = col[z * RW];
        const float u = fminf(fmaxf(v, D.lo), D.hi);
        float dv = __fdividef(gT, 1.0f - u);
        if (z == 0) dv = fmaf(gp, D.c0 - 1.0f, dv);
        if ((u != v) || !(wbits & 1u)) dv = 0.0f;
        wbits >>= 1;
        dsv = fmaf(dv, v, dsv);
        col[z * RW] = dv * s;
      }
    }
    ds = dsv * inv_s;
  } else {
    const int ty = tid >> 6, x = tid & 63, y = y0 + ty;
    const int yo = a.flip_y ? (V - 1 - y) : y;
    const float* gv = a.g_vox ? a.g_vox + ((size_t)b * Vz * V + y) * V + x : nullptr;
    float* col = tile + ty * V + x;
    const float gp = a.g_proj ? a.g_proj[((size_t)b * V + yo) * V + x] : 0.0f;
    uint32_t mlo = 0xffffffffu, mhi = 0xffffffffu;
    if (a.mask2 && has_s) {
      const uint2 mw = *reinterpret_cast<const uint2*>(a.mask2 + (((size_t)b * V + y) * V + x) * 2);
      mlo = mw.x; mhi = mw.y;
    }
    float share = 0.0f, mx = -INFINITY, gT = 0.0f;
    if (a.mode == DPC_PROJ_MAX) {
#pragma unroll 8
      for (int z = 0; z < Vz; ++z) mx = fmaxf(mx, col[z * RW]);
      int cnt = 0;
#pragma unroll 8
      for (int z = 0; z < Vz; ++z) cnt += (col[z * RW] == mx) ? 1 : 0;
      share = gp / (float)cnt;     // TF _MaxGrad: ties share the gradient equally
    } else if (a.mode != DPC_PROJ_NONE) {
      float T0 = 1.0f, T1 = 1.0f, T2 = 1.0f, T3 = 1.0f;   // four partial products: shorter dependency chains
#pragma unroll 4
      for (int z = 0; z < Vz; z += 4) {
        const float v0 = col[(z + 0) * RW], v1 = col[(z + 1) * RW], v2 = col[(z + 2) * RW], v3 = col[(z + 3) * RW];
        T0 *= 1.0f - (D.clampu ? fminf(fmaxf(v0, D.lo), D.hi) : v0);
        T1 *= 1.0f - (D.clampu ? fminf(fmaxf(v1, D.lo), D.hi) : v1);
        T2 *= 1.0f - (D.clampu ? fminf(fmaxf(v2, D.lo), D.hi) : v2);
        T3 *= 1.0f - (D.clampu ? fminf(fmaxf(v3, D.lo), D.hi) : v3);
      }
      gT = gp * ((T0 * T1) * (T2 * T3));
    }
#pragma unroll 8
    for (int z = 0; z < Vz; ++z) {
      const float v = col[z * RW];
      float dv = 0.0f;
      if (a.mode == DPC_PROJ_MAX) {
        dv = (v == mx) ? share : 0.0f;
      } else if (a.mode != DPC_PROJ_NONE) {
        const float u = D.clampu ? fminf(fmaxf(v, D.lo), D.hi) : v;
        dv = __fdividef(gT, 1.0f - u);
        if (z == 0) dv += gp * (D.c0 - 1.0f);
        if (u != v) dv = 0.0f;                       // clip_by_value passes lo <= v <= hi only
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
    const float* col = tile + ty * V + 2 * xp;
    float* dout = a.d_in + ((size_t)b * Vz * V + y) * V + 2 * xp;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int zc = (4 * h + c) * 8;
      float2 acc[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = dpc_f2(0.0f, 0.0f);
      dpc_col_conv_pairs<K, 8>(col, RW, zc, Vz, tzd, acc);
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

// Lean backward (training configuration: DRC silhouette gradient only, occupancy scale present),
// V = Vz in {32, 64, 128}, same 128-ray tiles as the forward.  Thread = (depth half, ray): the
// product over a ray is split in two halves exchanged through smem, after which every level's
// gradient is independent (quotient form, see dpc_conv_z64_bwd_kernel).  Then 4 depth segments x 64
// ray pairs for the transposed correlation.
template <int V, int K, int MINB, bool CT>
#ifndef DPC_EMU
__global__ void __launch_bounds__(DPC_ZF_THREADS, MINB)
#else
static void
#endif
dpc_conv_z_fast_bwd_lean_kernel(const DPC_GRID_CONSTANT DpcConvZBwdArgs a) {
  constexpr int Vz = V, TY = 128 / V, RW = 128, CPS = Vz / 32, NW = Vz / 32, HL = Vz / 2;
  DPC_DYN_SMEM(float, tile);                // [Vz][128]: forward voxels, overwritten by dL/d(smoothed)
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) float2 tzd[K];
  __shared__ float pp[2][128];
  __shared__ float red[DPC_ZF_THREADS / 32];
  const int tid = threadIdx.x;
  const int b = blockIdx.y, y0 = blockIdx.x * TY;
  if (!CT && tid < K) { const float t = dpc_tap(a.taps, K, tid, a.rev); tzd[tid] = dpc_f2(t, t); }
  if (tid == 0) dpc_mbar_init(&bar, 1);
  dpc_grid_dep_sync();
  __syncthreads();
  const bool use_cpasync = a.TY < 0;
  if (use_cpasync) dpc_cta_cpasync_rows(tile, a.vox + ((size_t)b * Vz * V + y0) * V, (size_t)V * V, Vz);
  else if (tid < 32) dpc_warp_bulk_rows(tile, RW, a.vox + ((size_t)b * Vz * V + y0) * V, (size_t)V * V, Vz, RW * 4, &bar);
  const float s = a.scale[b];
  const float inv_s = (s != 0.0f) ? 1.0f / s : 0.0f;
  const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
  const int h = tid >> 7, rx = tid & 127;            // depth half, ray within the tile
  const int ty = rx / V, x = rx % V, y = y0 + ty;
  const int yo = a.flip_y ? (V - 1 - y) : y;
  const float gp = a.g_proj[((size_t)b * V + yo) * V + x];
  const uint32_t* mrow = a.mask2 + (((size_t)b * V + y) * V + x) * NW;
  float* col = tile + (size_t)(HL * h) * RW + rx;
  if (!use_cpasync) dpc_mbar_wait(&bar, 0);
  {
    float T0 = 1.0f, T1 = 1.0f, T2 = 1.0f, T3 = 1.0f;
#pragma unroll 4
    for (int z = 0; z < HL; z += 4) {
      T0 *= 1.0f - fminf(fmaxf(col[(z + 0) * RW], D.lo), D.hi);
      T1 *= 1.0f - fminf(fmaxf(col[(z + 1) * RW], D.lo), D.hi);
      T2 *= 1.0f - fminf(fmaxf(col[(z + 2) * RW], D.lo), D.hi);
      T3 *= 1.0f - fminf(fmaxf(col[(z + 3) * RW], D.lo), D.hi);
    }
    pp[h][rx] = (T0 * T1) * (T2 * T3);
  }
  __syncthreads();
  const float gT = gp * (pp[0][rx] * pp[1][rx]);
  float dsv = 0.0f;
  constexpr int NB = HL < 32 ? HL : 32;              // levels served by one mask word
#pragma unroll 1
  for (int zb = 0; zb < HL; zb += NB) {
    const int zg = HL * h + zb;                       // global depth level of this block
    uint32_t wbits = mrow[zg >> 5] >> (zg & 31);
#pragma unroll 8
    for (int zz = 0; zz < NB; ++zz) {
      const float v = col[(zb + zz) * RW];
      const float u = fminf(fmaxf(v, D.lo), D.hi);
      float dv = __fdividef(gT, 1.0f - u);
      if (zg + zz == 0) dv = fmaf(gp, D.c0 - 1.0f, dv);
      if ((u != v) || !(wbits & 1u)) dv = 0.0f;
      wbits >>= 1;
      dsv = fmaf(dv, v, dsv);
      col[(zb + zz) * RW] = dv * s;
    }
  }
  const float ds = dsv * inv_s;
  __syncthreads();
  {
    const int seg = (V == 64) ? ((tid >> 5) & 3) : (tid >> 6);
    const int pidx = (V == 64) ? (((tid >> 7) << 5) | (tid & 31)) : (tid & 63);
    const int wy = pidx / (V / 2), xp = pidx % (V / 2), yy = y0 + wy;
    const float* c2 = tile + 2 * pidx;
    float* dout = a.d_in + ((size_t)b * Vz * V + yy) * V + 2 * xp;
    float2 ttr[CT ? 1 : K];
    if (!CT) {
#pragma unroll
      for (int j = 0; j < K; ++j) ttr[CT ? 0 : j] = tzd[j];
    }
    const float2* tt = CT ? a.ht.dz : ttr;
#pragma unroll 1
    for (int c = 0; c < CPS; ++c) {
      const int zc = (CPS * seg + c) * 8;
      float2 acc[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = dpc_f2(0.0f, 0.0f);
      dpc_col_conv_pairs<K, 8>(c2, RW, zc, Vz, tt, acc);
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
static DPC_KNOB_T dpc_xy_dbg = 0;           // diagnostics knob (dpc_debug_set key 7)
static DPC_KNOB_T dpc_ignore_host_taps = 0; // experiment knob (dpc_debug_set key 5): 1 = run the vector-register kernels even when host taps are given

static inline bool dpc_fast_k(int K) { return K == 21 || K == 11; }

static inline bool dpc_conv_xy_fast_supported(int V, int Kx, int plx, int Ky, int ply) {
  return (V == 128 || V == 64 || V == 32) && Kx == Ky && dpc_fast_k(Kx) && plx == (Kx - 1) / 2 && ply == (Ky - 1) / 2;
}

template <int V, int K, int TS>
static inline int dpc_conv_xy_fast_go(const DpcConvXY64Args& a, void* stream) {
  constexpr int NT = 256, S = V + 4, MR = (V == 32) ? 128 : V, AR = (V == 128) ? 64 : MR;
  const size_t smem = (size_t)(AR + MR) * S * sizeof(float);
#ifndef DPC_EMU
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(dpc_conv_xy_fast_kernel<V, K, NT, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return DPC_ERR_CUDA;
#endif
  DPC_LAUNCH((dpc_conv_xy_fast_kernel<V, K, NT, TS>), dim3(a.nslices), dim3(NT), smem, stream, a);
  return DPC_OK;
}

template <int V, int K>
static inline int dpc_conv_xy_fast_pick(const DpcConvXY64Args& a, bool host_taps, void* stream) {
  constexpr int CTAS = (V == 128) ? 2 : 5;       // the 99 KiB of smem at V = 128 allow two CTAs per SM either way
  return host_taps ? dpc_conv_xy_fast_go<V, K, CTAS>(a, stream) : dpc_conv_xy_fast_go<V, K, 0>(a, stream);
}

// hx / hy: the same taps on the HOST (nullable): given both, they travel in the launch parameters.
static inline int dpc_conv_xy_fast_launch(const float* in, float* out, const float* taps_x, const float* taps_y, int K,
                                          int B, int Vz, int V, int clip_in, uint32_t* mask_out, const uint32_t* mask_in,
                                          int rev, float* zero_ptr, const float* hx, const float* hy, void* stream) {
  if ((((uintptr_t)in) & 15u) || (((uintptr_t)out) & 7u)) return DPC_ERR_ARG;
  const int64_t voxels = (int64_t)B * Vz * V * V;
  const int unit = (V == 128) ? 16384 : 4096;
  if (voxels % unit != 0) return DPC_ERR_SHAPE;     // V=32 needs B*Vz to be a multiple of 4
  DpcConvXY64Args a;
  a.in = in; a.out = out; a.taps_x = taps_x; a.taps_y = taps_y; a.clip_in = clip_in; a.mask_out = mask_out; a.mask_in = mask_in;
  a.nslices = (int)(voxels / unit); a.rev = rev; a.zero_ptr = zero_ptr; a.dbg = dpc_xy_dbg;
  const bool host_taps = hx && hy && !dpc_ignore_host_taps;
  if (host_taps) { dpc_build_taps_pairs(hx, K, rev, a.ht.px); dpc_build_taps_dup(hy, K, rev, a.ht.dy); }
  else { for (int i = 0; i < 24; ++i) { a.ht.px[i] = make_float2(0.f, 0.f); a.ht.dy[i] = make_float2(0.f, 0.f); } }
  if (V == 64) return K == 21 ? dpc_conv_xy_fast_pick<64, 21>(a, host_taps, stream) : dpc_conv_xy_fast_pick<64, 11>(a, host_taps, stream);
  if (V == 32) return K == 21 ? dpc_conv_xy_fast_pick<32, 21>(a, host_taps, stream) : dpc_conv_xy_fast_pick<32, 11>(a, host_taps, stream);
  return K == 21 ? dpc_conv_xy_fast_pick<128, 21>(a, host_taps, stream) : dpc_conv_xy_fast_pick<128, 11>(a, host_taps, stream);
}

// `extras`: drc_probs / proj_depth outputs (forward) or their gradients (backward) requested
static inline bool dpc_fast_v(int V) { return V == 32 || V == 64 || V == 128; }
static inline bool dpc_conv_z_fast_supported(int V, int Vz, int Kz, int plz, bool extras) {
  return dpc_fast_v(V) && Vz == V && dpc_fast_k(Kz) && plz == (Kz - 1) / 2 && !extras;
}

// CT (host taps): 40 registers -> six 256-thread CTAs per SM at V <= 64 (the 64 KiB tile of V = 128 allows three either way)
template <int V, int K, bool CT>
static inline int dpc_conv_z_fwd_fast_go1(const DpcConvZArgs& a, int B, void* stream) {
  constexpr int MINB = (V == 128) ? 3 : (CT ? 6 : 4);
  const size_t smem = (size_t)V * 128 * sizeof(float);
#ifndef DPC_EMU
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(dpc_conv_z_fast_fwd_kernel<V, K, MINB, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return DPC_ERR_CUDA;
#endif
  DPC_LAUNCH((dpc_conv_z_fast_fwd_kernel<V, K, MINB, CT>), dim3(V / (128 / V), B), dim3(DPC_ZF_THREADS), smem, stream, a);
  return DPC_OK;
}
template <int V, int K>
static inline int dpc_conv_z_fwd_fast_go(const DpcConvZArgs& a, int B, void* stream) {
  return a.use_ht ? dpc_conv_z_fwd_fast_go1<V, K, true>(a, B, stream) : dpc_conv_z_fwd_fast_go1<V, K, false>(a, B, stream);
}

static inline void dpc_set_taps_z(DpcTapsZ* ht, int* use_ht, const float* hz, int K, int rev) {
  *use_ht = (hz && !dpc_ignore_host_taps) ? 1 : 0;
  if (*use_ht) dpc_build_taps_dup(hz, K, rev, ht->dz);
  else for (int i = 0; i < 24; ++i) ht->dz[i] = make_float2(0.f, 0.f);
}

static inline int dpc_conv_z_fwd_fast_launch(const float* in, const float* taps_z, int Kz, const float* scale, int mode,
                                             float eps, float cam_dist, float max_depth, int flip_y, int B, int Vz, int V,
                                             float* vox_out, uint32_t* mask2_out, float* proj, float* probs, float* depth,
                                             const float* hz, void* stream) {
  DpcConvZArgs a;
  a.in = in; a.taps = taps_z; a.K = Kz; a.pl = (Kz - 1) / 2; a.rev = 0; a.scale = scale; a.mode = mode; a.eps = eps;
  a.cam_dist = cam_dist; a.max_depth = max_depth; a.flip_y = flip_y; a.B = B; a.Vz = Vz; a.V = V;
  a.TY = dpc_z_tile_cpasync ? -(128 / V) : (128 / V);      // sign = tile load method (experiment knob)
  a.vox_out = vox_out; a.mask2_out = mask2_out; a.proj = proj; a.probs = probs; a.depth = depth;
  dpc_set_taps_z(&a.ht, &a.use_ht, hz, Kz, 0);
  if (V == 32) return Kz == 21 ? dpc_conv_z_fwd_fast_go<32, 21>(a, B, stream) : dpc_conv_z_fwd_fast_go<32, 11>(a, B, stream);
  if (V == 64) return Kz == 21 ? dpc_conv_z_fwd_fast_go<64, 21>(a, B, stream) : dpc_conv_z_fwd_fast_go<64, 11>(a, B, stream);
  return Kz == 21 ? dpc_conv_z_fwd_fast_go<128, 21>(a, B, stream) : dpc_conv_z_fwd_fast_go<128, 11>(a, B, stream);
}

template <int V, int K, bool CT>
static inline int dpc_conv_z_bwd_lean_go1(const DpcConvZBwdArgs& a, int B, void* stream) {
  constexpr int MINB = (V == 128) ? 3 : (CT ? 6 : 4);
  const size_t smem = (size_t)V * 128 * sizeof(float);
#ifndef DPC_EMU
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(dpc_conv_z_fast_bwd_lean_kernel<V, K, MINB, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return DPC_ERR_CUDA;
#endif
  DPC_LAUNCH((dpc_conv_z_fast_bwd_lean_kernel<V, K, MINB, CT>), dim3(V / (128 / V), B), dim3(DPC_ZF_THREADS), smem, stream, a);
  return DPC_OK;
}
template <int V, int K>
static inline int dpc_conv_z_bwd_lean_go(const DpcConvZBwdArgs& a, int B, void* stream) {
  return a.use_ht ? dpc_conv_z_bwd_lean_go1<V, K, true>(a, B, stream) : dpc_conv_z_bwd_lean_go1<V, K, false>(a, B, stream);
}

// hz: the FORWARD taps on the host (nullable); `rev` says the kernel is to read them back to front.
static inline int dpc_conv_z_bwd_fast_launch(const float* vox, const uint32_t* mask2, const float* scale,
                                             const float* taps_rev, int Kz, int mode, float eps, float cam_dist,
                                             float max_depth, int flip_y, int B, int Vz, int V, const float* g_proj,
                                             const float* g_vox, const float* g_probs, const float* g_depth, float* d_in,
                                             float* d_scale, int rev, const float* hz, void* stream) {
  DpcConvZBwdArgs a = {};
  a.vox = vox; a.mask2 = mask2; a.scale = scale; a.taps = taps_rev; a.K = Kz; a.pl = (Kz - 1) / 2; a.rev = rev;
  a.mode = mode; a.eps = eps; a.cam_dist = cam_dist; a.max_depth = max_depth; a.flip_y = flip_y;
  a.B = B; a.Vz = Vz; a.V = V; a.TY = DPC_ZF_TY;
  a.g_proj = g_proj; a.g_vox = g_vox; a.g_probs = g_probs; a.g_depth = g_depth; a.d_in = d_in; a.d_scale = d_scale;
  const bool lean = (mode == DPC_PROJ_DRC) && scale && mask2 && g_proj && !g_vox;
  dpc_set_taps_z(&a.ht, &a.use_ht, lean ? hz : nullptr, Kz, rev);
  if (lean) {
    if (dpc_z_tile_cpasync) a.TY = -1;
    if (V == 32) return Kz == 21 ? dpc_conv_z_bwd_lean_go<32, 21>(a, B, stream) : dpc_conv_z_bwd_lean_go<32, 11>(a, B, stream);
    if (V == 64) return Kz == 21 ? dpc_conv_z_bwd_lean_go<64, 21>(a, B, stream) : dpc_conv_z_bwd_lean_go<64, 11>(a, B, stream);
    return Kz == 21 ? dpc_conv_z_bwd_lean_go<128, 21>(a, B, stream) : dpc_conv_z_bwd_lean_go<128, 11>(a, B, stream);
  }
  if (V != 64) return DPC_ERR_ARG;   // caller falls back to the generic kernel (see dpc_conv_z_bwd_fast_general_ok)
  const size_t smem = (size_t)Vz * DPC_ZF_TY * V * sizeof(float);
  dim3 grid(V / DPC_ZF_TY, B), block(DPC_ZF_THREADS);
#ifndef DPC_EMU
#define DPC_ZB_LAUNCH(KK, LL) do { \
    if (cudaFuncSetAttribute(dpc_conv_z64_bwd_kernel<KK, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DPC_ERR_CUDA; \
    DPC_LAUNCH((dpc_conv_z64_bwd_kernel<KK, LL>), grid, block, smem, stream, a); } while (0)
#else
#define DPC_ZB_LAUNCH(KK, LL) do { DPC_LAUNCH((dpc_conv_z64_bwd_kernel<KK, LL>), grid, block, smem, stream, a); } while (0)
#endif
  if (Kz == 21) DPC_ZB_LAUNCH(21, false); else DPC_ZB_LAUNCH(11, false);
#undef DPC_ZB_LAUNCH
  return DPC_OK;
}

// the non-lean fast backward exists for the 64^3 grid only
static inline bool dpc_conv_z_bwd_fast_general_ok(int V) { return V == 64; }
