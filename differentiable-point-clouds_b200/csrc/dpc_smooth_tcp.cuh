// Persistent, TMA-staged, warp-specialised pipelines over the tensor-core smoothing passes
// (included by dpc_smooth_tc.cuh, which holds the Toeplitz operand, the 3xTF32 split and the MMA helpers).
//
// One CTA per SM, 13 warps (at most four per scheduler: 128 registers per thread), tiles strided over the grid:
//   warps 0-3   producers, thread = operand row m (TMEM lane m): raw tile from the staging ring in smem ->
//               [clip / DRC gradient] -> 3xTF32 split -> A operand planes in TENSOR MEMORY (tcgen05.st)
//   warps 4-11  consumers, two threads per row (column halves): accumulator (TMEM) -> registers -> epilogue -> global
//   warp  12    one elected lane: issues the MMAs of a tile as soon as its operand is published (A from TMEM, the
//               Toeplitz operand from smem), and keeps the staging ring full (TMA, DPC_TCP_NS tiles ahead)
// Everything is double-buffered in TMEM (512 columns: 2 x accumulators, 2 x A planes), so the producers of tile
// i+1, the MMAs of tile i and the consumers of tile i-1 run concurrently.
//
// How it got here -- each step measured on B200 (scripts/mma_bench.py, scripts/tcp_trace.py, profiles/r01_i_*):
//   * single-tile CTAs (dpc_smooth_tc.cuh) are latency-bound: every phase of a CTA is exposed, 2 CTAs per SM;
//   * LDG-fed producers expose the load latency (one tile in flight per SM): 3.2 us per tile;
//   * a `lane == 0` issue branch costs ~72 cycles per tcgen05.mma (R2UR / ELECT per instruction); elect.sync in a
//     warp-uniform branch lets the compiler emit the 24 MMAs of a tile back to back from uniform registers;
//   * a depth-pass tile fetched as 64 bulk copies of 512 B keeps the TMA unit busy for 3600 cycles (~56 per copy),
//     more than the whole rest of the tile: one cp.async.bulk.tensor.2d box {128 floats, 64 levels} replaces them;
//   * with both operands in shared memory the pipeline is smem-bandwidth-bound (per tile 32 KiB staged in, 32 KiB
//     read back, 64 KiB of hi/lo planes written, 144 KiB read by the MMAs: ~3500 cycles); with A in tensor memory the
//     MMAs read only the 2 KiB Toeplitz slice per instruction and the producers do not store to smem at all;
//   * an issuing lane that also does producer work serialises the two (2300-2900 cycles per tile): own warp.
// smem (1024-aligned): Toeplitz operand 32 KiB | xy transpose buffers 2 x 32 KiB | staging ring NS x 32 KiB.
#pragma once
#include <cuda.h>      // CUtensorMap and its enums (types only; the encoder is fetched through the runtime)

#define DPC_TCP_THREADS 416
#define DPC_TCP_NPROD 128                  // producer threads (warps 0-3)
#define DPC_TCP_ISSUER 12                  // issuing / loading warp
#define DPC_TCP_NS 4                       // staging slots (raw tiles in flight towards HBM)
#define DPC_TCP_T_OFF 0u                   // Toeplitz operand (hi 16 KiB | lo 16 KiB)
#define DPC_TCP_X_OFF 32768u               // xy: two 32 KiB transpose buffers
#define DPC_TCP_S_OFF 98304u               // staging ring
#define DPC_TCP_SLOT 32768u
#define DPC_TCP_SMEM_BYTES (98304 + DPC_TCP_NS * 32768)
// TMEM columns (512 allocated): accumulators D1[s] = 64 s, D2[s] = 128 + 64 s (xy only); A operand planes of stage s:
// hi = 256 + 128 s, lo = 320 + 128 s (row m in lane m, K element k in column + k)
#define DPC_TCP_D1(s) ((uint32_t)(s) * 64u)
#define DPC_TCP_D2(s) (128u + (uint32_t)(s) * 64u)
#define DPC_TCP_AHI(s) (256u + (uint32_t)(s) * 128u)
#define DPC_TCP_ALO(s) (320u + (uint32_t)(s) * 128u)

// diagnostics: per-tile timeline of CTA 0 (clock64 at pipeline hand-offs), enabled by dpc_debug_set(9, 1)
__device__ long long dpc_tcp_trace[16 * 16];
__device__ int dpc_tcp_trace_on = 0;
__device__ unsigned long long dpc_tcp_cta_ns[3 * 160];    // per CTA: globaltimer at entry, after setup + grid dependency, at exit
DPC_DEV unsigned long long dpc_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define DPC_TR(ev) do { if (trace && i < 16) dpc_tcp_trace[(ev) * 16 + i] = clock64(); } while (0)

struct DpcTcpBars {
  uint64_t sfull[DPC_TCP_NS];   // TMA -> producers: raw tile landed in staging slot
  uint64_t sfree[DPC_TCP_NS];   // 4 producer warps -> loader: slot read, may be refilled
  uint64_t opfull[2];    // 4 producer warps -> issuer: A planes of stage s complete
  uint64_t done[2];      // tcgen05.commit of the (first) GEMM of stage s: A planes read, accumulator D1[s] ready
  uint64_t accfree[2];   // depth pass: 8 consumer warps -> issuer: accumulator s drained
  uint64_t opfull2[2];   // xy: 8 consumer warps -> issuer: transposed A planes complete
  uint64_t done2[2];     // xy: second GEMM finished, D2[s] ready
};

DPC_DEV void dpc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dpc_tc_s32(bar)) : "memory");
}
DPC_DEV void dpc_named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// all lanes of a warp are done with a resource (staging slot read / tcgen05.ld or tcgen05.st complete): one arrival per warp
DPC_DEV void dpc_tcp_warp_arrive(uint64_t* bar) {
  dpc_tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) dpc_mbar_arrive(bar);
}
// one thread: post the byte count of the copies that will complete on `bar` (its single arrival)
DPC_DEV void dpc_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dpc_tc_s32(bar)), "r"(bytes) : "memory");
}
// one thread: the box of a 2-D tensor map at (c0 = inner, c1 = outer coordinate) -> shared memory
DPC_DEV void dpc_tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  const unsigned d = dpc_tc_s32(smem_dst), b = dpc_tc_s32(bar);
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(d), "l"(reinterpret_cast<uint64_t>(map)), "r"(b), "r"(c0), "r"(c1) : "memory");
}
// a thread's 32 operand values (row = its TMEM lane, K half h) -> A planes of stage s (not yet waited for)
DPC_DEV void dpc_tcp_put_a(uint32_t tmem, int s, int h, const float* v) {
  const uint32_t lane_t = tmem + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16);
  dpc_tc_split_st32(lane_t + DPC_TCP_AHI(s) + 32u * (uint32_t)h, lane_t + DPC_TCP_ALO(s) + 32u * (uint32_t)h, v);
}

// ---- cross-CTA hand-over inside a persistent kernel (per-sample completion counters): the consumer warps of a CTA count
// their finished tile stores in shared memory (release at CTA scope), ONE thread with no stores of its own turns that into
// a GPU-scope release (fence + atomic on the sample's counter in global memory) and the readers acquire-poll it.
DPC_DEV unsigned dpc_ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
DPC_DEV unsigned dpc_ld_acquire_cta_shared(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(dpc_tc_s32(p)) : "memory");
  return v;
}
DPC_DEV void dpc_red_release_cta_shared(unsigned* p, unsigned v) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(dpc_tc_s32(p)), "r"(v) : "memory");
}
// The taps as launch parameters when the host knows them (sigma is a host value in the reference's schedule): the
// prologue then builds the Toeplitz operand from the constant bank instead of waiting ~1 us for a global load that
// misses L2 -- the prologue of these 226 KB CTAs cannot overlap the previous kernel, so it is on the critical path.
struct DpcTcpTaps { float t[DPC_MAX_TAPS + 1]; int valid; };

#define DPC_TCP_SETUP(TAPS, KK, PL, REV)                                            \
  extern __shared__ __align__(1024) unsigned char dpc_tcp_dsm[];                    \
  __shared__ __align__(8) DpcTcpBars B;                                             \
  dpc_kt_mark(kt_id, 0);                                                            \
  const bool cta_trace = (dpc_tcp_trace_on != 0) && threadIdx.x == 0 && blockIdx.x < 160; \
  if (cta_trace) dpc_tcp_cta_ns[3 * blockIdx.x] = dpc_globaltimer();                \
  __shared__ uint32_t tmem_slot;                                                    \
  unsigned char* sm = dpc_tcp_dsm;                                                  \
  const uint32_t sbase = dpc_tc_s32(sm);                                            \
  if (sbase & 1023u) __trap();                                                      \
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                    \
  const int step = (int)gridDim.x;                                                  \
  const bool trace = (blockIdx.x == 0) && (dpc_tcp_trace_on != 0);                  \
  (void)trace; (void)lane;                                                          \
  if (warp == 0) dpc_tc_alloc(&tmem_slot, 512);                                     \
  if (tid == 0) {                                                                   \
    for (int q = 0; q < DPC_TCP_NS; ++q) { dpc_mbar_init(&B.sfull[q], 1); dpc_mbar_init(&B.sfree[q], 4); } \
    for (int q = 0; q < 2; ++q) {                                                   \
      dpc_mbar_init(&B.opfull[q], 4); dpc_mbar_init(&B.done[q], 1); dpc_mbar_init(&B.accfree[q], 8); \
      dpc_mbar_init(&B.opfull2[q], 8); dpc_mbar_init(&B.done2[q], 1);               \
    }                                                                               \
  }                                                                                 \
  {                                                                                 \
    float* tp_hi = reinterpret_cast<float*>(sm + DPC_TCP_S_OFF);      /* prologue scratch inside the staging ring */ \
    if (ht.valid) {                                                                 \
      for (int i_ = tid; i_ < 192; i_ += DPC_TCP_THREADS) {                         \
        const int j_ = i_ - 64;                                                     \
        const float t_ = (j_ >= 0 && j_ < (KK)) ? ht.t[(REV) ? ((KK) - 1 - j_) : j_] : 0.0f; \
        const float th_ = dpc_tc_hi(t_);                                            \
        tp_hi[i_] = th_;                                                            \
        tp_hi[192 + i_] = dpc_tc_hi(t_ - th_);                                      \
      }                                                                             \
      __syncthreads();                                                              \
      dpc_tc_build_toeplitz_from(sm + DPC_TCP_T_OFF, tp_hi, tp_hi + 192, PL);       \
    } else {                                                                        \
      dpc_tc_build_toeplitz(sm + DPC_TCP_T_OFF, tp_hi, tp_hi + 192, TAPS, KK, PL, REV); \
    }                                                                               \
  }                                                                                 \
  dpc_fence_proxy_async();                                                          \
  dpc_tc_fence_before();                                                            \
  __syncthreads();                                                                  \
  dpc_tc_fence_after();                                                             \
  const uint32_t tmem = tmem_slot;                                                  \
  dpc_grid_dep_sync();                                                              \
  dpc_kt_mark(kt_id, 1);                                                            \
  if (cta_trace) dpc_tcp_cta_ns[3 * blockIdx.x + 1] = dpc_globaltimer()

#define DPC_TCP_TEARDOWN()                                                          \
  dpc_tc_fence_before();                                                            \
  __syncthreads();                                                                  \
  if (cta_trace) dpc_tcp_cta_ns[3 * blockIdx.x + 2] = dpc_globaltimer();            \
  dpc_kt_mark(kt_id, 3);                                                            \
  if (warp == 0) { __syncwarp(); dpc_tc_dealloc(tmem, 512); }

// staging slot <- the 64 depth levels x 128 rays of depth-pass tile (b, image rows 2t, 2t+1): one TMA box; one thread.
// zmap: 2-D view of the grid, inner dimension the 4096 (y, x) positions of a level, outer the B*64 (b, level) rows.
DPC_DEV void dpc_tcp_fill_z(unsigned char* sm, DpcTcpBars& B, const CUtensorMap* zmap, int slot, int tile) {
  dpc_mbar_expect_tx(&B.sfull[slot], DPC_TCP_SLOT);
  dpc_tma_load_2d(sm + DPC_TCP_S_OFF + (uint32_t)slot * DPC_TCP_SLOT, zmap, (tile & 31) * 128, (tile >> 5) * 64, &B.sfull[slot]);
}

// issuing / loading lane of the depth-pass pipelines
DPC_DEV void dpc_tcp_issuer_z(unsigned char* sm, uint32_t sbase, uint32_t tmem, DpcTcpBars& B, const CUtensorMap* zmap,
                              int ntiles, int step, bool trace) {
  for (int j = 0; j < DPC_TCP_NS; ++j) { const int t = (int)blockIdx.x + j * step; if (t < ntiles) dpc_tcp_fill_z(sm, B, zmap, j, t); }
  int i = 0, slot = 0, sph = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
    const int s = i & 1, k = i >> 1;
    dpc_mbar_wait(&B.opfull[s], k & 1);
    DPC_TR(4);
    if (k >= 1) dpc_mbar_wait(&B.accfree[s], (k - 1) & 1);
    dpc_tc_fence_after();
    dpc_tc_issue_ts(tmem + DPC_TCP_AHI(s), tmem + DPC_TCP_ALO(s), sbase + DPC_TCP_T_OFF, tmem + DPC_TCP_D1(s), &B.done[s]);
    DPC_TR(5);
    const int nt = tile + DPC_TCP_NS * step;        // the producers released this tile's slot before they published
    if (nt < ntiles) { dpc_mbar_wait(&B.sfree[slot], sph); dpc_tcp_fill_z(sm, B, zmap, slot, nt); }
    if (++slot == DPC_TCP_NS) { slot = 0; sph ^= 1; }
  }
}

// ------------------------------------------------------------------------------ depth pass, forward
// tile = (sample b, image rows 2t, 2t+1): 128 rays x 64 levels.
template <int MODE, bool HAS_S>
__global__ void __launch_bounds__(DPC_TCP_THREADS, 1)
dpc_tcp_conv_z_fwd_kernel(const __grid_constant__ DpcConvZArgs a, const __grid_constant__ CUtensorMap zmap, int ntiles,
                          const __grid_constant__ DpcTcpTaps ht) {
  constexpr int V = 64, Vz = 64;
  constexpr bool CLAMPU = (MODE == DPC_PROJ_DRC);
  __shared__ __align__(8) float2 comb[128];
  constexpr int kt_id = DPC_KT_Z_F;
  DPC_TCP_SETUP(a.taps, a.K, a.pl, a.rev);
  if (warp < 4) {
    // ---------------- producers: thread = ray
    const int m = tid;
    int i = 0, slot = 0, sph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
      const int s = i & 1, k = i >> 1;
      if (tid == 32) DPC_TR(0);
      dpc_mbar_wait(&B.sfull[slot], sph);
      if (tid == 32) DPC_TR(1);
      if (k >= 1) dpc_mbar_wait(&B.done[s], (k - 1) & 1);     // the MMAs of tile i-2 have read the A planes of stage s
      if (tid == 32) DPC_TR(2);
      const float* stg = reinterpret_cast<const float*>(sm + DPC_TCP_S_OFF + (uint32_t)slot * DPC_TCP_SLOT) + m;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[32];
#pragma unroll
        for (int z = 0; z < 32; ++z) v[z] = stg[(32 * h + z) * 128];
        dpc_tcp_put_a(tmem, s, h, v);
      }
      dpc_tcp_warp_arrive(&B.sfree[slot]);
      dpc_tc_wait_st();
      dpc_tcp_warp_arrive(&B.opfull[s]);
      if (tid == 32) DPC_TR(3);
      if (++slot == DPC_TCP_NS) { slot = 0; sph ^= 1; }
    }
  } else if (warp < DPC_TCP_ISSUER) {
    // ---------------- consumers
    const int c = tid - DPC_TCP_NPROD, m = c & 127, ch = c >> 7;
    const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
      const int s = i & 1, k = i >> 1;
      const int b = tile >> 5, y = (tile & 31) * 2 + (m >> 6), x = m & 63;
      const float sc = HAS_S ? a.scale[b] : 1.0f;
      dpc_mbar_wait(&B.done[s], k & 1);
      if (tid == DPC_TCP_NPROD) DPC_TR(6);
      dpc_tc_fence_after();
      float r[32];
      dpc_tc_ld32(tmem + DPC_TCP_D1(s) + (uint32_t)(ch * 32) + ((uint32_t)((warp & 3) * 32) << 16), r);
      dpc_tc_wait_ld();
      dpc_tcp_warp_arrive(&B.accfree[s]);
      const size_t ray = ((size_t)b * Vz * V + y) * V + x;
      float* vout = a.vox_out + ray + (size_t)(32 * ch) * V * V;
      float T0 = 1.f, T1 = 1.f, S0 = 0.f, S1 = 0.f, mx = -INFINITY;
      uint32_t mw = 0u;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float v = r[j];
        if (HAS_S) {
          const float t = __fmul_rn(v, sc);
          v = __saturatef(t);
          if (v == t) mw |= 1u << j;
        }
        vout[(size_t)j * V * V] = v;
        if (MODE == DPC_PROJ_MAX) {
          mx = fmaxf(mx, v);
        } else if (MODE != DPC_PROJ_NONE) {
          const float u = CLAMPU ? fminf(fmaxf(v, D.lo), D.hi) : v;
          if (j < 16) { float p = u * T0; T0 -= p; if (j == 0 && ch == 0) p *= D.c0; S0 += p; }
          else { const float p = u * T1; T1 -= p; S1 += p; }
        }
      }
      if (HAS_S && a.mask2_out) a.mask2_out[(((size_t)b * V + y) * V + x) * 2 + ch] = mw;
      if (MODE != DPC_PROJ_NONE) {
        // halves of a ray meet through smem: proj = S_lo + T_lo * S_hi (same association as the FFMA2 kernel)
        const float Sh = fmaf(T0, S1, S0), Th = T0 * T1;
        if (ch == 1) comb[m] = (MODE == DPC_PROJ_MAX) ? make_float2(mx, 0.f) : make_float2(Th, Sh);
        dpc_named_bar(1, 256);
        if (ch == 0) {
          const float2 hi = comb[m];
          const float out = (MODE == DPC_PROJ_MAX) ? fmaxf(mx, hi.x) : fmaf(Th, hi.y, Sh);
          const int yo = a.flip_y ? (V - 1 - y) : y;
          a.proj[((size_t)b * V + yo) * V + x] = out;
        }
        dpc_named_bar(1, 256);      // comb[] is reused by the next tile
      }
      if (tid == DPC_TCP_NPROD) DPC_TR(7);
    }
  } else {
    if (dpc_elect_one()) dpc_tcp_issuer_z(sm, sbase, tmem, B, &zmap, ntiles, step, trace);
    __syncwarp();
  }
  DPC_TCP_TEARDOWN();
}

// ------------------------------------------------------------------------------ depth pass, backward, lean
// (DRC silhouette gradient only, occupancy scale present; see dpc_conv_z64_bwd_kernel for the quotient form)
// Roles are mirrored here: the work is in front of the GEMM (two sweeps over the ray for the gradient of every
// level), the back end only stores.  So EIGHT producer warps (thread = (ray, depth half); the two partial products of
// a ray meet through smem) and FOUR consumer warps (thread = ray, all 64 levels).
// MODE = DPC_PROJ_DRC (quotient form) or DPC_PROJ_MAX (the levels that attain the ray's maximum share g equally, TF
// _MaxGrad; max and tie count of the two half rays meet through smem).
template <int MODE>
__global__ void __launch_bounds__(DPC_TCP_THREADS, 1)
dpc_tcp_conv_z_bwd_lean_kernel(const __grid_constant__ DpcConvZBwdArgs a, const __grid_constant__ CUtensorMap zmap, int ntiles,
                               const __grid_constant__ DpcTcpTaps ht) {
  constexpr int V = 64, Vz = 64;
  __shared__ float pp[2][2][128];
  __shared__ unsigned char pc[2][2][128];      // tie counts of the half rays (<= 32); bytes: static smem is nearly full
  constexpr int kt_id = DPC_KT_Z_B;
  DPC_TCP_SETUP(a.taps, a.K, a.pl, a.rev);
  // barrier arrival counts differ from the forward's: 8 producer warps, 4 consumer warps
  if (tid == 0) {
    for (int q = 0; q < DPC_TCP_NS; ++q) dpc_mbar_init(&B.sfree[q], 8);
    for (int q = 0; q < 2; ++q) { dpc_mbar_init(&B.opfull[q], 8); dpc_mbar_init(&B.accfree[q], 4); }
  }
  __syncthreads();
  if (warp < 8) {
    // ---------------- producers: forward voxels of a half ray -> DRC gradient -> A planes
    const int m = tid & 127, h = tid >> 7;
    const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
    struct Extra { float gp, sc; uint32_t wbits; };
    auto fetch = [&](int tile) {
      Extra e;
      const int b = tile >> 5, y = (tile & 31) * 2 + (m >> 6), x = m & 63;
      const int yo = a.flip_y ? (V - 1 - y) : y;
      e.gp = a.g_proj[((size_t)b * V + yo) * V + x];
      e.wbits = a.mask2[(((size_t)b * V + y) * V + x) * 2 + h];
      e.sc = a.scale[b];
      return e;
    };
    Extra en = fetch(blockIdx.x < (unsigned)ntiles ? (int)blockIdx.x : 0);
    int i = 0, slot = 0, sph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
      const int s = i & 1, k = i >> 1;
      const Extra e = en;
      if (tile + step < ntiles) en = fetch(tile + step);      // the next tile's per-ray scalars, one iteration ahead
      dpc_mbar_wait(&B.sfull[slot], sph);
      const float* stg = reinterpret_cast<const float*>(sm + DPC_TCP_S_OFF + (uint32_t)slot * DPC_TCP_SLOT) + (32 * h) * 128 + m;
      float v[32];
#pragma unroll
      for (int z = 0; z < 32; ++z) v[z] = stg[z * 128];
      dpc_tcp_warp_arrive(&B.sfree[slot]);
      const float gp = e.gp, sc = e.sc;
      uint32_t wbits = e.wbits;
      const float inv_s = (sc != 0.0f) ? 1.0f / sc : 0.0f;
      float dsv = 0.0f;
      if (MODE == DPC_PROJ_MAX) {
        float m0 = fmaxf(v[0], v[1]), m1 = fmaxf(v[2], v[3]);
#pragma unroll
        for (int z = 4; z < 32; z += 2) { m0 = fmaxf(m0, v[z]); m1 = fmaxf(m1, v[z + 1]); }
        pp[i & 1][h][m] = fmaxf(m0, m1);
        dpc_named_bar(2, 256);
        const float mx = fmaxf(pp[i & 1][0][m], pp[i & 1][1][m]);
        int cnt = 0;
#pragma unroll
        for (int z = 0; z < 32; ++z) cnt += (v[z] == mx) ? 1 : 0;
        pc[i & 1][h][m] = (unsigned char)cnt;
        dpc_named_bar(2, 256);          // pp / pc [i & 1] are rewritten two tiles later, behind that tile's barriers
        const float share = gp / (float)((int)pc[i & 1][0][m] + (int)pc[i & 1][1][m]);
#pragma unroll
        for (int z = 0; z < 32; ++z) {
          const float vv = v[z];
          float dv = (vv == mx) ? share : 0.0f;
          if (!(wbits & 1u)) dv = 0.0f;
          wbits >>= 1;
          dsv = fmaf(dv, vv, dsv);
          v[z] = dv * sc;
        }
      } else {
        float P0 = 1.0f, P1 = 1.0f, P2 = 1.0f, P3 = 1.0f;
#pragma unroll
        for (int z = 0; z < 32; z += 4) {
          P0 *= 1.0f - fminf(fmaxf(v[z + 0], D.lo), D.hi);
          P1 *= 1.0f - fminf(fmaxf(v[z + 1], D.lo), D.hi);
          P2 *= 1.0f - fminf(fmaxf(v[z + 2], D.lo), D.hi);
          P3 *= 1.0f - fminf(fmaxf(v[z + 3], D.lo), D.hi);
        }
        pp[i & 1][h][m] = (P0 * P1) * (P2 * P3);
        dpc_named_bar(2, 256);          // pp[i & 1] is rewritten two tiles later, behind the next tile's barrier
        const float gT = gp * (pp[i & 1][0][m] * pp[i & 1][1][m]);
#pragma unroll
        for (int z = 0; z < 32; ++z) {
          const float vv = v[z];
          const float u = fminf(fmaxf(vv, D.lo), D.hi);
          float dv = __fdividef(gT, 1.0f - u);
          if (z == 0 && h == 0) dv = fmaf(gp, D.c0 - 1.0f, dv);
          if ((u != vv) || !(wbits & 1u)) dv = 0.0f;
          wbits >>= 1;
          dsv = fmaf(dv, vv, dsv);
          v[z] = dv * sc;
        }
      }
      if (a.d_scale_part) {
        const float w = dpc_warp_sum(dsv * inv_s);
        if (lane == 0) a.d_scale_part[(size_t)tile * 8 + warp] = w;
      } else if (a.d_scale) {
        const float w = dpc_warp_sum(dsv * inv_s);
        if (lane == 0) atomicAdd(a.d_scale + (tile >> 5), w);
      }
      if (k >= 1) dpc_mbar_wait(&B.done[s], (k - 1) & 1);
      dpc_tcp_put_a(tmem, s, h, v);
      dpc_tc_wait_st();
      dpc_tcp_warp_arrive(&B.opfull[s]);
      if (++slot == DPC_TCP_NS) { slot = 0; sph ^= 1; }
    }
  } else if (warp < DPC_TCP_ISSUER) {
    // ---------------- consumers: thread = ray, dL/d(xy-smoothed) of all 64 levels out
    const int m = tid - 256;
    if (blockIdx.x == 0) {     // the splat backward's accumulation targets (it runs two kernels from now)
      for (int q = 0; q < 4; ++q)
        if (a.zero.p[q]) for (int e = m; e < a.zero.n[q]; e += 128) a.zero.p[q][e] = 0.0f;
    }
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
      const int s = i & 1, k = i >> 1;
      const int b = tile >> 5, y = (tile & 31) * 2 + (m >> 6), x = m & 63;
      dpc_mbar_wait(&B.done[s], k & 1);
      dpc_tc_fence_after();
      float r[64];
      const uint32_t taddr = tmem + DPC_TCP_D1(s) + ((uint32_t)((warp & 3) * 32) << 16);
      dpc_tc_ld32(taddr, r);
      dpc_tc_ld32(taddr + 32, r + 32);
      dpc_tc_wait_ld();
      dpc_tcp_warp_arrive(&B.accfree[s]);
      float* dout = a.d_in + ((size_t)b * Vz * V + y) * V + x;
#pragma unroll
      for (int j = 0; j < 64; ++j) dout[(size_t)j * V * V] = r[j];
    }
  } else {
    if (dpc_elect_one()) dpc_tcp_issuer_z(sm, sbase, tmem, B, &zmap, ntiles, step, false);
    __syncwarp();
  }
  DPC_TCP_TEARDOWN();
}

// ------------------------------------------------------------------------------ x and y passes
// tile = two depth slices = 128 rows (slice, y) of 64 x.  xymap: 2-D view [B*64*64 rows][64 x], box {32 x, 128 rows},
// SWIZZLE_128B: a staging slot holds two x-halves of [128 rows][128 B], 16-byte chunk c of row r at c ^ (r & 7), so
// that thread = row reads its chunks without bank conflicts.  Tile i uses TMEM stage s = i & 1:
//   producers: row -> [clip, pass bits] -> split -> A[s]  -> GEMM 1 -> D1[s]
//   consumers: D1[s] row (slice, y') -> transposed into smem X[s] (raw fp32, same chunk swizzle) -> row (slice, x)
//              half -> split -> A[s] again -> GEMM 2 -> D2[s] -> store (slice, y, x), one tile late so that it
//              overlaps GEMM 2 of the next tile.
// PDRAIN (knob 2): the four PRODUCER warps store the finished tiles instead of the consumers.  The consumers' chain per
// tile is  D1 -> transpose through smem -> split -> A planes  and then  D2 -> global; the producers only split one
// staged row each.  With the store moved over, a stage's critical path is  P -> GEMM 1 -> transpose -> GEMM 2  and the
// store of tile i-2 runs under it (accfree[s] tells the issuer that D2[s] has been drained before GEMM 2 of tile i).
template <bool CLIP, bool MOUT, bool MIN, bool PDRAIN>
__global__ void __launch_bounds__(DPC_TCP_THREADS, 1)
dpc_tcp_conv_xy_kernel(const __grid_constant__ DpcConvXY64Args a, const __grid_constant__ CUtensorMap xymap, int K, int pl, int ntiles,
                       const __grid_constant__ DpcTcpTaps ht, int ns) {
  // ns: staging slots in use (at most 4)
  constexpr int V = 64;
  constexpr int kt_id = MIN ? DPC_KT_XY_B : DPC_KT_XY_F;
  DPC_TCP_SETUP(a.taps_x, K, pl, a.rev);
  if (PDRAIN) {           // accfree[s]: 4 producer warps -> issuer, D2[s] drained
    if (tid == 0) { dpc_mbar_init(&B.accfree[0], 4); dpc_mbar_init(&B.accfree[1], 4); }
    __syncthreads();
  }
  if (warp < 4) {
    // ---------------- producers: thread = row (slice, y)
    const int m = tid;
    // PDRAIN: store of tile j (D2[s] ready: done2 of tile j has been waited for): thread = (slice, x), all 64 rows y
    auto pdrain = [&](int j, int tile) {
      const int s = j & 1;
      const size_t base = (size_t)tile * (2 * V * V);
      const int sl = m >> 6, rx = m & 63;
      dpc_tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t mw = 0xffffffffu;
        if (MIN) mw = a.mask_in[(base >> 5) + (size_t)sl * 128 + 2 * (32 * h + lane) + ((m >> 5) & 1)];
        float r[32];
        dpc_tc_ld32(tmem + DPC_TCP_D2(s) + (uint32_t)(h * 32) + ((uint32_t)(warp * 32) << 16), r);
        dpc_tc_wait_ld();
        float* dst = a.out + base + (size_t)sl * V * V + (size_t)(32 * h) * V + rx;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          float v = r[q];
          if (MIN) {
            const uint32_t w = __shfl_sync(DPC_FULL, mw, q);
            if (!((w >> lane) & 1u)) v = 0.0f;
          }
          dst[q * V] = v;
        }
      }
      dpc_tcp_warp_arrive(&B.accfree[s]);
    };
    int i = 0, slot = 0, sph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
      const int s = i & 1, k = i >> 1;
      const size_t base = (size_t)tile * (2 * V * V);
      dpc_mbar_wait(&B.sfull[slot], sph);
      if (k >= 1) dpc_mbar_wait(&B.done2[s], (k - 1) & 1);     // GEMM 2 of tile i-2 has read the A planes of stage s
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const unsigned char* row = sm + DPC_TCP_S_OFF + (uint32_t)slot * DPC_TCP_SLOT + (uint32_t)h * 16384u + (uint32_t)m * 128u;
        float v[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 f = *reinterpret_cast<const float4*>(row + ((q ^ (m & 7)) << 4));
          v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
        }
        if (MOUT) {       // these 32 voxels are exactly one word of the clip-pass bit plane
          uint32_t word = 0u;
#pragma unroll
          for (int j = 0; j < 32; ++j) word |= (v[j] >= 0.0f && v[j] <= 1.0f) ? (1u << j) : 0u;
          a.mask_out[(base >> 5) + 2 * m + h] = word;
        }
        if (CLIP) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = dpc_clip01(v[j]);
        }
        dpc_tcp_put_a(tmem, s, h, v);
      }
      dpc_tcp_warp_arrive(&B.sfree[slot]);
      dpc_tc_wait_st();
      dpc_tcp_warp_arrive(&B.opfull[s]);
      if (PDRAIN && k >= 1) pdrain(i - 2, tile - 2 * step);    // its done2 was waited for above
      if (++slot == ns) { slot = 0; sph ^= 1; }
    }
    if (PDRAIN) {          // the last two tiles of this CTA
      for (int j = (i >= 2 ? i - 2 : 0); j < i; ++j) {
        dpc_mbar_wait(&B.done2[j & 1], (j >> 1) & 1);
        pdrain(j, (int)blockIdx.x + j * step);
      }
    }
  } else if (warp < DPC_TCP_ISSUER) {
    // ---------------- consumers
    const int c = tid - DPC_TCP_NPROD, m = c & 127, h = c >> 7;
    const int sl = m >> 6, rx = m & 63;                 // D1: (slice, y' = rx); D2: (slice, x = rx)
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    // store of tile j (its D2 is ready once done2 fires): thread = (slice, x), rows y = 32 h .. 32 h + 31
    auto drain = [&](int j, int tile) {
      const int s = j & 1, k = j >> 1;
      const size_t base = (size_t)tile * (2 * V * V);
      uint32_t mw = 0xffffffffu;      // saved clip mask of the rows this thread writes: lane l fetches row 32 h + l
      if (MIN) mw = a.mask_in[(base >> 5) + (size_t)sl * 128 + 2 * (32 * h + lane) + ((m >> 5) & 1)];
      dpc_mbar_wait(&B.done2[s], k & 1);
      dpc_tc_fence_after();
      float r[32];
      dpc_tc_ld32(tmem + DPC_TCP_D2(s) + (uint32_t)(h * 32) + lane_addr, r);
      dpc_tc_wait_ld();
      dpc_tc_fence_before();
      float* dst = a.out + base + (size_t)sl * V * V + (size_t)(32 * h) * V + rx;
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        float v = r[q];
        if (MIN) {
          const uint32_t w = __shfl_sync(DPC_FULL, mw, q);
          if (!((w >> lane) & 1u)) v = 0.0f;
        }
        dst[q * V] = v;
      }
    };
    int i = 0, prev_tile = -1;
    for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
      const int s = i & 1, k = i >> 1;
      unsigned char* X = sm + DPC_TCP_X_OFF + (uint32_t)s * 32768u;
      dpc_mbar_wait(&B.done[s], k & 1);
      dpc_tc_fence_after();
      float r[32];
      dpc_tc_ld32(tmem + DPC_TCP_D1(s) + (uint32_t)(h * 32) + lane_addr, r);
      dpc_tc_wait_ld();
      {
        // element (row m2 = (slice, x = 32 h + j), column y' = rx) of X; (m2 & 7) == (j & 7)
        unsigned char* colb = X + (uint32_t)(sl * 64 + 32 * h) * 256u + (uint32_t)(rx >> 5) * 128u + (uint32_t)((rx & 3) * 4);
        const int cc = (rx & 31) >> 2;
#pragma unroll
        for (int j = 0; j < 32; ++j) *reinterpret_cast<float*>(colb + (uint32_t)j * 256u + (uint32_t)((cc ^ (j & 7)) << 4)) = r[j];
      }
      dpc_named_bar(1, 256);
      {
        const unsigned char* row = X + (uint32_t)m * 256u + (uint32_t)h * 128u;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 f = *reinterpret_cast<const float4*>(row + ((q ^ (m & 7)) << 4));
          r[4 * q] = f.x; r[4 * q + 1] = f.y; r[4 * q + 2] = f.z; r[4 * q + 3] = f.w;
        }
      }
      dpc_tcp_put_a(tmem, s, h, r);                    // A planes of stage s: GEMM 1 of this tile has finished reading them
      dpc_tc_wait_st();
      dpc_tcp_warp_arrive(&B.opfull2[s]);
      if (!PDRAIN && prev_tile >= 0) drain(i - 1, prev_tile);      // overlaps GEMM 2 of tile i
      prev_tile = tile;
    }
    if (!PDRAIN && prev_tile >= 0) drain(i - 1, prev_tile);
  } else {
    // ---------------- issuing / loading lane
    if (dpc_elect_one()) {
      auto fill = [&](int slot, int tile) {
        unsigned char* dst = sm + DPC_TCP_S_OFF + (uint32_t)slot * DPC_TCP_SLOT;
        dpc_mbar_expect_tx(&B.sfull[slot], DPC_TCP_SLOT);
        dpc_tma_load_2d(dst, &xymap, 0, tile * 128, &B.sfull[slot]);
        dpc_tma_load_2d(dst + 16384, &xymap, 32, tile * 128, &B.sfull[slot]);
      };
      for (int j = 0; j < ns; ++j) { const int t = (int)blockIdx.x + j * step; if (t < ntiles) fill(j, t); }
      int slot = 0, sph = 0;
      auto gemm1 = [&](int j, int tile) {      // D1[s] drained: implied by done2[s] of tile j-2, which the producers waited for
        const int s = j & 1, k = j >> 1;
        dpc_mbar_wait(&B.opfull[s], k & 1);
        dpc_tc_fence_after();
        dpc_tc_issue_ts(tmem + DPC_TCP_AHI(s), tmem + DPC_TCP_ALO(s), sbase + DPC_TCP_T_OFF, tmem + DPC_TCP_D1(s), &B.done[s]);
        const int nt = tile + ns * step;
        if (nt < ntiles) { dpc_mbar_wait(&B.sfree[slot], sph); fill(slot, nt); }
        if (++slot == ns) { slot = 0; sph ^= 1; }
      };
      auto gemm2 = [&](int j) {                // D2[s] drained: every consumer stored tile j-2 before it published tile j
        const int s = j & 1, k = j >> 1;
        dpc_mbar_wait(&B.opfull2[s], k & 1);
        if (PDRAIN && k >= 1) dpc_mbar_wait(&B.accfree[s], (k - 1) & 1);     // the producers have stored tile j-2 (D2[s])
        dpc_tc_fence_after();
        dpc_tc_issue_ts(tmem + DPC_TCP_AHI(s), tmem + DPC_TCP_ALO(s), sbase + DPC_TCP_T_OFF, tmem + DPC_TCP_D2(s), &B.done2[s]);
      };
      int i = 0;
      if ((int)blockIdx.x < ntiles) gemm1(0, blockIdx.x);
      for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
        if (tile + step < ntiles) gemm1(i + 1, tile + step);
        gemm2(i);
      }
    }
    __syncwarp();
  }
  DPC_TCP_TEARDOWN();
}

// ------------------------------------------------------------------------------ MMA micro-benchmark (diagnostics)
// One CTA per SM; thread 0 issues `nmma` tcgen05.mma (M128 N64 K8 tf32, SS) + commit and waits, `reps` times.
// out[3*cta + 0] = cycles per repetition, [1] = cycles spent issuing, [2] = cycles waiting after the issue.
// spin != 0: every other thread of the CTA polls the same mbarrier (as the pipeline's waiters do).
DPC_DEV void dpc_tc_mma_idesc(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__global__ void __launch_bounds__(512, 1) dpc_tc_mma_bench_kernel(long long* out, int reps, int nmma, int spin, int M, int N) {
  extern __shared__ __align__(1024) unsigned char dpc_tcp_dsm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  unsigned char* sm = dpc_tcp_dsm;
  const uint32_t sbase = dpc_tc_s32(sm);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) stop = 0;
  for (int i = tid; i < 196608 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.0f;
  if (warp == 0) dpc_tc_alloc(&tmem_slot, 512);
  if (tid == 0) dpc_mbar_init(&bar, 1);
  dpc_fence_proxy_async();
  dpc_tc_fence_before();
  __syncthreads();
  dpc_tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = 0x10u | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (warp == 0 && dpc_elect_one()) {
    long long t_issue = 0, t_wait = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const long long a0 = clock64();
      const uint64_t dA = dpc_tc_desc(sbase), dT = dpc_tc_desc(sbase + 65536u);
      for (int k = 0; k < nmma; ++k) {
        const uint32_t kk = (uint32_t)(k & 7);
        const uint64_t off = (uint64_t)(((kk >> 2) * 16384u + (kk & 3) * 32u) >> 4);
        const uint64_t offb = (uint64_t)(((kk >> 2) * 8192u + (kk & 3) * 32u) >> 4);
        if (spin == 3) {
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                       ::"r"(tmem), "r"(tmem + 256u + 8u * kk), "l"(dT + offb), "r"(idesc), "r"((uint32_t)(k > 0)) : "memory");
        } else {
          dpc_tc_mma_idesc(tmem, dA + off, dT + offb, idesc, k > 0);
        }
      }
      dpc_tc_commit(&bar);
      const long long a1 = clock64();
      dpc_mbar_wait(&bar, r & 1);
      dpc_tc_fence_after();
      const long long a2 = clock64();
      t_issue += a1 - a0; t_wait += a2 - a1;
    }
    const long long t1 = clock64();
    out[3 * blockIdx.x + 0] = (t1 - t0) / reps;
    out[3 * blockIdx.x + 1] = t_issue / reps;
    out[3 * blockIdx.x + 2] = t_wait / reps;
    stop = 1;
  } else if (spin) {
    // poll the barrier the way pipeline waiters do (try_wait, retry), until thread 0 is finished
    const unsigned ba = dpc_tc_s32(&bar);
    unsigned r = 0;
    while (!stop) {
      unsigned ok;
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(ba), "r"(r & 1u) : "memory");
      if (ok) ++r; else if (spin == 2) __nanosleep(200);
    }
  }
  dpc_tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); dpc_tc_dealloc(tmem, 512); }
}
