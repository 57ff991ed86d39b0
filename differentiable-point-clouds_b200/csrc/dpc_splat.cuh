// K1 / K1b: fused camera transform + validity + trilinear splat, and its backward.
//
// Replaces ~55 element-wise TF kernels, 9-17 boolean_mask compactions (dynamic shapes, host
// syncs), eight zero-filled dense grids + eight scatter_nd + add_n of the reference
// (point_cloud.py:60-136,157-216; quaternion.py:96-117) by ONE pass: a point tile is staged
// into shared memory by the TMA engine (cp.async.bulk, mbarrier completion), transformed in
// registers with the reference's exact fp32 op order, written back coalesced as tr_pc, and its
// 8 corner weights are added to the single grid with warp-aggregated REDG (v2 where aligned).
//
// Two generations of kernels live here.  dpc_splat_fwd_kernel / dpc_splat_bwd_kernel: one tile of points per CTA
// (TMA-staged); they serve the rgb grid and a backward without the forward's tr_pc.
// dpc_splat_fwd_warp_kernel / dpc_splat_bwd_warp_kernel (further down, the default): every warp an independent worker
// over several 32-point tiles, next tile prefetched with cp.async -- twice as fast in the backward.
#pragma once
#include "dpc_math.cuh"
#ifndef DPC_EMU
#include <cooperative_groups.h>
#endif

#define DPC_SPLAT_THREADS 256
#define DPC_SPLAT_MAX_PPT 4
// Points per thread (PPT) is a template parameter: more points per thread = more independent
// work per warp and fewer CTAs paying the tile-load latency; fewer = more warps per SM.
// dpc_splat_ppt() picks the instantiation (tunable for experiments, see dpc_debug_set).

struct DpcSplatArgs {
  const float* pc; const float* pose; const float* trans; const float* focal; const float* rgb;
  int pose_kind; float focal_const; float cam_dist;
  int B, N, Vz, V;
  float* tr_pc; float* vox; float* vox_rgb; int32_t* idx_out; uint8_t* valid_out;
  int early;   // the stream predecessor only zeroes `vox` (fused path): stage + transform + tr_pc before the grid dependency
  int red4;    // x pairs go out as one 16-byte red.v4 when they do not straddle a 4-voxel group (experiment knob 11)
  // f-2, point dropout consumed by the load stage (point_cloud.py:293-319): sel != NULL => pc is [B,N_src,3] and point i
  // of sample b is pc[b, sel[b*N + i]]; the points that were dropped are never read.  Outputs are [B,N,..] as usual.
  const int32_t* sel; int N_src;
  unsigned* zero_u32; int n_zero;   // optional: words zeroed by CTA (0, 0) behind the grid dependency (per-sample counters of the fused x/y + depth kernel)
};

// Stage `n` points (3n floats) into smem: one TMA bulk copy when the 16-byte rules allow it
// (issued by thread 0, completion on `bar`), else a cooperative copy.  Returns with the tile
// visible to every thread.  While the copy is in flight thread 32 prepares the sample's camera.
// sel != NULL: `src` is the SAMPLE's first point and the tile is gathered through the index list instead (f-2).
DPC_DEV void dpc_stage_points(float* tile, uint64_t* bar, const float* src, int n, DpcPose* pose_sm,
                              const float* pose, int pose_kind, const float* trans, const float* focal,
                              float focal_const, float cam_dist, int b, const int32_t* sel = nullptr) {
  const unsigned bytes = (unsigned)n * 12u;
  if (sel) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float* q = src + (size_t)sel[i] * 3;
      tile[i * 3 + 0] = q[0]; tile[i * 3 + 1] = q[1]; tile[i * 3 + 2] = q[2];
    }
    if (threadIdx.x == 32) dpc_pose_load(*pose_sm, pose, pose_kind, trans, focal, focal_const, cam_dist, b);
    __syncthreads();
    return;
  }
  const bool bulk_ok = ((((uintptr_t)src) & 15u) == 0) && ((bytes & 15u) == 0);
  if (bulk_ok && threadIdx.x == 0) dpc_mbar_init(bar, 1);
  __syncthreads();
  if (bulk_ok) {
    if (threadIdx.x == 0) dpc_bulk_load(tile, src, bytes, bar);
  } else {
    for (int i = threadIdx.x; i < n * 3; i += blockDim.x) tile[i] = src[i];
  }
  if (threadIdx.x == 32) dpc_pose_load(*pose_sm, pose, pose_kind, trans, focal, focal_const, cam_dist, b);
  if (bulk_ok) dpc_mbar_wait(bar, 0);
  __syncthreads();
}

// Write `n` points (3n floats) from smem to global: TMA bulk store when aligned.
DPC_DEV void dpc_unstage_points(float* dst, const float* tile, int n) {
  const unsigned bytes = (unsigned)n * 12u;
  const bool bulk_ok = ((((uintptr_t)dst) & 15u) == 0) && ((bytes & 15u) == 0);
  if (bulk_ok) {
    if (threadIdx.x == 0) dpc_bulk_store(dst, tile, bytes);
  } else {
    for (int i = threadIdx.x; i < n * 3; i += blockDim.x) dst[i] = tile[i];
  }
}

template <int DPC_SPLAT_PPT>
#ifndef DPC_EMU
__global__ void __launch_bounds__(DPC_SPLAT_THREADS)
#else
static void
#endif
dpc_splat_fwd_kernel(DpcSplatArgs a) {
  constexpr int DPC_SPLAT_TILE = DPC_SPLAT_THREADS * DPC_SPLAT_PPT;
  __shared__ __align__(128) float tile[DPC_SPLAT_TILE * 3];
  __shared__ __align__(8) uint64_t bar;
  __shared__ DpcPose pose_sm;
  const int b = blockIdx.y;
  const int p_first = blockIdx.x * DPC_SPLAT_TILE;
  const int n = min(DPC_SPLAT_TILE, a.N - p_first);
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int V = a.V, Vz = a.Vz;

  dpc_kt_mark(DPC_KT_SPLAT_F, 0);
  dpc_ph_mark(0, 0);
  dpc_grid_dep_trigger();
  if (!a.early) { dpc_grid_dep_wait(); dpc_kt_mark(DPC_KT_SPLAT_F, 1); }
  if (a.zero_u32 && blockIdx.x == 0 && blockIdx.y == 0 && !a.early)
    for (int i = tid; i < a.n_zero; i += DPC_SPLAT_THREADS) a.zero_u32[i] = 0u;
  if (a.sel) dpc_stage_points(tile, &bar, a.pc + (size_t)b * a.N_src * 3, n, &pose_sm, a.pose, a.pose_kind, a.trans, a.focal,
                              a.focal_const, a.cam_dist, b, a.sel + (size_t)b * a.N + p_first);
  else dpc_stage_points(tile, &bar, a.pc + ((size_t)b * a.N + p_first) * 3, n, &pose_sm,
                        a.pose, a.pose_kind, a.trans, a.focal, a.focal_const, a.cam_dist, b);
  const DpcPose P = pose_sm;
  dpc_ph_mark(0, 1);

  float z[DPC_SPLAT_PPT], y[DPC_SPLAT_PPT], x[DPC_SPLAT_PPT];
#pragma unroll
  for (int j = 0; j < DPC_SPLAT_PPT; ++j) {
    const int i = j * DPC_SPLAT_THREADS + tid;
    z[j] = y[j] = x[j] = 0.0f;
    if (i < n) {
      DpcCamPoint cam;
      dpc_transform_point(P, tile[i * 3 + 0], tile[i * 3 + 1], tile[i * 3 + 2], z[j], y[j], x[j], cam);
    }
  }
  dpc_ph_mark(0, 2);
  if (a.tr_pc) {
    __syncthreads();  // everyone has read its points
#pragma unroll
    for (int j = 0; j < DPC_SPLAT_PPT; ++j) {
      const int i = j * DPC_SPLAT_THREADS + tid;
      if (i < n) { tile[i * 3 + 0] = z[j]; tile[i * 3 + 1] = y[j]; tile[i * 3 + 2] = x[j]; }
    }
    dpc_fence_proxy_async();
    __syncthreads();
    dpc_unstage_points(a.tr_pc + ((size_t)b * a.N + p_first) * 3, tile, n);
  }
  // the grid is zero (and everything older than the zeroing kernel complete) from here on
  dpc_ph_mark(0, 3);
  if (a.early) { dpc_grid_dep_wait(); dpc_kt_mark(DPC_KT_SPLAT_F, 1); }
  dpc_ph_mark(0, 4);

#pragma unroll
  for (int j = 0; j < DPC_SPLAT_PPT; ++j) {
    const int i = j * DPC_SPLAT_THREADS + tid;
    const bool live = i < n;
    DpcCell c = dpc_cell(z[j], y[j], x[j], Vz, V);
    c.valid = c.valid && live;
    if (live) {
      const size_t pi = (size_t)b * a.N + p_first + i;
      if (a.idx_out) { a.idx_out[pi * 3 + 0] = c.iz; a.idx_out[pi * 3 + 1] = c.iy; a.idx_out[pi * 3 + 2] = c.ix; }
      if (a.valid_out) a.valid_out[pi] = c.valid ? 1 : 0;
    }
    if (!a.vox) continue;   // uniform

    // corner weights, reference association: (rr[k].z * rr[j].y) * rr[i].x  (point_cloud.py:99)
    const float wz[2] = {__fsub_rn(1.0f, c.rz), c.rz};
    const float wy[2] = {__fsub_rn(1.0f, c.ry), c.ry};
    const float wx[2] = {__fsub_rn(1.0f, c.rx), c.rx};
    float w[8];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int ii = 0; ii < 2; ++ii)
          w[k * 4 + jj * 2 + ii] = c.valid ? __fmul_rn(__fmul_rn(wz[k], wy[jj]), wx[ii]) : 0.0f;

    const int base = (c.iz * V + c.iy) * V + c.ix;

    if (a.rgb && c.valid) {
      // 3-channel grid (point_cloud.py:111-118): plain per-lane reductions (non-default path).
      const float* col = a.rgb + ((size_t)b * a.N + p_first + i) * 3;
      const float cr = col[0], cg = col[1], cb = col[2];
      float* g3 = a.vox_rgb + (size_t)b * Vz * V * V * 3;
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
          for (int ii = 0; ii < 2; ++ii) {
            if (c.iz + k < Vz && c.iy + jj < V && c.ix + ii < V) {
              const float ww = w[k * 4 + jj * 2 + ii];
              float* q = g3 + (size_t)(base + (k * V + jj) * V + ii) * 3;
              dpc_red_add(q + 0, __fmul_rn(ww, cr));
              dpc_red_add(q + 1, __fmul_rn(ww, cg));
              dpc_red_add(q + 2, __fmul_rn(ww, cb));
            }
          }
    }

    // ---- warp aggregation: lanes whose points share a base voxel fold their 8 weights into the
    // lowest lane of the group, so a clustered cloud (decoder init, stddev 0.025) issues one
    // reduction per corner per group instead of 32 serialised same-address ones.
    const int key = c.valid ? base : (-1 - lane);
    const unsigned peers = __match_any_sync(DPC_FULL, key);
    const int cnt = __popc(peers);
    const int maxcnt = __reduce_max_sync(DPC_FULL, cnt);
    if (maxcnt > 1) {
      float s[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) s[q] = w[q];
      unsigned rem = peers & ~(1u << lane);
      for (int it = 1; it < maxcnt; ++it) {
        const int src = rem ? (__ffs(rem) - 1) : lane;
        rem &= rem - 1;
        const bool take = it < cnt;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float v = __shfl_sync(DPC_FULL, w[q], src);
          if (take) s[q] += v;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) w[q] = s[q];
    }
    const bool leader = c.valid && (lane == (__ffs(peers) - 1));
    if (leader) {
      float* g = a.vox + (size_t)b * Vz * V * V + base;
      const bool pair_ok = ((V & 1) == 0) && ((c.ix & 1) == 0) && ((((uintptr_t)a.vox) & 7u) == 0);
      const int o4 = c.ix & 3;
      const bool quad_ok = a.red4 && ((V & 3) == 0) && (o4 != 3) && ((((uintptr_t)a.vox) & 15u) == 0);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (c.iz + k >= Vz) continue;  // only for a coordinate of exactly +0.5 (weight is 0): TF-GPU drops it
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          if (c.iy + jj >= V) continue;
          float* row = g + (k * V + jj) * V;
          if (quad_ok) {
            const float w0 = w[k * 4 + jj * 2 + 0], w1 = w[k * 4 + jj * 2 + 1];
            dpc_red_add4(row - o4, o4 == 0 ? w0 : 0.0f, o4 == 0 ? w1 : (o4 == 1 ? w0 : 0.0f),
                         o4 == 1 ? w1 : (o4 == 2 ? w0 : 0.0f), o4 == 2 ? w1 : 0.0f);
          } else if (pair_ok) {
            dpc_red_add2(row, w[k * 4 + jj * 2 + 0], w[k * 4 + jj * 2 + 1]);
          } else {
            dpc_red_add(row, w[k * 4 + jj * 2 + 0]);
            if (c.ix + 1 < V) dpc_red_add(row + 1, w[k * 4 + jj * 2 + 1]);
          }
        }
      }
    }
  }
  dpc_ph_mark(0, 5);
  dpc_kt_mark(DPC_KT_SPLAT_F, 3);
}

// ------------------------------------------------------------------------------ backward
struct DpcSplatBwdArgs {
  const float* pc; const float* pose; const float* trans; const float* focal; const float* rgb;
  int pose_kind; float focal_const; float cam_dist; int rgb_stop_grad;
  int B, N, Vz, V;
  const float* d_vox; const float* d_vox_rgb; const float* d_tr_pc_in;
  float* d_pc; float* d_pose; float* d_trans; float* d_focal; float* d_rgb;
  int early;   // transform before the grid dependency (experiment knob 14)
  int gather4; // 16-byte gathers of x pairs (experiment knob 11)
  int gather_cg; // gathers through ld.global.cg (L2 only, no L1 line allocated per scattered miss; experiment knob 22)
  int stagger_ns; // experiment knob 23: every other CTA sleeps this long at entry (are the phases of a wave in lock-step?)
  // per-warp partial sums of dL/dscale left by the depth-pass backward ([B, n_part]); CTA (0, b) folds them into
  // d_scale_out[b], so the fused backward needs neither atomics on d_scale nor a launch that zeroes it
  const float* d_scale_part; int n_part; float* d_scale_out;
  // f-2 (see DpcSplatArgs): pc and d_pc are [B,N_src,3], addressed through sel[B,N]; d_pc is ZEROED by the launcher and
  // the selected rows are written (the indices of a sample are distinct: plain stores)
  const int32_t* sel; int N_src;
  const float* tr_pc;   // the forward's tr_pc [B,N,3] or NULL: lets dpc_splat_bwd_warp_kernel start a tile's gathers before its transform
};

// gathers of dL/d(raw): through the read-only path normally; past L1 (ld.global.cg) when the producer kernel is still
// running on other samples (co-resident mode)
DPC_DEV float dpc_ld_gather(const float* p, bool coherent) {
#ifndef DPC_EMU
  return coherent ? __ldcg(p) : __ldg(p);
#else
  (void)coherent; return *p;
#endif
}
DPC_DEV float4 dpc_ld_gather4(const float4* p, bool coherent) {
#ifndef DPC_EMU
  return coherent ? __ldcg(p) : __ldg(p);
#else
  (void)coherent; return *p;
#endif
}

// The 8 corner values of dL/d(raw) around a cell, for a grid whose rows are 16-byte aligned (V % 4 == 0): per (z, y) row
// ONE 16-byte load of the 4-voxel group that holds ix, plus one scalar load of the neighbour for the lanes whose x pair
// straddles two groups (ix % 4 == 3).  All loads are UNCONDITIONAL (a lane without a valid row, or without a straddling
// pair, reads the sample's first voxels instead) and every one is issued before anything consumes a result.  The first form of
// this code guarded each load with `if (inb)`: nvcc then wrapped every load in its own divergence region and reused one
// destination quad, so the four gathers of a point became four DEPENDENT L2 round trips (cuobjdump: LDG.E.128 R28 x 3
// with the selects in between) -- the "gather issue 12.9 us" of profiles/r01_m_splat_phases.txt.
template <bool COHERENT>
DPC_DEV void dpc_gather_corners(const float* dv, const DpcCell& c, int Vz, int V, float* dw) {
  const int base = (c.iz * V + c.iy) * V + c.ix;
  const int o4 = c.ix & 3;
  float4 q[4];
  float e[4];
  bool ok[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int k = r >> 1, jj = r & 1;
    ok[r] = c.valid && (c.iz + k < Vz) && (c.iy + jj < V);
    const float* rp = ok[r] ? dv + base - o4 + (k * V + jj) * V : dv;
    q[r] = dpc_ld_gather4(reinterpret_cast<const float4*>(rp), COHERENT);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int k = r >> 1, jj = r & 1;
    const bool need = ok[r] && (o4 == 3) && (c.ix + 1 < V);
    const float* ep = need ? dv + base + 1 + (k * V + jj) * V : dv;
    e[r] = dpc_ld_gather(ep, COHERENT);       // unconditional as well: lanes that do not need it all read dv[0]
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float w0 = o4 == 0 ? q[r].x : (o4 == 1 ? q[r].y : (o4 == 2 ? q[r].z : q[r].w));
    const float w1 = o4 == 0 ? q[r].y : (o4 == 1 ? q[r].z : (o4 == 2 ? q[r].w : ((c.ix + 1 < V) ? e[r] : 0.0f)));
    dw[r * 2 + 0] = ok[r] ? w0 : 0.0f;
    dw[r * 2 + 1] = ok[r] ? w1 : 0.0f;
  }
}

// The same 8 corner values with PREDICATED loads (inline PTX): no divergence regions, so all loads of a point are issued
// back to back into distinct registers (one round trip), and a lane whose predicate is off generates no wavefront --
// the load count of the guarded code (four 16-byte loads per valid lane, four scalar loads on the straddling lanes
// only) with the issue pattern of dpc_gather_corners.
DPC_DEV void dpc_gather_corners_pred(const float* dv, const DpcCell& c, int Vz, int V, float* dw) {
  const int base = (c.iz * V + c.iy) * V + c.ix;
  const int o4 = c.ix & 3;
  const bool straddle = (o4 == 3) && (c.ix + 1 < V);
  float4 q[4];
  float e[4];
  bool ok[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int k = r >> 1, jj = r & 1;
    ok[r] = c.valid && (c.iz + k < Vz) && (c.iy + jj < V);
    const float* rp = dv + base - o4 + (k * V + jj) * V;
    q[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    e[r] = 0.0f;
#ifndef DPC_EMU
    asm("{\n\t.reg .pred pq;\n\tsetp.ne.u32 pq, %5, 0;\n\t@pq ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "+f"(q[r].x), "+f"(q[r].y), "+f"(q[r].z), "+f"(q[r].w) : "l"(rp), "r"((unsigned)ok[r]));
    asm("{\n\t.reg .pred pq;\n\tsetp.ne.u32 pq, %2, 0;\n\t@pq ld.global.nc.f32 %0, [%1];\n\t}"
        : "+f"(e[r]) : "l"(rp + 4), "r"((unsigned)(ok[r] && straddle)));
#else
    if (ok[r]) q[r] = *reinterpret_cast<const float4*>(rp);
    if (ok[r] && straddle) e[r] = rp[4];
#endif
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    dw[r * 2 + 0] = o4 == 0 ? q[r].x : (o4 == 1 ? q[r].y : (o4 == 2 ? q[r].z : q[r].w));
    dw[r * 2 + 1] = o4 == 0 ? q[r].y : (o4 == 1 ? q[r].z : (o4 == 2 ? q[r].w : e[r]));
  }
}

// MINB: minimum resident CTAs per SM the kernel is compiled for (0 = unconstrained: ptxas settles at 61 registers, four 256-thread CTAs per SM);
// INDEP: 0 = four guarded 16-byte gathers, a second (scalar) path for the lanes whose x pair straddles two 4-voxel groups;
// 1 = independent, un-guarded loads (dpc_gather_corners); 2 = ONE guarded path: per row the group that holds ix and, for
// the straddling lanes only, the next group (same lines touched as 0, without the second code path every warp walks);
// 3 = the loads of 2 as predicated PTX loads: independent, no divergence regions (dpc_gather_corners_pred).
template <int DPC_SPLAT_PPT, int NT, int MINB = 0, int INDEP = 0>
#ifndef DPC_EMU
__global__ void __launch_bounds__(NT, MINB)
#else
static void
#endif
dpc_splat_bwd_kernel(DpcSplatBwdArgs a) {
  constexpr int DPC_SPLAT_TILE = NT * DPC_SPLAT_PPT;
  __shared__ __align__(128) float tile[DPC_SPLAT_TILE * 3];
  __shared__ __align__(8) uint64_t bar;
  __shared__ DpcPose pose_sm;
  __shared__ float red[NT / 32][12];
  const int b = blockIdx.y;
  const int p_first = blockIdx.x * DPC_SPLAT_TILE;
  const int n = min(DPC_SPLAT_TILE, a.N - p_first);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int V = a.V, Vz = a.Vz;

  // The points and the camera are inputs of the forward (nothing in front of this kernel writes them), so they are
  // staged and transformed while the x/y pass of the backward is still draining; the gathers wait for it.
  dpc_kt_mark(DPC_KT_SPLAT_B, 0);
  dpc_ph_mark(1, 0);
  dpc_grid_dep_trigger();
#if defined(DPC_EXPERIMENTS) && !defined(DPC_EMU)
  if (a.stagger_ns > 0) {
    const unsigned phase = (blockIdx.x + blockIdx.y) & 3u;
    if (phase) __nanosleep((unsigned)a.stagger_ns * phase);
  }
#endif
  if (!a.early) { dpc_grid_dep_wait(); dpc_kt_mark(DPC_KT_SPLAT_B, 1); }
  if (a.sel) dpc_stage_points(tile, &bar, a.pc + (size_t)b * a.N_src * 3, n, &pose_sm, a.pose, a.pose_kind, a.trans, a.focal,
                              a.focal_const, a.cam_dist, b, a.sel + (size_t)b * a.N + p_first);
  else dpc_stage_points(tile, &bar, a.pc + ((size_t)b * a.N + p_first) * 3, n, &pose_sm,
                        a.pose, a.pose_kind, a.trans, a.focal, a.focal_const, a.cam_dist, b);
  const DpcPose P = pose_sm;
  dpc_ph_mark(1, 1);

  float acc[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) acc[q] = 0.0f;
  float p0[DPC_SPLAT_PPT], p1[DPC_SPLAT_PPT], p2[DPC_SPLAT_PPT];
  DpcCamPoint cam[DPC_SPLAT_PPT];
  DpcCell cell[DPC_SPLAT_PPT];
  float dw[DPC_SPLAT_PPT][8];
  const float* dv = a.d_vox ? a.d_vox + (size_t)b * Vz * V * V : nullptr;

  // pass 1: recompute the transform (same code as forward => same cell) and issue all corner
  // gathers of all four points before anything consumes them
#pragma unroll
  for (int j = 0; j < DPC_SPLAT_PPT; ++j) {
    const int i = j * NT + tid;
    p0[j] = p1[j] = p2[j] = 0.0f;
    float z = 0.f, y = 0.f, x = 0.f;
    cam[j].xs = cam[j].ys = 0.f; cam[j].zs = 1.f;
    if (i < n) {
      p0[j] = tile[i * 3 + 0]; p1[j] = tile[i * 3 + 1]; p2[j] = tile[i * 3 + 2];
      dpc_transform_point(P, p0[j], p1[j], p2[j], z, y, x, cam[j]);
    }
    cell[j] = dpc_cell(z, y, x, Vz, V);
    cell[j].valid = cell[j].valid && (i < n);
  }
  dpc_ph_mark(1, 2);
  if (a.early) { dpc_grid_dep_wait(); dpc_kt_mark(DPC_KT_SPLAT_B, 1); }
  dpc_ph_mark(1, 3);
  if (a.d_scale_part && blockIdx.x == 0 && warp == NT / 32 - 1) {
    float v = 0.0f;
    for (int q = lane; q < a.n_part; q += 32) v += a.d_scale_part[(size_t)b * a.n_part + q];
    v = dpc_warp_sum(v);
    if (lane == 0) a.d_scale_out[b] = v;
  }
  // An uncoalesced warp load costs one L1 wavefront per lane, and the gathers are what this kernel waits for: the x
  // pair of a row comes in ONE 16-byte load whenever it does not straddle a 4-voxel group (3 of 4 points): 5
  // wavefronts per point on average instead of 8.
  const bool coherent = a.gather_cg != 0;
  const bool quad = a.gather4 && dv && ((V & 3) == 0) && ((((uintptr_t)dv) & 15u) == 0);
#pragma unroll
  for (int j = 0; j < DPC_SPLAT_PPT; ++j) {
    const int base = (cell[j].iz * V + cell[j].iy) * V + cell[j].ix;
    const int o4 = cell[j].ix & 3;
    if (INDEP == 1 && quad) {
      dpc_gather_corners<false>(dv, cell[j], Vz, V, dw[j]);
    } else if (INDEP == 3 && quad) {
      dpc_gather_corners_pred(dv, cell[j], Vz, V, dw[j]);
    } else if (INDEP == 2 && quad) {
      const bool straddle = (o4 == 3) && (cell[j].ix + 1 < V);
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const bool inb = cell[j].valid && (cell[j].iz + k < Vz) && (cell[j].iy + jj < V);
          const float* rp = dv + base + (k * V + jj) * V - o4;
          float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
          float nx = 0.0f;
          if (inb) q4 = dpc_ld_gather4(reinterpret_cast<const float4*>(rp), coherent);
          if (inb && straddle) nx = dpc_ld_gather(rp + 4, coherent);
          dw[j][k * 4 + jj * 2 + 0] = o4 == 0 ? q4.x : (o4 == 1 ? q4.y : (o4 == 2 ? q4.z : q4.w));
          dw[j][k * 4 + jj * 2 + 1] = o4 == 0 ? q4.y : (o4 == 1 ? q4.z : (o4 == 2 ? q4.w : nx));
        }
    } else if (quad && o4 != 3) {        // ix + 1 < V is implied
      // (the four loads below end up as four DEPENDENT round trips -- nvcc wraps each in its own divergence region and
      // reuses one destination quad; the independent form, dpc_gather_corners, was measured SLOWER in this kernel at full
      // occupancy: 23.8 vs 20.2 us, profiles/r02_b_timeline_sep.txt.  Neither the chain nor the L1 miss path bounds this
      // kernel -- its lock-step shape does, see dpc_splat_bwd_warp_kernel and profiles/r02_w_splat_bwd_warp.md)
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const bool inb = cell[j].valid && (cell[j].iz + k < Vz) && (cell[j].iy + jj < V);
          float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (inb) q4 = dpc_ld_gather4(reinterpret_cast<const float4*>(dv + base + (k * V + jj) * V - o4), coherent);
          dw[j][k * 4 + jj * 2 + 0] = o4 == 0 ? q4.x : (o4 == 1 ? q4.y : q4.z);
          dw[j][k * 4 + jj * 2 + 1] = o4 == 0 ? q4.y : (o4 == 1 ? q4.z : q4.w);
        }
    } else {
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
          for (int ii = 0; ii < 2; ++ii) {
            const bool inb = cell[j].valid && dv && (cell[j].iz + k < Vz) && (cell[j].iy + jj < V) && (cell[j].ix + ii < V);
            dw[j][k * 4 + jj * 2 + ii] = inb ? dpc_ld_gather(dv + base + (k * V + jj) * V + ii, coherent) : 0.0f;
          }
    }
  }
  dpc_ph_mark(1, 4);   // gathers issued
  __syncthreads();  // every thread has read its points: the tile can take the results

  // pass 2: weights' derivative, chain rule through the camera
  const bool want_tf = (a.d_trans != nullptr) || (a.d_focal != nullptr);
#pragma unroll
  for (int j = 0; j < DPC_SPLAT_PPT; ++j) {
    const int i = j * NT + tid;
    const bool live = i < n;
    const DpcCell c = cell[j];
    const size_t pi = (size_t)b * a.N + p_first + i;
    float gz = 0.f, gy = 0.f, gx = 0.f;
    if (c.valid && (a.d_vox || a.d_vox_rgb)) {
      const float wz[2] = {1.0f - c.rz, c.rz}, wy[2] = {1.0f - c.ry, c.ry}, wx[2] = {1.0f - c.rx, c.rx};
      const int base = (c.iz * V + c.iy) * V + c.ix;
      const float* dv3 = a.d_vox_rgb ? a.d_vox_rgb + (size_t)b * Vz * V * V * 3 : nullptr;
      float cr = 0.f, cg = 0.f, cb = 0.f, dr = 0.f, dg = 0.f, db = 0.f;
      if (dv3) { const float* col = a.rgb + pi * 3; cr = col[0]; cg = col[1]; cb = col[2]; }
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
          for (int ii = 0; ii < 2; ++ii) {
            if (c.iz + k < Vz && c.iy + jj < V && c.ix + ii < V) {
              float dwc = dw[j][k * 4 + jj * 2 + ii];  // dL/dw of this corner
              if (dv3) {
                const size_t off = (size_t)(base + (k * V + jj) * V + ii) * 3;
                const float e0 = __ldg(dv3 + off + 0), e1 = __ldg(dv3 + off + 1), e2 = __ldg(dv3 + off + 2);
                const float ww = (wz[k] * wy[jj]) * wx[ii];
                dr += ww * e0; dg += ww * e1; db += ww * e2;
                if (!a.rgb_stop_grad) dwc += e0 * cr + e1 * cg + e2 * cb;
              }
              gz += (k ? dwc : -dwc) * (wy[jj] * wx[ii]);
              gy += (jj ? dwc : -dwc) * (wz[k] * wx[ii]);
              gx += (ii ? dwc : -dwc) * (wz[k] * wy[jj]);
            }
          }
      gz *= (float)(Vz - 1); gy *= (float)(V - 1); gx *= (float)(V - 1);  // d grid / d coordinate
      if (a.d_rgb) { a.d_rgb[pi * 3 + 0] = dr; a.d_rgb[pi * 3 + 1] = dg; a.d_rgb[pi * 3 + 2] = db; }
    } else if (a.d_rgb && live) {
      a.d_rgb[pi * 3 + 0] = 0.f; a.d_rgb[pi * 3 + 1] = 0.f; a.d_rgb[pi * 3 + 2] = 0.f;
    }
    if (live) {
      if (a.d_tr_pc_in) {
        gz += a.d_tr_pc_in[pi * 3 + 0]; gy += a.d_tr_pc_in[pi * 3 + 1]; gx += a.d_tr_pc_in[pi * 3 + 2];
      }
      // chain rule applied unconditionally: an invalid but finite point gets exact zeros, a NaN
      // point propagates NaN into the pose gradient exactly as TF's autodiff does (0 * NaN).
      float d0, d1, d2;
      dpc_transform_point_bwd(P, p0[j], p1[j], p2[j], cam[j], gz, gy, gx, d0, d1, d2, acc, want_tf);
      if (a.d_pc) { tile[i * 3 + 0] = d0; tile[i * 3 + 1] = d1; tile[i * 3 + 2] = d2; }
    }
  }
  dpc_ph_mark(1, 5);   // gathers consumed, chain rule done
  if (a.d_pc && a.sel) {
    __syncthreads();
    const int32_t* sl = a.sel + (size_t)b * a.N + p_first;
    float* dst = a.d_pc + (size_t)b * a.N_src * 3;
    for (int i = tid; i < n; i += NT) {
      float* q = dst + (size_t)sl[i] * 3;
      q[0] = tile[i * 3 + 0]; q[1] = tile[i * 3 + 1]; q[2] = tile[i * 3 + 2];
    }
  } else if (a.d_pc) {
    dpc_fence_proxy_async();
    __syncthreads();
    dpc_unstage_points(a.d_pc + ((size_t)b * a.N + p_first) * 3, tile, n);
  }
  dpc_ph_mark(1, 6);
  dpc_kt_mark(DPC_KT_SPLAT_B, 3);
  if (a.pose_kind == DPC_POSE_NONE) return;
  const bool want_pose = a.d_pose != nullptr, want_t = a.d_trans != nullptr, want_f = a.d_focal != nullptr;
  if (!(want_pose || want_t || want_f)) return;
  // block reduction of the per-sample pose gradients: warp shuffles, then one atomic per CTA.  Only the components
  // somebody asked for are reduced (quaternion pose: 4 for d_pose, +3 translation, +1 focal; 10 shuffle + add
  // instructions per component and thread -- the full 12 were 16 % of this kernel's instructions, ncu r02_p)
  {
    const int q_lo = (a.pose_kind == DPC_POSE_QUAT && !want_pose) ? 4 : 0;
    const int q_hi = (a.pose_kind == DPC_POSE_QUAT) ? (want_f ? 8 : (want_t ? 7 : 4)) : 12;
#pragma unroll
    for (int q = 0; q < 12; ++q) {
      if (q < q_lo || q >= q_hi) { if (lane == 0) red[warp][q] = 0.0f; continue; }      // warp-uniform
      const float v = dpc_warp_sum(acc[q]);
      if (lane == 0) red[warp][q] = v;
    }
  }
  __syncthreads();
  if (tid < 12) {
    float v = 0.f;
    for (int wgi = 0; wgi < NT / 32; ++wgi) v += red[wgi][tid];
    red[0][tid] = v;
  }
  __syncthreads();
  if (tid == 0) {
    if (a.pose_kind == DPC_POSE_QUAT) {
      if (want_pose) {
        float dq[4];
        dpc_quat_norm_bwd(P, red[0], dq);
        for (int q = 0; q < 4; ++q) atomicAdd(a.d_pose + b * 4 + q, dq[q]);
      }
      if (want_t) for (int q = 0; q < 3; ++q) atomicAdd(a.d_trans + b * 3 + q, red[0][4 + q]);
      if (want_f) atomicAdd(a.d_focal + b, red[0][7]);
    } else if (want_pose) {
      // dL/dE = K^T dL/dM: rows 1,2 scaled by f; row 3 of the extrinsic does not reach the output
      for (int k = 0; k < 4; ++k) {
        atomicAdd(a.d_pose + b * 16 + 0 + k, red[0][0 + k]);
        atomicAdd(a.d_pose + b * 16 + 4 + k, red[0][4 + k] * a.focal_const);
        atomicAdd(a.d_pose + b * 16 + 8 + k, red[0][8 + k] * a.focal_const);
      }
    }
  }
}

// ------------------------------------------------------------------------------ backward, software-pipelined form
// Every warp is an independent worker with k tiles of 32 points each (tiles j, j + wps, ... of sample blockIdx.y;
// DPC_SPLAT_WPC warps share a CTA only for the launch rate and the sample's camera), every warp of the grid resident at once.  What the tile-per-CTA kernel above cannot do -- its CTAs walk stage / transform / gather / chain
// rule in lock-step, so an SM alternates between waiting for memory and being issue-bound (scripts/gather_bench.py: the
// gathers alone take 4 us, that kernel 20) -- this one does by prefetching: the cells of the NEXT tile are known from the
// forward's tr_pc (12 B per point, no transform needed), so its corner rows and its points are fetched with cp.async
// (global -> shared, no destination registers, the warp does not wait) while the current tile's transform and chain
// rule are computed.  Rows that are out of range are zero-filled through cp.async's src-size operand.  The copies of a
// tile are issued in four portions spread over the current tile's arithmetic (a burst of 11 scattered LDGSTS backs up
// the SM's load/store queue and the warp then stalls on the address registers it wants to reuse: ncu, r02_w).  The
// pose gradients are accumulated in registers over all tiles of the warp (same sample) and reduced once.
// Preconditions (the launcher falls back to dpc_splat_bwd_kernel otherwise): tr_pc given, no rgb, V % 4 == 0 and a
// 16-byte aligned gradient grid.
#ifndef DPC_EMU
DPC_DEV void dpc_cp_async16(void* smem, const void* g, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(g), "r"(pred ? 16u : 0u) : "memory");
}
DPC_DEV void dpc_cp_async4(void* smem, const void* g, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(g), "r"(pred ? 4u : 0u) : "memory");
}
DPC_DEV void dpc_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_PENDING>
DPC_DEV void dpc_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_PENDING) : "memory"); }
#else
// CPU emulation (tests/emu): the copies complete at once; the zero-fill of the src-size operand is kept
DPC_DEV void dpc_cp_async16(void* smem, const void* g, bool pred) { if (pred) memcpy(smem, g, 16); else memset(smem, 0, 16); }
DPC_DEV void dpc_cp_async4(void* smem, const void* g, bool pred) { if (pred) memcpy(smem, g, 4); else memset(smem, 0, 4); }
DPC_DEV void dpc_cp_async_commit() {}
template <int N_PENDING>
DPC_DEV void dpc_cp_async_wait() {}
#endif

struct DpcSplatBwdWarpSmem {   // per warp
  float4 gq[2][4][32];   // per buffer, per (z, y) row of the cell: the 4-voxel group that holds ix, one per lane
  float ge[2][4][32];    // the voxel behind that group, for the lanes whose x pair straddles two groups
  float pts[2][96];      // the tile's points (AoS, as in global memory); reused for d_pc on the way out
};
#define DPC_SPLAT_WPC 4      // warps per CTA of the software-pipelined splat kernels: the warps are independent workers, the
                             // CTA only exists because 1-warp CTAs launch at ~0.45 per ns (4000 of them: 9 us)

// what a lane needs to fetch the four corner rows of its next cell
struct DpcGatherPlan {
  const float* g0;       // the 4-voxel group of (iz, iy, ix)
  unsigned ok;           // bit r: row r = (k, jj) is inside the grid and the point is valid; bit 4: the x pair straddles
  int base;              // linear index of the cell, -1 for a lane without a valid point: what the plan was made for
};

// SEL: the dropout index list (f-2) is compiled in only where it is used -- the instantiation without it is instruction
// for instruction the kernel that was tuned above (the list's bookkeeping cost 1.2 us when it was a run-time branch).
template <int MINB, bool SEL>
#ifndef DPC_EMU
__global__ void __launch_bounds__(32 * DPC_SPLAT_WPC, MINB)
#else
static void
#endif
dpc_splat_bwd_warp_kernel(DpcSplatBwdArgs a) {
  __shared__ __align__(16) DpcSplatBwdWarpSmem sm_all[DPC_SPLAT_WPC];
  __shared__ DpcPose pose_sm;
  DpcSplatBwdWarpSmem& sm = sm_all[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int V = a.V, Vz = a.Vz, N = a.N;
  const int tiles = (N + 31) >> 5;
  const int wps = gridDim.x * DPC_SPLAT_WPC;
  const float* dv = a.d_vox + (size_t)b * Vz * V * V;
  // f-2 (dropout list): pc / d_pc are [B,N_src,3] and point i of the sample is row sel[b*N + i]; tr_pc stays [B,N,3]
  const int32_t* sel_b = SEL ? a.sel + (size_t)b * N : nullptr;
  const float* pc_b = a.pc + (size_t)b * (SEL ? a.N_src : N) * 3;
  const float* tr_b = a.tr_pc + (size_t)b * N * 3;
  int t = blockIdx.x * DPC_SPLAT_WPC + (threadIdx.x >> 5);
  const bool kt = dpc_kt_enabled();      // read once: a load of the flag behind the grid dependency would sit on the critical path

  int row_next = 0, row_cur = 0;         // dropout list: this lane's source row in the tile being fetched / computed
  auto load_row = [&](int tt) {
    if (!SEL) return;
    const int i = tt * 32 + lane;
    row_next = (i < N) ? __ldg(sel_b + i) : 0;
  };
  auto issue_points = [&](int tt, int buf) {
    if (SEL) {                           // the lane's own point: three 4-byte copies from its row (row_next = sel of tile tt)
      const bool live = tt * 32 + lane < N;
      const float* src = pc_b + (size_t)row_next * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) dpc_cp_async4(&sm.pts[buf][lane * 3 + c], live ? src + c : pc_b, live);
      return;
    }
    const int n3 = min(32, N - tt * 32) * 3;
    const float* src = pc_b + (size_t)tt * 96;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int i = lane + 32 * c;
      dpc_cp_async4(&sm.pts[buf][i], i < n3 ? src + i : pc_b, i < n3);
    }
  };
  auto plan = [&](const DpcCell& c) {      // c.valid already says whether the lane has a point at all
    const int o4 = c.ix & 3;
    DpcGatherPlan g;
    g.base = c.valid ? (c.iz * V + c.iy) * V + c.ix : -1;
    g.g0 = dv + (c.iz * V + c.iy) * V + (c.ix - o4);
    g.ok = 0u;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (c.valid && (c.iz + (r >> 1) < Vz) && (c.iy + (r & 1) < V)) g.ok |= 1u << r;
    if ((o4 == 3) && (c.ix + 1 < V)) g.ok |= 16u;
    return g;
  };
  auto issue_row = [&](const DpcGatherPlan& g, int r, int buf) {
    const bool ok = (g.ok >> r) & 1u, ex = ok && (g.ok & 16u);
    const float* rp = ok ? g.g0 + ((r >> 1) * V + (r & 1)) * V : dv;
    dpc_cp_async16(&sm.gq[buf][r][lane], rp, ok);
    dpc_cp_async4(&sm.ge[buf][r][lane], ex ? rp + 4 : dv, ex);
  };
  // tr_pc of a tile: loaded one iteration before the cell is needed (the raw values stay in flight, the cell is computed late)
  float rz, ry, rx;
  auto load_raw = [&](int tt) {
    const int i = tt * 32 + lane;
    rz = ry = rx = 2.0f;        // outside the grid: an invalid cell
    if (i < N) { rz = __ldg(tr_b + (size_t)i * 3 + 0); ry = __ldg(tr_b + (size_t)i * 3 + 1); rx = __ldg(tr_b + (size_t)i * 3 + 2); }
  };

  // prologue: tr_pc, the points and the camera are forward data (nothing in front of this kernel writes them)
  dpc_kt_mark_if(kt, DPC_KT_SPLAT_B, 0);
  dpc_grid_dep_trigger();
  if (t < tiles) {
    load_raw(t);
    load_row(t);
    issue_points(t, 0);
    row_cur = row_next;
    if (t + wps < tiles) load_row(t + wps);
  }
  if (threadIdx.x == 0) dpc_pose_load(pose_sm, a.pose, a.pose_kind, a.trans, a.focal, a.focal_const, a.cam_dist, b);
  __syncthreads();               // the only CTA-wide step: the camera of the sample
  if (t >= tiles) return;
  DpcGatherPlan gc = plan(dpc_cell(rz, ry, rx, Vz, V));    // the plan the CURRENT tile's corners were requested with
  {
    if (t + wps < tiles) load_raw(t + wps);
    dpc_grid_dep_wait();
    dpc_kt_mark_if(kt, DPC_KT_SPLAT_B, 1);
#pragma unroll
    for (int r = 0; r < 4; ++r) issue_row(gc, r, 0);
  }
  dpc_cp_async_commit();
  if (a.d_scale_part && blockIdx.x == 0 && threadIdx.x < 32) {
    float v = 0.0f;
    for (int q = lane; q < a.n_part; q += 32) v += a.d_scale_part[(size_t)b * a.n_part + q];
    v = dpc_warp_sum(v);
    if (lane == 0) a.d_scale_out[b] = v;
  }

  float acc[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) acc[q] = 0.0f;
  const bool want_tf = (a.d_trans != nullptr) || (a.d_focal != nullptr);
  int buf = 0;
  for (; t < tiles; t += wps, buf ^= 1) {
    const int tn = t + wps;
    const bool more = tn < tiles;                  // warp-uniform
    dpc_cp_async_wait<0>();                        // this tile's points and corners (issued during the previous tile)
    __syncwarp();                                  // ... of every lane
    DpcGatherPlan gn;
    gn.g0 = dv; gn.ok = 0u; gn.base = -1;
    const int row_here = row_cur;                  // dropout list: where this tile's d_pc rows go
    if (more) {
      issue_points(tn, buf ^ 1);                   // (uses row_next = the list entries of tile tn)
      if (SEL) row_cur = row_next;
      gn = plan(dpc_cell(rz, ry, rx, Vz, V));
      issue_row(gn, 0, buf ^ 1);
      if (tn + wps < tiles) { load_raw(tn + wps); load_row(tn + wps); }
    }

    const int i = t * 32 + lane;
    const int n = min(32, N - t * 32);
    const bool live = lane < n;
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, z = 0.f, y = 0.f, x = 0.f;
    DpcCamPoint cam;
    cam.xs = cam.ys = 0.f; cam.zs = 1.f;
    if (live) {
      p0 = sm.pts[buf][lane * 3 + 0]; p1 = sm.pts[buf][lane * 3 + 1]; p2 = sm.pts[buf][lane * 3 + 2];
      dpc_transform_point(pose_sm, p0, p1, p2, z, y, x, cam);
    }
    if (more) issue_row(gn, 1, buf ^ 1);
    DpcCell c = dpc_cell(z, y, x, Vz, V);      // the forward's cell again: same code, same inputs as the prefetch's tr_pc
    c.valid = c.valid && live;
    float gz = 0.f, gy = 0.f, gx = 0.f;
    // The prefetch was aimed by tr_pc, the result must not depend on it: a lane whose recomputed cell is not the one its
    // corners were requested for (tr_pc is not the tr_pc of this call's forward) fetches its rows now, the slow way.
    const bool stale = gc.base != (c.valid ? (c.iz * V + c.iy) * V + c.ix : -1);
    if (c.valid) {
      const int o4 = c.ix & 3;
      const float wz[2] = {1.0f - c.rz, c.rz}, wy[2] = {1.0f - c.ry, c.ry}, wx[2] = {1.0f - c.rx, c.rx};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int k = r >> 1, jj = r & 1;
        float4 q4 = sm.gq[buf][r][lane];
        float e = sm.ge[buf][r][lane];
        if (stale) {
          const DpcGatherPlan pv = plan(c);
          const bool ok = (pv.ok >> r) & 1u;
          const float* rp = pv.g0 + (k * V + jj) * V;
          q4 = ok ? __ldg(reinterpret_cast<const float4*>(rp)) : make_float4(0.f, 0.f, 0.f, 0.f);
          e = (ok && (pv.ok & 16u)) ? __ldg(rp + 4) : 0.0f;
        }
        const float w0 = o4 == 0 ? q4.x : (o4 == 1 ? q4.y : (o4 == 2 ? q4.z : q4.w));
        const float w1 = o4 == 0 ? q4.y : (o4 == 1 ? q4.z : (o4 == 2 ? q4.w : e));
        if (c.iz + k < Vz && c.iy + jj < V) {
          // same expression order as dpc_splat_bwd_kernel (corner ii = 0 before ii = 1)
          gz += (k ? w0 : -w0) * (wy[jj] * wx[0]);
          gy += (jj ? w0 : -w0) * (wz[k] * wx[0]);
          gx += (-w0) * (wz[k] * wy[jj]);
          if (c.ix + 1 < V) {
            gz += (k ? w1 : -w1) * (wy[jj] * wx[1]);
            gy += (jj ? w1 : -w1) * (wz[k] * wx[1]);
            gx += w1 * (wz[k] * wy[jj]);
          }
        }
      }
      gz *= (float)(Vz - 1); gy *= (float)(V - 1); gx *= (float)(V - 1);
    }
    if (more) { issue_row(gn, 2, buf ^ 1); issue_row(gn, 3, buf ^ 1); }
    dpc_cp_async_commit();
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    if (live) {
      const size_t pi = (size_t)b * N + i;
      if (a.d_tr_pc_in) { gz += a.d_tr_pc_in[pi * 3 + 0]; gy += a.d_tr_pc_in[pi * 3 + 1]; gx += a.d_tr_pc_in[pi * 3 + 2]; }
      dpc_transform_point_bwd(pose_sm, p0, p1, p2, cam, gz, gy, gx, d0, d1, d2, acc, want_tf);
    }
    if (SEL && a.d_pc) {                           // rows of a sample's list are distinct: plain stores (the launcher zeroed d_pc)
      if (live) {
        float* dst = a.d_pc + ((size_t)b * a.N_src + (size_t)row_here) * 3;
        dst[0] = d0; dst[1] = d1; dst[2] = d2;
      }
    } else if (a.d_pc) {
      if (live) { sm.pts[buf][lane * 3 + 0] = d0; sm.pts[buf][lane * 3 + 1] = d1; sm.pts[buf][lane * 3 + 2] = d2; }
      __syncwarp();
      float* dst = a.d_pc + ((size_t)b * N + (size_t)t * 32) * 3;
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int q = lane + 32 * cc;
        if (q < n * 3) dst[q] = sm.pts[buf][q];
      }
    }
    __syncwarp();   // the buffer is free for the prefetch of the tile after next
    gc = gn;
  }
  dpc_cp_async_wait<0>();
  dpc_kt_mark_if(kt, DPC_KT_SPLAT_B, 3);
  if (a.pose_kind == DPC_POSE_NONE) return;
  const bool want_pose = a.d_pose != nullptr, want_t = a.d_trans != nullptr, want_f = a.d_focal != nullptr;
  if (!(want_pose || want_t || want_f)) return;
  const int q_lo = (a.pose_kind == DPC_POSE_QUAT && !want_pose) ? 4 : 0;
  const int q_hi = (a.pose_kind == DPC_POSE_QUAT) ? (want_f ? 8 : (want_t ? 7 : 4)) : 12;
  float red[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) red[q] = (q < q_lo || q >= q_hi) ? 0.0f : dpc_warp_sum(acc[q]);
  if (lane == 0) {
    if (a.pose_kind == DPC_POSE_QUAT) {
      if (want_pose) {
        float dq[4];
        dpc_quat_norm_bwd(pose_sm, red, dq);
        for (int q = 0; q < 4; ++q) atomicAdd(a.d_pose + b * 4 + q, dq[q]);
      }
      if (want_t) for (int q = 0; q < 3; ++q) atomicAdd(a.d_trans + b * 3 + q, red[4 + q]);
      if (want_f) atomicAdd(a.d_focal + b, red[7]);
    } else if (want_pose) {
      for (int k = 0; k < 4; ++k) {
        atomicAdd(a.d_pose + b * 16 + 0 + k, red[0 + k]);
        atomicAdd(a.d_pose + b * 16 + 4 + k, red[4 + k] * a.focal_const);
        atomicAdd(a.d_pose + b * 16 + 8 + k, red[8 + k] * a.focal_const);
      }
    }
  }
}

// ------------------------------------------------------------------------------ forward, software-pipelined form
// The forward splat in the shape of dpc_splat_bwd_warp_kernel: independent warps, k tiles of 32 points per warp, every
// warp resident at once, the next tile's points prefetched with cp.async while the current tile is transformed, written
// back as tr_pc and reduced into the grid.  The tile-per-CTA kernel above runs 256 CTAs of 1024 points on 148 SMs (108
// SMs get two CTAs, 40 get one) with 8-16 warps per SM; here every SM gets its 54 tiles and 27 warps.
// Same device functions for the transform and the cell => tr_pc and the voxel indices stay bit-exact.
// ZERO: the kernel also zeroes the grid it reduces into (cooperative launch: every CTA stores its share of zeros right
// after issuing its first loads, transforms its first tile while they drain, then a grid-wide barrier separates the zeros
// from the first reduction) -- the fused forward then needs no memset node in front of it.
// Preconditions (launcher): no rgb, no counters to zero.
template <int MINB, bool ZERO, bool SEL>
#ifndef DPC_EMU
__global__ void __launch_bounds__(32 * DPC_SPLAT_WPC, MINB)
#else
static void
#endif
dpc_splat_fwd_warp_kernel(DpcSplatArgs a) {
  __shared__ __align__(16) float pts_all[DPC_SPLAT_WPC][2][96];
  __shared__ DpcPose pose_sm;
  float (*pts)[96] = pts_all[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int V = a.V, Vz = a.Vz, N = a.N;
  const int tiles = (N + 31) >> 5;
  const int wps = gridDim.x * DPC_SPLAT_WPC;
  // f-2 (dropout list): pc is [B,N_src,3] and point i of the sample is row sel[b*N + i]; the list is written by the
  // kernel in front of this one (dpc_dropout_indices), so with a list everything waits for the grid dependency first
  const int32_t* sel_b = SEL ? a.sel + (size_t)b * N : nullptr;
  const float* pc_b = a.pc + (size_t)b * (SEL ? a.N_src : N) * 3;
  int t = blockIdx.x * DPC_SPLAT_WPC + (threadIdx.x >> 5);
  const bool kt = dpc_kt_enabled();
  int row_next = 0;
  auto load_row = [&](int tt) {
    if (!SEL) return;
    const int i = tt * 32 + lane;
    row_next = (i < N) ? __ldg(sel_b + i) : 0;
  };
  auto issue_points = [&](int tt, int buf) {
    if (SEL) {                           // the lane's own point: three 4-byte copies from its row (row_next = sel of tile tt)
      const bool lv = tt * 32 + lane < N;
      const float* src = pc_b + (size_t)row_next * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) dpc_cp_async4(&pts[buf][lane * 3 + c], lv ? src + c : pc_b, lv);
      return;
    }
    const int n3 = min(32, N - tt * 32) * 3;
    const float* src = pc_b + (size_t)tt * 96;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int i = lane + 32 * c;
      dpc_cp_async4(&pts[buf][i], i < n3 ? src + i : pc_b, i < n3);
    }
  };
  dpc_kt_mark_if(kt, DPC_KT_SPLAT_F, 0);
  dpc_grid_dep_trigger();
  bool waited = ZERO;
  if (SEL && !waited) { dpc_grid_dep_wait(); dpc_kt_mark_if(kt, DPC_KT_SPLAT_F, 1); waited = true; }
  // the points and the camera are inputs: nothing in front of this kernel writes them (see dpc_splat_fwd_kernel)
  if (t < tiles) {
    load_row(t);
    issue_points(t, 0);
    if (t + wps < tiles) load_row(t + wps);
  }
  dpc_cp_async_commit();
  if (threadIdx.x == 0) dpc_pose_load(pose_sm, a.pose, a.pose_kind, a.trans, a.focal, a.focal_const, a.cam_dist, b);
  if (ZERO) {
    // launched without a programmatic dependency (everything older has completed): the zeros may go out at once
    float4* z4 = reinterpret_cast<float4*>(a.vox);
    const size_t n4 = (size_t)a.B * Vz * V * (V >> 2);
    const size_t nth = (size_t)gridDim.x * gridDim.y * blockDim.x;
    const float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += nth) z4[i] = zz;
  }
  __syncthreads();               // the camera of the sample
  if (!waited && !a.early) { dpc_grid_dep_wait(); dpc_kt_mark_if(kt, DPC_KT_SPLAT_F, 1); waited = true; }
  float* grid = a.vox ? a.vox + (size_t)b * Vz * V * V : nullptr;
  const bool quad_al = a.red4 && ((V & 3) == 0) && ((((uintptr_t)a.vox) & 15u) == 0);
  const bool pair_al = ((V & 1) == 0) && ((((uintptr_t)a.vox) & 7u) == 0);
  // front half of a tile: its points have been requested; transform them, write tr_pc, request the next tile's points
  float z = 0.f, y = 0.f, x = 0.f;
  bool live = false;
  auto front = [&](int tt, int buf) {
    dpc_cp_async_wait<0>();
    __syncwarp();
    if (tt + wps < tiles) {
      issue_points(tt + wps, buf ^ 1);             // (uses row_next = the list entries of that tile)
      if (tt + 2 * wps < tiles) load_row(tt + 2 * wps);
    }
    dpc_cp_async_commit();
    const int n = min(32, N - tt * 32);
    live = lane < n;
    z = y = x = 0.f;
    if (live) {
      DpcCamPoint cam;
      dpc_transform_point(pose_sm, pts[buf][lane * 3 + 0], pts[buf][lane * 3 + 1], pts[buf][lane * 3 + 2], z, y, x, cam);
    }
    if (a.tr_pc) {       // (early mode: written ahead of the grid dependency, as in dpc_splat_fwd_kernel)
      if (live) { pts[buf][lane * 3 + 0] = z; pts[buf][lane * 3 + 1] = y; pts[buf][lane * 3 + 2] = x; }
      __syncwarp();
      float* dst = a.tr_pc + ((size_t)b * N + (size_t)tt * 32) * 3;
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int q = lane + 32 * cc;
        if (q < n * 3) dst[q] = pts[buf][q];
      }
    }
  };
  if (t < tiles) front(t, 0);
#ifndef DPC_EMU
  if (ZERO) {
    __threadfence();                         // my zeros are visible device-wide ...
    cooperative_groups::this_grid().sync();  // ... and so are everybody else's: the grid is zero from here on
    dpc_kt_mark_if(kt, DPC_KT_SPLAT_F, 1);
  }
#endif
  if (t >= tiles) return;
  if (!waited) { dpc_grid_dep_wait(); dpc_kt_mark_if(kt, DPC_KT_SPLAT_F, 1); waited = true; }
  int buf = 0;
  while (true) {
    DpcCell c = dpc_cell(z, y, x, Vz, V);
    c.valid = c.valid && live;
    if (live) {
      const size_t pi = (size_t)b * N + (size_t)t * 32 + lane;
      if (a.idx_out) { a.idx_out[pi * 3 + 0] = c.iz; a.idx_out[pi * 3 + 1] = c.iy; a.idx_out[pi * 3 + 2] = c.ix; }
      if (a.valid_out) a.valid_out[pi] = c.valid ? 1 : 0;
    }
    if (grid) {
      // corner weights, reference association: (rr[k].z * rr[j].y) * rr[i].x  (point_cloud.py:99)
      const float wz[2] = {__fsub_rn(1.0f, c.rz), c.rz};
      const float wy[2] = {__fsub_rn(1.0f, c.ry), c.ry};
      const float wx[2] = {__fsub_rn(1.0f, c.rx), c.rx};
      float w[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) w[q] = c.valid ? __fmul_rn(__fmul_rn(wz[q >> 2], wy[(q >> 1) & 1]), wx[q & 1]) : 0.0f;
      const int base = (c.iz * V + c.iy) * V + c.ix;
      // warp aggregation of lanes that share a base voxel (see dpc_splat_fwd_kernel)
      const int key = c.valid ? base : (-1 - lane);
      const unsigned peers = __match_any_sync(DPC_FULL, key);
      const int cnt = __popc(peers);
      const int maxcnt = __reduce_max_sync(DPC_FULL, cnt);
      if (maxcnt > 1) {
        float s8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) s8[q] = w[q];
        unsigned rem = peers & ~(1u << lane);
        for (int it = 1; it < maxcnt; ++it) {
          const int src = rem ? (__ffs(rem) - 1) : lane;
          rem &= rem - 1;
          const bool take = it < cnt;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float v = __shfl_sync(DPC_FULL, w[q], src);
            if (take) s8[q] += v;
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) w[q] = s8[q];
      }
      if (c.valid && (lane == (__ffs(peers) - 1))) {
        float* g = grid + base;
        const int o4 = c.ix & 3;
        const bool quad_ok = quad_al && (o4 != 3);
        const bool pair_ok = pair_al && ((c.ix & 1) == 0);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (c.iz + k >= Vz) continue;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            if (c.iy + jj >= V) continue;
            float* row = g + (k * V + jj) * V;
            const float w0 = w[k * 4 + jj * 2 + 0], w1 = w[k * 4 + jj * 2 + 1];
            if (quad_ok) {
              dpc_red_add4(row - o4, o4 == 0 ? w0 : 0.0f, o4 == 0 ? w1 : (o4 == 1 ? w0 : 0.0f),
                           o4 == 1 ? w1 : (o4 == 2 ? w0 : 0.0f), o4 == 2 ? w1 : 0.0f);
            } else if (pair_ok) {
              dpc_red_add2(row, w0, w1);
            } else {
              dpc_red_add(row, w0);
              if (c.ix + 1 < V) dpc_red_add(row + 1, w1);
            }
          }
        }
      }
    }
    __syncwarp();
    t += wps;
    buf ^= 1;
    if (t >= tiles) break;
    front(t, buf);
  }
  dpc_cp_async_wait<0>();
  dpc_kt_mark_if(kt, DPC_KT_SPLAT_F, 3);
}

// ------------------------------------------------------------------------------ grid zeroing (lab build: knob 10)
#ifdef DPC_EXPERIMENTS
// Zeroes the raw grid ahead of the splat.  Besides initialising the accumulation target this
// leaves the grid resident (dirty) in L2, so the splat's reductions hit L2 instead of fetching
// 32-byte sectors from HBM at random (measured: the splat kernel takes ~10 us longer on a cold grid).
#ifndef DPC_EMU
__global__ void __launch_bounds__(256)
#else
static void
#endif
dpc_zero_kernel(float4* dst, size_t n4, float* tail, int ntail) {
  // wait FIRST, trigger second: the splat that follows reads its inputs before its own wait, so by the time it may
  // start everything older than this kernel has to be complete (see dpc_grid_dep_trigger)
  dpc_kt_mark(DPC_KT_ZERO, 0);
  dpc_grid_dep_wait();
  dpc_grid_dep_trigger();
  dpc_kt_mark(DPC_KT_ZERO, 1);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
  for (; i + 3 * stride < n4; i += 4 * stride) { dst[i] = zz; dst[i + stride] = zz; dst[i + 2 * stride] = zz; dst[i + 3 * stride] = zz; }
  for (; i < n4; i += stride) dst[i] = zz;
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0.0f;
  dpc_kt_mark(DPC_KT_ZERO, 3);
}

// Same role, zeros written by the TMA engine (bulk stores from a zeroed 16 KiB of shared memory) instead of by the
// SMs' store path: experiment knob 10 = 2 (does the placement of the zero lines in L2 matter to the reductions?).
#ifndef DPC_EMU
__global__ void __launch_bounds__(128)
dpc_zero_bulk_kernel(unsigned char* dst, size_t bytes) {
  __shared__ __align__(128) unsigned char z[16384];
  dpc_kt_mark(DPC_KT_ZERO, 0);
  dpc_grid_dep_wait();
  dpc_grid_dep_trigger();
  dpc_kt_mark(DPC_KT_ZERO, 1);
  for (int i = threadIdx.x; i < 16384 / 16; i += blockDim.x) reinterpret_cast<float4*>(z)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  dpc_fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(z);
    for (size_t off = (size_t)blockIdx.x * 16384; off < bytes; off += (size_t)gridDim.x * 16384)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(sa), "r"(16384u) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  dpc_kt_mark(DPC_KT_ZERO, 3);
}
#endif

#endif  // DPC_EXPERIMENTS

// Zeroes up to four small accumulation targets (pose / translation / focal / scale gradients) in ONE launch: every
// stream operation costs ~2.5 us of dependency latency on B200, four cudaMemsetAsync of a few bytes each cost 10 us.
struct DpcZero4Args { float* p[4]; int n[4]; };
#ifndef DPC_EMU
__global__ void __launch_bounds__(256)
#else
static void
#endif
dpc_zero4_kernel(DpcZero4Args a) {
  dpc_kt_mark(DPC_KT_ZERO4, 0);
  dpc_grid_dep_sync();
  dpc_kt_mark(DPC_KT_ZERO4, 1);
  for (int k = 0; k < 4; ++k)
    if (a.p[k]) for (int i = threadIdx.x; i < a.n[k]; i += blockDim.x) a.p[k][i] = 0.0f;
}

// ------------------------------------------------------------------------------ f-2: dropout gather
#ifndef DPC_EMU
__global__ void
#else
static void
#endif
dpc_gather_kernel(const float* in, const int64_t* sel, int N, int n_keep, int C, float* out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  dpc_grid_dep_wait();   // no early trigger: the output is point data, which a splat reads before its own wait
  if (i >= n_keep * C) return;
  const int r = i / C, ch = i - r * C;
  const int64_t s = sel[(size_t)b * n_keep + r];
  out[((size_t)b * n_keep + r) * C + ch] = in[((size_t)b * N + s) * C + ch];
}

#ifndef DPC_EMU
__global__ void
#else
static void
#endif
dpc_gather_bwd_kernel(const float* g_out, const int64_t* sel, int N, int n_keep, int C, float* g_in) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  dpc_grid_dep_wait();
  if (i >= n_keep * C) return;
  const int r = i / C, ch = i - r * C;
  const int64_t s = sel[(size_t)b * n_keep + r];
  atomicAdd(g_in + ((size_t)b * N + s) * C + ch, g_out[((size_t)b * n_keep + r) * C + ch]);
}

// ------------------------------------------------------------------------------ f-2: dropout index lists on the device
// pc_point_dropout (point_cloud.py:293-319) keeps int(N * keep_prob) points per sample, drawn without replacement with
// np.random.choice on the host through tf.py_func.  Here the subset of sample b is {pi_b(0), ..., pi_b(n_keep - 1)} for a
// pseudo-random PERMUTATION pi_b of [0, N): a 6-round Feistel network on ceil(log2 N) bits (cycle-walking back into
// [0, N)), keyed per sample and per draw by Philox4x32-10(counter = (b, draw), key = seed).  O(1) per index, no sort, no
// B x N noise tensor; distinct by construction.  state (device, optional): {seed, draw} read on the device, so a
// captured CUDA graph draws a fresh subset on every replay (the caller bumps `draw` with a one-element add).
DPC_DEV void dpc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
DPC_DEV uint32_t dpc_mix32(uint32_t x) {      // finaliser of murmur3: the Feistel round function
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
#ifndef DPC_EMU
__global__ void
#else
static void
#endif
dpc_dropout_indices_kernel(unsigned long long seed, unsigned long long draw, const unsigned long long* state,
                           int N, int n_keep, int32_t* sel) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  dpc_grid_dep_wait();
  if (i >= n_keep) return;
  if (state) { seed = state[0]; draw = state[1]; }
  uint32_t key[8];
  dpc_philox4x32_10((uint32_t)b, (uint32_t)draw, (uint32_t)(draw >> 32), 0u, (uint32_t)seed, (uint32_t)(seed >> 32), key);
  dpc_philox4x32_10((uint32_t)b, (uint32_t)draw, (uint32_t)(draw >> 32), 1u, (uint32_t)seed, (uint32_t)(seed >> 32), key + 4);
  int bits = 1;
  while ((1u << bits) < (unsigned)N) ++bits;
  if (bits < 2) bits = 2;
  const int lb = bits / 2, hb = bits - lb;               // low half lb bits, high half hb bits (unbalanced when bits is odd)
  const uint32_t lmask = (1u << lb) - 1u, hmask = (1u << hb) - 1u;
  uint32_t x = (uint32_t)i;
  do {
    uint32_t lo = x & lmask, hi = x >> lb;
#pragma unroll
    for (int r = 0; r < 6; r += 2) {
      hi = (hi ^ dpc_mix32(lo ^ key[r])) & hmask;
      lo = (lo ^ dpc_mix32(hi ^ key[r + 1])) & lmask;
    }
    x = (hi << lb) | lo;
  } while (x >= (uint32_t)N);
  sel[(size_t)b * n_keep + i] = (int32_t)x;
}
