// f-4: nearest-neighbour projection between two point sets (the chamfer evaluation's inner op).
//
// Replaces point_cloud_distance (util/point_cloud_distance.py:26-39): the reference tiles both sets to
// [VsN, VtN, 3] (24 * VsN * VtN bytes in fp64 -- the evaluation has to cut the source into 10 parts,
// run/eval_chamfer.py:18-34, default_config.yaml:155, to fit), takes sqrt(sum(diff^2)) of every pair, argmin over
// the targets and two gathers.  Here: thread = source point, target tiles staged in shared memory, a running
// (distance, first index) pair in registers, the targets split over blockIdx.y so that a few thousand sources still
// fill 148 SMs; a second small kernel folds the splits in target order.  Nothing of size VsN * VtN is materialised.
//
// Bit-exactness of `idx`: tf.argmin runs on the ROUNDED sqrt values and returns the first minimum.  sqrt is
// monotone, so the scan compares squared distances and takes a square root only when a strictly smaller squared
// distance shows up (O(log VtN) times per source): the index moves only if the rounded root is strictly smaller
// too, exactly what a scan over the rounded roots would do.  diff = Vt - Vs and (d0^2 + d1^2) + d2^2 follow the
// reference's op order (`:33-34`), each operation rounded once (no FMA contraction).
#pragma once
#include "dpc_common.cuh"

#define DPC_NN_THREADS 256
#define DPC_NN_TILE 1024      // targets staged per tile: 12 KiB (fp32) / 24 KiB (fp64) of shared memory

#ifndef DPC_EMU
DPC_DEV float dpc_nn_sub(float a, float b) { return __fsub_rn(a, b); }
DPC_DEV float dpc_nn_mul(float a, float b) { return __fmul_rn(a, b); }
DPC_DEV float dpc_nn_add(float a, float b) { return __fadd_rn(a, b); }
DPC_DEV float dpc_nn_sqrt(float a) { return __fsqrt_rn(a); }
DPC_DEV double dpc_nn_sub(double a, double b) { return __dsub_rn(a, b); }
DPC_DEV double dpc_nn_mul(double a, double b) { return __dmul_rn(a, b); }
DPC_DEV double dpc_nn_add(double a, double b) { return __dadd_rn(a, b); }
DPC_DEV double dpc_nn_sqrt(double a) { return __dsqrt_rn(a); }
#else
DPC_DEV float dpc_nn_sub(float a, float b) { return a - b; }
DPC_DEV float dpc_nn_mul(float a, float b) { return a * b; }
DPC_DEV float dpc_nn_add(float a, float b) { return a + b; }
DPC_DEV float dpc_nn_sqrt(float a) { return sqrtf(a); }
DPC_DEV double dpc_nn_sub(double a, double b) { return a - b; }
DPC_DEV double dpc_nn_mul(double a, double b) { return a * b; }
DPC_DEV double dpc_nn_add(double a, double b) { return a + b; }
DPC_DEV double dpc_nn_sqrt(double a) { return sqrt(a); }
#endif

template <typename T> DPC_DEV T dpc_nn_inf() { return (T)INFINITY; }

// grid (ceil(ns / 256), splits): block (bx, by) scans targets [by * chunk, min(nt, (by + 1) * chunk)) for its
// 256 sources and writes the split's (rounded distance, first index) to part_s / part_i [splits, ns].
template <typename T>
#ifndef DPC_EMU
__global__ void __launch_bounds__(DPC_NN_THREADS)
#else
static void
#endif
dpc_nn_partial_kernel(const T* vs, int ns, const T* vt, int nt, int chunk, T* part_s, int32_t* part_i) {
  __shared__ T tile[DPC_NN_TILE * 3];
  const int i = blockIdx.x * DPC_NN_THREADS + threadIdx.x;
  const int t_first = blockIdx.y * chunk;
  const int t_last = min(nt, t_first + chunk);
  dpc_grid_dep_sync();
  T sx = 0, sy = 0, sz = 0;
  if (i < ns) { sx = vs[(size_t)i * 3 + 0]; sy = vs[(size_t)i * 3 + 1]; sz = vs[(size_t)i * 3 + 2]; }
  T best_d2 = dpc_nn_inf<T>(), best_s = dpc_nn_inf<T>();
  int best_i = t_first < nt ? t_first : 0;
  for (int t0 = t_first; t0 < t_last; t0 += DPC_NN_TILE) {
    const int cnt = min(DPC_NN_TILE, t_last - t0);
    __syncthreads();      // the previous tile has been scanned by every thread
    for (int q = threadIdx.x; q < cnt * 3; q += DPC_NN_THREADS) tile[q] = vt[(size_t)t0 * 3 + q];
    __syncthreads();
    if (i < ns) {
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const T d0 = dpc_nn_sub(tile[3 * j + 0], sx), d1 = dpc_nn_sub(tile[3 * j + 1], sy), d2 = dpc_nn_sub(tile[3 * j + 2], sz);
        const T q2 = dpc_nn_add(dpc_nn_add(dpc_nn_mul(d0, d0), dpc_nn_mul(d1, d1)), dpc_nn_mul(d2, d2));
        if (q2 < best_d2) {       // rare: the running minimum of a random sequence improves O(log n) times
          best_d2 = q2;
          const T s = dpc_nn_sqrt(q2);
          if (s < best_s) { best_s = s; best_i = t0 + j; }
        }
      }
    }
  }
  if (i < ns) {
    part_s[(size_t)blockIdx.y * ns + i] = best_s;
    part_i[(size_t)blockIdx.y * ns + i] = best_i;
  }
}

// thread = source point: fold the splits in target order (strict <, so the first minimum wins), gather the projection
template <typename T>
#ifndef DPC_EMU
__global__ void __launch_bounds__(DPC_NN_THREADS)
#else
static void
#endif
dpc_nn_final_kernel(const T* vt, int ns, int splits, const T* part_s, const int32_t* part_i,
                    T* proj, T* min_dist, int32_t* idx) {
  const int i = blockIdx.x * DPC_NN_THREADS + threadIdx.x;
  dpc_grid_dep_sync();
  if (i >= ns) return;
  T best_s = part_s[i];
  int best_i = part_i[i];
  for (int k = 1; k < splits; ++k) {
    const T s = part_s[(size_t)k * ns + i];
    if (s < best_s) { best_s = s; best_i = part_i[(size_t)k * ns + i]; }
  }
  if (min_dist) min_dist[i] = best_s;
  if (idx) idx[i] = best_i;
  if (proj) {
    proj[(size_t)i * 3 + 0] = vt[(size_t)best_i * 3 + 0];
    proj[(size_t)i * 3 + 1] = vt[(size_t)best_i * 3 + 1];
    proj[(size_t)i * 3 + 2] = vt[(size_t)best_i * 3 + 2];
  }
}

// Splits of the target set: enough blocks for ~2 waves of 148 SMs x 4 resident CTAs, at least one tile per split.
static inline int dpc_nn_splits(int ns, int nt) {
  const int bx = (ns + DPC_NN_THREADS - 1) / DPC_NN_THREADS;
  int want = (148 * 8 + bx - 1) / bx;
  const int max_by_tiles = (nt + DPC_NN_TILE - 1) / DPC_NN_TILE;
  if (want > max_by_tiles) want = max_by_tiles;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  return want;
}
