// Silhouette loss of the training step and its gradient in one pass.
//
// Replaces `tf.nn.l2_loss(gt - pred) / num_samples` (models/model_pc.py:414-415; l2_loss = sum(x^2) / 2) and the
// element-wise kernels TF's autodiff derives from it: loss = sum((gt - pred)^2) / 2 * inv_count and
// dL/dpred = (pred - gt) * inv_count, written while the difference is in registers.  It sits between the forward and
// the backward of the projection path; as a PDL-aware kernel of this library it lets the first backward kernel run
// its (non-overlappable, ~2.5 us) start-up underneath it, which a framework's element-wise kernel does not.
// Deterministic: per-CTA partial sums, folded in index order by whichever CTA finishes last (self-resetting counter).
#pragma once
#include "dpc_common.cuh"

#define DPC_LOSS_THREADS 256
#define DPC_LOSS_MAX_CTAS 512

#ifndef DPC_EMU
__global__ void __launch_bounds__(DPC_LOSS_THREADS)
#else
static void
#endif
dpc_proj_l2_loss_kernel(const float* pred, const float* gt, long long n, float inv_count,
                        float* loss, float* g_pred, float* partial, unsigned* counter) {
  __shared__ float red[DPC_LOSS_THREADS / 32];
  __shared__ int last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  dpc_grid_dep_sync();
  float acc = 0.0f;
  const long long stride = (long long)gridDim.x * DPC_LOSS_THREADS;
  const bool vec = ((n & 3) == 0) && ((((uintptr_t)pred | (uintptr_t)gt | (uintptr_t)g_pred) & 15u) == 0);
  if (vec) {
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * DPC_LOSS_THREADS + tid; i < n4; i += stride) {
      const float4 p = reinterpret_cast<const float4*>(pred)[i], t = reinterpret_cast<const float4*>(gt)[i];
      const float d0 = t.x - p.x, d1 = t.y - p.y, d2 = t.z - p.z, d3 = t.w - p.w;
      acc += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      if (g_pred) reinterpret_cast<float4*>(g_pred)[i] = make_float4(-d0 * inv_count, -d1 * inv_count, -d2 * inv_count, -d3 * inv_count);
    }
  } else {
    for (long long i = (long long)blockIdx.x * DPC_LOSS_THREADS + tid; i < n; i += stride) {
      const float d = gt[i] - pred[i];
      acc += d * d;
      if (g_pred) g_pred[i] = -d * inv_count;
    }
  }
  if (!loss) return;        // gradient only (uniform): no reduction, no fence, no counter
  acc = dpc_warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    float v = 0.0f;
    for (int w = 0; w < DPC_LOSS_THREADS / 32; ++w) v += red[w];
    partial[blockIdx.x] = v;
#ifndef DPC_EMU
    __threadfence();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
#else
    last = (blockIdx.x == gridDim.x - 1) ? 1 : 0;      // the emulation runs the CTAs one after the other
#endif
  }
  __syncthreads();
  if (last && warp == 0) {
#ifndef DPC_EMU
    __threadfence();
#endif
    float v = 0.0f;
    for (unsigned i = lane; i < gridDim.x; i += 32) v += ((volatile float*)partial)[i];
    v = dpc_warp_sum(v);
    if (lane == 0) { *loss = 0.5f * v * inv_count; *counter = 0u; }
  }
}
