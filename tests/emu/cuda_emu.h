// TEST INFRASTRUCTURE ONLY: a small CPU emulation of the CUDA execution model, just enough to
// run the kernels of differentiable-point-clouds_b200/csrc unchanged on the host (compiled with
// `g++ -x c++ -DDPC_EMU -ffp-contract=off`).  Every CUDA thread of a block is an OS thread;
// blocks run one after another; __syncthreads and the warp collectives are barriers.  It lets
// the CPU test-suite exercise the real kernel sources (indexing, tiling, halos, gradient
// formulas) without a GPU.  It is never built into, loaded by, or selected by the product.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <barrier>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
#define cudaSuccess 0
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }

namespace dpc_emu {
struct WarpCtx {
  std::barrier<> bar{32};
  uint32_t slot[32];
};
struct BlockCtx {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<std::unique_ptr<WarpCtx>> warps;
  std::vector<unsigned char> dyn;
};
extern thread_local uint3_emu t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
extern thread_local int t_lane, t_warp;
extern BlockCtx* g_block;
unsigned char* dyn_smem();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
uint32_t exchange(uint32_t v, int src);          // value of lane `src`
void gather(uint32_t v, uint32_t* all32);         // all lanes' values
// mbarrier model: the word counts completed phases; wait(parity P) passes once the phase of
// parity P has completed, i.e. while (count & 1) == P it blocks.
static inline void mbar_complete(uint64_t* bar) { std::atomic_ref<uint64_t>(*bar).fetch_add(1, std::memory_order_release); }
static inline void mbar_wait(uint64_t* bar, unsigned parity) {
  while ((std::atomic_ref<uint64_t>(*bar).load(std::memory_order_acquire) & 1u) == parity) std::this_thread::yield();
}
}  // namespace dpc_emu

#define threadIdx dpc_emu::t_threadIdx
#define blockIdx dpc_emu::t_blockIdx
#define blockDim dpc_emu::t_blockDim
#define gridDim dpc_emu::t_gridDim

#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__
#define INFINITY_F INFINITY

static inline void __syncthreads() { dpc_emu::g_block->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { dpc_emu::g_block->warps[dpc_emu::t_warp]->bar.arrive_and_wait(); }

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline float __shfl_sync(unsigned, float v, int src) { return u2f(dpc_emu::exchange(f2u(v), src & 31)); }
static inline int __shfl_sync(unsigned, int v, int src) { return (int)dpc_emu::exchange((uint32_t)v, src & 31); }
static inline float __shfl_xor_sync(unsigned, float v, int m) { return u2f(dpc_emu::exchange(f2u(v), (dpc_emu::t_lane ^ m) & 31)); }
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int m) { return dpc_emu::exchange(v, (dpc_emu::t_lane ^ m) & 31); }
static inline float __shfl_down_sync(unsigned, float v, int d) {
  int s = dpc_emu::t_lane + d; if (s > 31) s = dpc_emu::t_lane;
  return u2f(dpc_emu::exchange(f2u(v), s));
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  uint32_t all[32]; dpc_emu::gather(pred ? 1u : 0u, all);
  unsigned r = 0; for (int i = 0; i < 32; ++i) r |= (all[i] & 1u) << i; return r;
}
static inline unsigned __match_any_sync(unsigned, int key) {
  uint32_t all[32]; dpc_emu::gather((uint32_t)key, all);
  unsigned r = 0; for (int i = 0; i < 32; ++i) if (all[i] == (uint32_t)key) r |= 1u << i; return r;
}
static inline int __reduce_max_sync(unsigned, int v) {
  uint32_t all[32]; dpc_emu::gather((uint32_t)v, all);
  int r = (int)all[0]; for (int i = 1; i < 32; ++i) r = std::max(r, (int)all[i]); return r;
}
static inline int __all_sync(unsigned, int pred) { return __ballot_sync(0xffffffffu, pred) == 0xffffffffu; }
static inline int __any_sync(unsigned, int pred) { return __ballot_sync(0xffffffffu, pred) != 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }

static inline float atomicAdd(float* addr, float v) {
  std::atomic_ref<float> a(*addr);
  float old = a.load(std::memory_order_relaxed);
  while (!a.compare_exchange_weak(old, old + v, std::memory_order_relaxed)) {}
  return old;
}
static inline float __ldg(const float* p) { return *p; }
static inline float4 __ldg(const float4* p) { return *p; }
static inline int __ldg(const int* p) { return *p; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fdividef(float a, float b) { return a / b; }
using std::max;
using std::min;
