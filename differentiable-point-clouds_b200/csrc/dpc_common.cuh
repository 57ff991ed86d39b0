// Common definitions for the sm_100a kernels.  The same sources compile in two ways:
//   * nvcc -gencode arch=compute_100a,code=sm_100a   -> the product library (libdpc_b200.so)
//   * g++ -x c++ -DDPC_EMU                            -> tests/emu: every CUDA thread is an OS
//     thread, used ONLY by the CPU test-suite to check indexing / gradient logic of these very
//     kernels before GPU time is spent.  The emulation build is never loaded by the product.
#pragma once

#include <stdint.h>

#ifdef DPC_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

#include "../../include/dpc_b200.h"

// Experiment knobs (dpc_debug_set) exist only in the LAB build (-DDPC_EXPERIMENTS: libdpc_b200_lab.so, and the CPU
// emulation build of the test-suite).  In the product build every knob is a compile-time constant (its default), the
// experimental kernels are not compiled, and the library has no process-wide mutable state besides the two diagnostics
// switches (stage events, kernel timeline) and the kernel-family selector.
#ifdef DPC_EXPERIMENTS
#define DPC_KNOB_T int
#else
#define DPC_KNOB_T const int
#endif

#define DPC_WARP 32
#define DPC_FULL 0xffffffffu

#ifndef DPC_EMU
#define DPC_GRID_CONSTANT __grid_constant__
#define DPC_DEV __device__ __forceinline__
// Every kernel is launched with Programmatic Dependent Launch allowed: its CTAs may be scheduled
// while the previous kernel of the stream is still draining, run their prologue (taps into smem,
// mbarrier init), and block in dpc_grid_dep_sync() until the previous grid has completed and its
// memory is visible.  That hides launch latency and the tail of each kernel behind the next one.
template <typename... KArgs, typename... Args>
static inline void dpc_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define DPC_LAUNCH(kernel, grid, block, smem, stream, ...) \
  dpc_launch_pdl(kernel, (grid), (block), (smem), (stream), __VA_ARGS__)
#define DPC_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; \
  type* name = reinterpret_cast<type*>(name##_raw)
#else
#define DPC_GRID_CONSTANT
#define DPC_DEV static inline
#define DPC_LAUNCH(kernel, grid, block, smem, stream, ...) \
  dpc_emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define DPC_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(dpc_emu::dyn_smem())
#endif

// Block until the kernel(s) this launch depends on have completed (no-op without PDL).  Must
// precede the first access to anything a previous kernel wrote, and the first global write.
// It first signals that this grid's own dependents may be scheduled: that takes effect once EVERY
// CTA of this grid has started (i.e. during its last wave), so the next kernel's CTAs fill the SM
// slots freed by this kernel's tail and wait here for its completion.
DPC_DEV void dpc_grid_dep_sync() {
#ifndef DPC_EMU
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// The two halves separately.  A kernel whose first phase reads nothing its stream predecessor wrote may
// run that phase before the wait (the splat stages and transforms its points while the raw grid is being
// zeroed).  The rule that keeps this safe transitively: a kernel whose OUTPUT a dependent may read early
// (point data: the dropout gather) or that sits directly in front of such a dependent (the grid-zeroing
// kernel) triggers only AFTER its own wait, so everything older than it has completed by then.
DPC_DEV void dpc_grid_dep_trigger() {
#ifndef DPC_EMU
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
DPC_DEV void dpc_grid_dep_wait() {
#ifndef DPC_EMU
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// ---- diagnostics: per-kernel timeline of one step (dpc_debug_set(12, 1); scripts/step_timeline.py).
// Thread 0 of every CTA folds %globaltimer into [kernel][0] = first entry, [1] = first CTA past its grid
// dependency, [2] = last CTA past it, [3] = last exit.
#ifndef DPC_EMU
__device__ unsigned long long dpc_kt[16 * 4];
__device__ int dpc_kt_on = 0;
DPC_DEV void dpc_kt_mark(int id, int slot) {
  if (threadIdx.x == 0 && dpc_kt_on) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (slot < 2) atomicMin(&dpc_kt[id * 4 + slot], t); else atomicMax(&dpc_kt[id * 4 + slot], t);
    if (slot == 1) atomicMax(&dpc_kt[id * 4 + 2], t);
  }
}
// the same with the switch read ONCE by the caller (a kernel that stamps behind its grid dependency would otherwise put
// a global load of the flag on its critical path)
DPC_DEV bool dpc_kt_enabled() { return dpc_kt_on != 0; }
DPC_DEV void dpc_kt_mark_if(bool on, int id, int slot) {
  if (on && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (slot < 2) atomicMin(&dpc_kt[id * 4 + slot], t); else atomicMax(&dpc_kt[id * 4 + slot], t);
    if (slot == 1) atomicMax(&dpc_kt[id * 4 + 2], t);
  }
}
// per-CTA phase stamps of the two splat kernels (same switch): [which][cta < 512][8 slots]
__device__ unsigned long long dpc_ph[2 * 512 * 8];
DPC_DEV void dpc_ph_mark(int which, int slot) {
  if (threadIdx.x == 0 && dpc_kt_on) {
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t) :: "memory");
    if (cta < 512u) dpc_ph[((unsigned)which * 512u + cta) * 8u + (unsigned)slot] = t;
  }
}
#else
DPC_DEV void dpc_kt_mark(int, int) {}
DPC_DEV bool dpc_kt_enabled() { return false; }
DPC_DEV void dpc_kt_mark_if(bool, int, int) {}
DPC_DEV void dpc_ph_mark(int, int) {}
#endif
enum { DPC_KT_ZERO = 0, DPC_KT_SPLAT_F, DPC_KT_XY_F, DPC_KT_Z_F, DPC_KT_ZERO4, DPC_KT_Z_B, DPC_KT_XY_B, DPC_KT_SPLAT_B };

// ------------------------------------------------------------------ small PTX wrappers
// red.global.add.f32: fire-and-forget fp32 add at L2 (SASS REDG.E.ADD.F32).
DPC_DEV void dpc_red_add(float* addr, float v) {
#ifndef DPC_EMU
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
#else
  atomicAdd(addr, v);
#endif
}

// red.global.add.v2.f32 (sm_90+): two adjacent floats in one L2 reduction (REDG.E.ADD.F32x2).
// addr must be 8-byte aligned.
DPC_DEV void dpc_red_add2(float* addr, float a, float b) {
#ifndef DPC_EMU
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
#else
  atomicAdd(addr, a);
  atomicAdd(addr + 1, b);
#endif
}

// red.global.add.v4.f32 (sm_90+): four adjacent floats in one L2 reduction.  addr must be 16-byte aligned.
DPC_DEV void dpc_red_add4(float* addr, float a, float b, float c, float d) {
#ifndef DPC_EMU
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
#else
  atomicAdd(addr, a); atomicAdd(addr + 1, b); atomicAdd(addr + 2, c); atomicAdd(addr + 3, d);
#endif
}

// Packed fp32x2 FMA (Blackwell FFMA2): d = a*b+c on both halves.
DPC_DEV float2 dpc_ffma2(float2 a, float2 b, float2 c) {
#ifndef DPC_EMU
  return __ffma2_rn(a, b, c);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

// ---- 1-D bulk async copy global -> shared through the TMA engine (SASS UBLKCP), completion
// on an mbarrier.  Requirements: 16-byte aligned src/dst, bytes % 16 == 0.
DPC_DEV void dpc_mbar_init(uint64_t* bar, unsigned count) {
#ifndef DPC_EMU
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
  *bar = 0;
  (void)count;
#endif
}

DPC_DEV void dpc_bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
#ifndef DPC_EMU
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
#else
  memcpy(smem_dst, gmem_src, bytes);
  dpc_emu::mbar_complete(bar);
#endif
}

DPC_DEV void dpc_mbar_wait(uint64_t* bar, unsigned phase) {
#ifndef DPC_EMU
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a), "r"(phase) : "memory");
#else
  dpc_emu::mbar_wait(bar, phase);
#endif
}

// Make this thread's generic-proxy writes to shared memory visible to the async proxy (TMA);
// every writer calls it before the barrier that precedes a bulk store.
DPC_DEV void dpc_fence_proxy_async() {
#ifndef DPC_EMU
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// bulk async copy shared -> global (TMA store), bulk-group completion.
DPC_DEV void dpc_bulk_store(void* gmem_dst, const void* smem_src, unsigned bytes) {
#ifndef DPC_EMU
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_src);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(s), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#else
  memcpy(gmem_dst, smem_src, bytes);
#endif
}

DPC_DEV float dpc_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DPC_FULL, v, o);
  return v;
}

DPC_DEV float dpc_clip01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
