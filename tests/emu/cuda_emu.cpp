// TEST INFRASTRUCTURE ONLY -- see cuda_emu.h.  Builds libdpc_b200_emu.so from the product's
// kernel sources for the CPU test-suite.
#include "cuda_emu.h"

namespace dpc_emu {
thread_local uint3_emu t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;
thread_local int t_lane, t_warp;
BlockCtx* g_block = nullptr;

unsigned char* dyn_smem() { return g_block->dyn.data(); }

uint32_t exchange(uint32_t v, int src) {
  WarpCtx& w = *g_block->warps[t_warp];
  w.slot[t_lane] = v;
  w.bar.arrive_and_wait();
  uint32_t r = w.slot[src];
  w.bar.arrive_and_wait();
  return r;
}
void gather(uint32_t v, uint32_t* all32) {
  WarpCtx& w = *g_block->warps[t_warp];
  w.slot[t_lane] = v;
  w.bar.arrive_and_wait();
  for (int i = 0; i < 32; ++i) all32[i] = w.slot[i];
  w.bar.arrive_and_wait();
}

// A kernel that `return`s early from some threads would deadlock a std::barrier sized for the
// whole block, so a thread that leaves the body drops out of the block barrier (arrive_and_drop)
// -- matching CUDA, where exited threads no longer count for bar.sync.
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  if (nthreads % 32 != 0) abort();
  const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  for (size_t bi = 0; bi < nblocks; ++bi) {
    BlockCtx ctx;
    ctx.bar = std::make_unique<std::barrier<>>(nthreads);
    for (unsigned w = 0; w < nthreads / 32; ++w) ctx.warps.emplace_back(new WarpCtx());
    ctx.dyn.assign(smem + 64, 0);
    g_block = &ctx;
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) {
      th.emplace_back([&, t]() {
        t_threadIdx.x = t % block.x;
        t_threadIdx.y = (t / block.x) % block.y;
        t_threadIdx.z = t / (block.x * block.y);
        t_blockIdx.x = (unsigned)(bi % grid.x);
        t_blockIdx.y = (unsigned)((bi / grid.x) % grid.y);
        t_blockIdx.z = (unsigned)(bi / ((size_t)grid.x * grid.y));
        t_blockDim = block;
        t_gridDim = grid;
        t_lane = t & 31;
        t_warp = t >> 5;
        body();
        ctx.bar->arrive_and_drop();
      });
    }
    for (auto& x : th) x.join();
    g_block = nullptr;
  }
}
}  // namespace dpc_emu
