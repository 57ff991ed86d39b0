// Backward x/y pass of the smoothing AND the gathers of the splat backward in ONE persistent kernel (64^3 grids,
// training case).
//
// Why: inside a step the splat backward (K1b) was the longest kernel (20 us of 94 at B=32) and what it waits for are its
// scattered 16-byte gathers from dL/d(raw) (ncu, profiles/r02_a_l1tex_splat.csv: 4.7 L2 sectors per point, 0.4 sectors /
// ns / SM, 43 % of the warp slots) -- nothing the x/y pipeline of the backward uses: that one is bound by TMA, tensor
// core and its coalesced stores.  So the gathers run UNDER the pipeline, on the same SMs:
//   warps 0-11   the x/y pipeline exactly as in dpc_tcp_conv_xy_kernel<false,false,true,false> (producers 0-3, consumers 4-11)
//   warp  12     MMA issue + TMA loads (one elected lane)
//   warp  13     signaller: once the eight consumer warps have stored a tile it makes those stores visible at GPU scope
//                (the fence is paid by a thread that has nothing else to do, not by the consumers) and adds 1 to the
//                tile's per-sample counter in global memory
//   warps 14..   gather warps: item = 32 consecutive points of a sample: cell from tr_pc (the forward's own output, so
//                the same cell), wait until the sample's 32 tiles are published, gather the 8 corners of dL/d(raw) as
//                four 16-byte loads, weights' derivative -> dL/d(tr_pc) of the point, 12 bytes out
// and the chain rule through the camera (+ pose-gradient sums) is left to dpc_splat_bwd_kernel, which then has no
// gathers to wait for (d_vox = NULL, d_tr_pc_in = this kernel's output; ~4 us instead of 20).
// A first version kept the whole splat backward in the gather warps: correct, but 600 dependent instructions per item
// on 2.5 warps per scheduler took 18 us of ALU latency alone, and one atomic per pose component per item another 11-17 us
// (profiles/r02_e_fused_roles.txt) -- the ALU-heavy half belongs in a kernel that fills the SM.
// Register file: the kernel starts at 96 (640 threads) / 80 (768 threads) registers per thread; THAT allocation (61440)
// is the pool setmaxnreg works in -- not the SM's 64 K (a version that tried to end up with 65536 hung in
// setmaxnreg.inc) -- and setmaxnreg moves registers from warps 12.. (48 each) to the pipeline warps 0-11 (128 / 112 each).
// Tiles are handed out in sample order (tile = blockIdx.x + i * gridDim.x, sample = tile / 32) and so are the gather
// items, so the gathers of sample b overlap the pipeline's work on samples b+4...; only the last few samples' gathers
// run after the pipeline has drained.
#pragma once
#ifndef DPC_EMU

#define DPC_XYG_SIGNAL 13
#define DPC_XYG_G0 14
#define DPC_XYG_AUX_REGS 48

struct DpcXYGatherArgs {
  const float* tr_pc;        // [B,N,3] camera-space points as the forward wrote them (depth, y, x)
  const float* d_raw;        // == the grid this kernel's pipeline writes (dL/d raw)
  const float* g_tr_pc;      // optional upstream gradient at tr_pc, added in
  float* d_tr;               // [B,N,3] out: dL/d(tr_pc)
  unsigned* sample_cnt;      // [B], zero on entry
  int B, N;
  int dbg;                   // experiments (dpc_debug_set(16, .)): 1 = do not wait, 2 = no gathers, 4 = gather warps idle, 8 = no fence
};

DPC_DEV void dpc_xyg_stamp(int id, int slot, bool leader) {
  if (leader && dpc_kt_on) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (slot < 2) atomicMin(&dpc_kt[id * 4 + slot], t); else atomicMax(&dpc_kt[id * 4 + slot], t);
  }
}

// One gather warp: items gw, gw + nw, ... of the B * ceil(N / 32) chunks, in sample order.  What an item needs from
// global memory before its gathers -- the sample's counter and the 32 points -- is loaded while the previous item is
// being processed.
DPC_DEV void dpc_xyg_gather_warp(const DpcXYGatherArgs& g, unsigned target, unsigned gw, unsigned nw) {
  constexpr int V = 64, Vz = 64;
  const int lane = threadIdx.x & 31;
  const unsigned chunks = (unsigned)(g.N + 31) >> 5;
  const unsigned items = (unsigned)g.B * chunks;
  if (g.dbg & 4) return;
  unsigned n_seen = 0;
  float n_z = 0.f, n_y = 0.f, n_x = 0.f;
  auto prefetch = [&](unsigned w) {
    const int b = (int)(w / chunks), c = (int)(w - (unsigned)b * chunks);
    // relaxed: the value only decides whether the acquire loop below can be skipped; the gathers are issued behind a
    // branch on it and bypass L1 (ld.global.cg), so they cannot be served before the counter was
    if (lane == 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(n_seen) : "l"(g.sample_cnt + b) : "memory");
    const int i = c * 32 + lane;
    n_z = n_y = n_x = 2.0f;            // outside the cube: an invalid point
    if (i < g.N) {
      const float* pp = g.tr_pc + ((size_t)b * g.N + i) * 3;
      n_z = __ldg(pp); n_y = __ldg(pp + 1); n_x = __ldg(pp + 2);
    }
  };
  unsigned long long t_wait = 0;
  dpc_xyg_stamp(9, 0, lane == 0);
  if (gw < items) prefetch(gw);
  for (unsigned w = gw; w < items; w += nw) {
    const int b = (int)(w / chunks), c = (int)(w - (unsigned)b * chunks);
    unsigned seen = n_seen;
    const float z = n_z, y = n_y, x = n_x;
    if (w + nw < items) prefetch(w + nw);
    const int i = c * 32 + lane;
    const bool live = i < g.N;
    const size_t pi = (size_t)b * g.N + i;
    const DpcCell cell = dpc_cell(z, y, x, Vz, V);
    // wait until every tile of sample b has been published by its CTA's signaller
    if (lane == 0 && !(g.dbg & 1)) {
      unsigned long long t0 = 0;
      if (dpc_kt_on && seen < target) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
      while (seen < target) { __nanosleep(100); seen = dpc_ld_acquire_gpu(g.sample_cnt + b); }
      if (t0) { unsigned long long t1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1)); t_wait += t1 - t0; }
    }
    __syncwarp();
    float dw[8];
    if (g.dbg & 2) {
#pragma unroll
      for (int q = 0; q < 8; ++q) dw[q] = 0.0f;
    } else {
      dpc_gather_corners<true>(g.d_raw + (size_t)b * Vz * V * V, cell, Vz, V, dw);
    }
    // weights' derivative (same formulas as dpc_splat_bwd_kernel)
    float gz = 0.f, gy = 0.f, gx = 0.f;
    if (cell.valid) {
      const float wz[2] = {1.0f - cell.rz, cell.rz}, wy[2] = {1.0f - cell.ry, cell.ry}, wx[2] = {1.0f - cell.rx, cell.rx};
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
          for (int ii = 0; ii < 2; ++ii) {
            if (cell.iz + k < Vz && cell.iy + jj < V && cell.ix + ii < V) {
              const float dwc = dw[k * 4 + jj * 2 + ii];
              gz += (k ? dwc : -dwc) * (wy[jj] * wx[ii]);
              gy += (jj ? dwc : -dwc) * (wz[k] * wx[ii]);
              gx += (ii ? dwc : -dwc) * (wz[k] * wy[jj]);
            }
          }
      gz *= (float)(Vz - 1); gy *= (float)(V - 1); gx *= (float)(V - 1);
    }
    if (live) {
      if (g.g_tr_pc) { gz += g.g_tr_pc[pi * 3 + 0]; gy += g.g_tr_pc[pi * 3 + 1]; gx += g.g_tr_pc[pi * 3 + 2]; }
      g.d_tr[pi * 3 + 0] = gz; g.d_tr[pi * 3 + 1] = gy; g.d_tr[pi * 3 + 2] = gx;
    }
  }
  dpc_xyg_stamp(9, 3, lane == 0);
  if (lane == 0 && dpc_kt_on) atomicAdd(&dpc_kt[10 * 4 + 2], t_wait);     // summed over all gather warps
}

// a, xymap, K, pl, ntiles, ht: as dpc_tcp_conv_xy_kernel (always: no input clip, saved clip mask applied to the output).
// NT = 640 (6 gather warps, pipeline warps at 128 registers), 768 (10 gather warps, pipeline warps at 112) or 448 (no
// gather warps, no setmaxnreg: the pipeline + the per-sample signalling alone, to price the signalling).
template <int NT>
__global__ void __launch_bounds__(NT, 1)
dpc_tcp_conv_xy_gather_kernel(const __grid_constant__ DpcConvXY64Args a, const __grid_constant__ CUtensorMap xymap, int K, int pl,
                              int ntiles, const __grid_constant__ DpcTcpTaps ht, const __grid_constant__ DpcXYGatherArgs g) {
  constexpr int V = 64;
  constexpr int kt_id = DPC_KT_XY_B;
  constexpr int ns = DPC_TCP_NS;
  constexpr int NG = NT / 32 - DPC_XYG_G0;
  constexpr int PIPE_REGS = (NT == 640) ? 128 : 112;
  constexpr int START_REGS = (NT == 640) ? 96 : 80;
  static_assert(NT == 448 || 12 * PIPE_REGS + (NT / 32 - 12) * DPC_XYG_AUX_REGS <= (NT / 32) * START_REGS, "setmaxnreg: the CTA's launch allocation is the pool");
  __shared__ unsigned G_stored;         // number of (consumer warp, tile) stores completed in this CTA
  if (threadIdx.x == 0) G_stored = 0u;
  DPC_TCP_SETUP(a.taps_x, K, pl, a.rev);
  if (NT != 448) {
    if (warp < DPC_TCP_ISSUER) {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PIPE_REGS));
    } else {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DPC_XYG_AUX_REGS));
    }
  }
  if (warp < 4) {
    // ---------------- producers: thread = row (slice, y)
    const int m = tid;
    int i = 0, slot = 0, sph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
      const int s = i & 1, k = i >> 1;
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
        dpc_tcp_put_a(tmem, s, h, v);
      }
      dpc_tcp_warp_arrive(&B.sfree[slot]);
      dpc_tc_wait_st();
      dpc_tcp_warp_arrive(&B.opfull[s]);
      if (++slot == ns) { slot = 0; sph ^= 1; }
    }
  } else if (warp < DPC_TCP_ISSUER) {
    // ---------------- consumers (see dpc_tcp_conv_xy_kernel)
    const int c = tid - DPC_TCP_NPROD, m = c & 127, h = c >> 7;
    const int sl = m >> 6, rx = m & 63;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    auto drain = [&](int j, int tile) {
      const int s = j & 1, k = j >> 1;
      const size_t base = (size_t)tile * (2 * V * V);
      const uint32_t mw = a.mask_in[(base >> 5) + (size_t)sl * 128 + 2 * (32 * h + lane) + ((m >> 5) & 1)];
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
        const uint32_t w = __shfl_sync(DPC_FULL, mw, q);
        if (!((w >> lane) & 1u)) v = 0.0f;
        dst[q * V] = v;
      }
      __syncwarp();
      if (lane == 0 && !(g.dbg & 32)) dpc_red_release_cta_shared(&G_stored, 1u);     // this warp's part of tile j is stored
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
      dpc_tcp_put_a(tmem, s, h, r);
      dpc_tc_wait_st();
      dpc_tcp_warp_arrive(&B.opfull2[s]);
      if (prev_tile >= 0) drain(i - 1, prev_tile);      // overlaps GEMM 2 of tile i
      prev_tile = tile;
    }
    if (prev_tile >= 0) drain(i - 1, prev_tile);
    dpc_xyg_stamp(8, 3, tid == DPC_TCP_NPROD);        // this CTA's pipeline is done
  } else if (warp == DPC_TCP_ISSUER) {
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
      auto gemm1 = [&](int j, int tile) {
        const int s = j & 1, k = j >> 1;
        dpc_mbar_wait(&B.opfull[s], k & 1);
        dpc_tc_fence_after();
        dpc_tc_issue_ts(tmem + DPC_TCP_AHI(s), tmem + DPC_TCP_ALO(s), sbase + DPC_TCP_T_OFF, tmem + DPC_TCP_D1(s), &B.done[s]);
        const int nt = tile + ns * step;
        if (nt < ntiles) { dpc_mbar_wait(&B.sfree[slot], sph); fill(slot, nt); }
        if (++slot == ns) { slot = 0; sph ^= 1; }
      };
      auto gemm2 = [&](int j) {
        const int s = j & 1, k = j >> 1;
        dpc_mbar_wait(&B.opfull2[s], k & 1);
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
  } else if (warp == DPC_XYG_SIGNAL) {
    // ---------------- signaller: publish finished tiles.  Every consumer warp stores tile i before tile i+1 and the
    // consumers meet at a named barrier once per tile, so stored >= 8 (i + 1) <=> all eight have stored tile i.
    if (lane == 0 && !(g.dbg & 32)) {
      int i = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
        while (dpc_ld_acquire_cta_shared(&G_stored) < 8u * (unsigned)(i + 1)) __nanosleep(64);
        if (!(g.dbg & 8)) __threadfence();       // cumulative: the consumers' stores (observed through G.stored) are visible GPU-wide first
        atomicAdd(g.sample_cnt + (tile >> 5), 1u);
      }
    }
    __syncwarp();
  } else if (NG > 0) {
    // ---------------- gather warps
    dpc_xyg_gather_warp(g, 32u, blockIdx.x * NG + (unsigned)(warp - DPC_XYG_G0), gridDim.x * NG);
  }
  DPC_TCP_TEARDOWN();
}

// Host side.  Returns DPC_OK after the launch; the caller has made sure the shapes are the pipeline's (V = Vz = 64).
static inline int dpc_tcp_conv_xy_gather_launch(float* grid, const float* taps, int K, int pl, int64_t nslices,
                                                const uint32_t* mask_in, int rev, const float* host_taps,
                                                const DpcXYGatherArgs& g, int wide, void* stream) {
  if ((((uintptr_t)grid) & 15u) != 0) return DPC_ERR_ARG;
  if (!mask_in || !g.sample_cnt || !g.tr_pc || !g.d_tr || (nslices % 64) != 0) return DPC_ERR_ARG;
  if ((int64_t)g.B * ((g.N + 31) / 32) > 2147483647LL) return DPC_ERR_SHAPE;
  DpcConvXY64Args a;
  a.in = grid; a.out = grid; a.taps_x = taps; a.taps_y = taps; a.clip_in = 0; a.mask_out = nullptr; a.mask_in = mask_in;
  a.nslices = (int)nslices; a.rev = rev; a.zero_ptr = nullptr; a.dbg = 0;
  for (int i = 0; i < 24; ++i) { a.ht.px[i] = make_float2(0.f, 0.f); a.ht.dy[i] = make_float2(0.f, 0.f); }
  const int ntiles = (int)(nslices / 2), grid_x = ntiles < dpc_tc_sm_count() ? ntiles : dpc_tc_sm_count();
  CUtensorMap xymap;
  if (dpc_tc_make_xymap(&xymap, grid, nslices) != DPC_OK) return DPC_ERR_CUDA;
  dpc_tcp_host_taps_next = host_taps;
  const DpcTcpTaps ht = dpc_tcp_take_host_taps(taps ? K : 0);
  if (wide == 2) {     // experiment: no gather warps -- the caller runs the splat backward as its own kernel afterwards
    if (cudaFuncSetAttribute(dpc_tcp_conv_xy_gather_kernel<448>, cudaFuncAttributeMaxDynamicSharedMemorySize, DPC_TCP_SMEM_BYTES) != cudaSuccess)
      return DPC_ERR_CUDA;
    DPC_LAUNCH(dpc_tcp_conv_xy_gather_kernel<448>, dim3(grid_x), dim3(448), (size_t)DPC_TCP_SMEM_BYTES, stream, a, xymap, K, pl, ntiles, ht, g);
  } else if (wide) {
    if (cudaFuncSetAttribute(dpc_tcp_conv_xy_gather_kernel<768>, cudaFuncAttributeMaxDynamicSharedMemorySize, DPC_TCP_SMEM_BYTES) != cudaSuccess)
      return DPC_ERR_CUDA;
    DPC_LAUNCH(dpc_tcp_conv_xy_gather_kernel<768>, dim3(grid_x), dim3(768), (size_t)DPC_TCP_SMEM_BYTES, stream, a, xymap, K, pl, ntiles, ht, g);
  } else {
    if (cudaFuncSetAttribute(dpc_tcp_conv_xy_gather_kernel<640>, cudaFuncAttributeMaxDynamicSharedMemorySize, DPC_TCP_SMEM_BYTES) != cudaSuccess)
      return DPC_ERR_CUDA;
    DPC_LAUNCH(dpc_tcp_conv_xy_gather_kernel<640>, dim3(grid_x), dim3(640), (size_t)DPC_TCP_SMEM_BYTES, stream, a, xymap, K, pl, ntiles, ht, g);
  }
  return DPC_OK;
}

#endif  // !DPC_EMU
