// x/y pass AND depth pass (+ scale / clip / projection) of the forward in ONE persistent kernel (64^3 grids).
//
// Between two pipeline kernels sit ~1 us of launch gap and ~2.5 us of start-up that PDL cannot hide (two 226 KB / 512-TMEM-
// column CTAs never share an SM: TMEM allocation, barrier init, Toeplitz build and the first TMA round trip all wait for
// the previous kernel's CTA to leave), plus the ramp-down of the first kernel behind its slowest CTA.  Here every CTA runs
// its x/y tiles and then its depth tiles on the same TMEM allocation / operand / staging ring:
//   phase A  = dpc_tcp_conv_xy_kernel<clip, mask out> in place on the raw grid; each finished tile is counted towards its
//              sample (consumer warps -> shared counter -> the signaller warp's GPU-scope release on sample_cnt[b])
//   barrier  = the 13 pipeline warps of THIS CTA only (its own MMAs have drained; other CTAs are not waited for)
//   phase B  = dpc_tcp_conv_z_fwd_kernel; the loading lane acquires sample_cnt[b] == 32 before the TMA box of a tile of
//              sample b (generic-proxy stores of other SMs -> async-proxy read: fence.proxy.async behind the acquire)
// Tiles are handed out in sample order in both phases, so by the time a CTA reaches phase B its first samples have long been
// complete; only the last samples' tiles wait for the slowest CTA of phase A.  sample_cnt is zeroed by the splat kernel.
#pragma once
#ifndef DPC_EMU

#define DPC_FXZ_THREADS 448
#define DPC_FXZ_SIGNAL 13

template <int MODE, bool HAS_S>
__global__ void __launch_bounds__(DPC_FXZ_THREADS, 1)
dpc_tcp_fwd_xyz_kernel(const __grid_constant__ DpcConvXY64Args a, const __grid_constant__ CUtensorMap xymap, int K, int pl, int ntiles,
                       const __grid_constant__ DpcTcpTaps ht, const __grid_constant__ DpcConvZArgs az,
                       const __grid_constant__ CUtensorMap zmap, unsigned* sample_cnt) {
  constexpr int V = 64, Vz = 64;
  constexpr bool CLAMPU = (MODE == DPC_PROJ_DRC);
  constexpr int kt_id = DPC_KT_XY_F;
  constexpr int ns = DPC_TCP_NS;
  __shared__ unsigned G_stored;
  __shared__ __align__(8) DpcTcpBars B2;        // phase B's barriers (fresh phases)
  __shared__ __align__(8) float2 comb[128];
  if (threadIdx.x == 0) G_stored = 0u;
  DPC_TCP_SETUP(a.taps_x, K, pl, a.rev);
  // (B2 is initialised behind the set-up barrier by thread 0 and published by the phase barrier below)
  if (tid == 0) {
    for (int q = 0; q < DPC_TCP_NS; ++q) { dpc_mbar_init(&B2.sfull[q], 1); dpc_mbar_init(&B2.sfree[q], 4); }
    for (int q = 0; q < 2; ++q) { dpc_mbar_init(&B2.opfull[q], 4); dpc_mbar_init(&B2.done[q], 1); dpc_mbar_init(&B2.accfree[q], 8); }
  }
  // =========================================================================================== phase A: x and y passes
  if (warp < 4) {
    const int m = tid;
    int i = 0, slot = 0, sph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
      const int s = i & 1, k = i >> 1;
      const size_t base = (size_t)tile * (2 * V * V);
      dpc_mbar_wait(&B.sfull[slot], sph);
      if (k >= 1) dpc_mbar_wait(&B.done2[s], (k - 1) & 1);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const unsigned char* row = sm + DPC_TCP_S_OFF + (uint32_t)slot * DPC_TCP_SLOT + (uint32_t)h * 16384u + (uint32_t)m * 128u;
        float v[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 f = *reinterpret_cast<const float4*>(row + ((q ^ (m & 7)) << 4));
          v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
        }
        {       // these 32 voxels are exactly one word of the clip-pass bit plane
          uint32_t word = 0u;
#pragma unroll
          for (int j = 0; j < 32; ++j) word |= (v[j] >= 0.0f && v[j] <= 1.0f) ? (1u << j) : 0u;
          a.mask_out[(base >> 5) + 2 * m + h] = word;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = dpc_clip01(v[j]);
        dpc_tcp_put_a(tmem, s, h, v);
      }
      dpc_tcp_warp_arrive(&B.sfree[slot]);
      dpc_tc_wait_st();
      dpc_tcp_warp_arrive(&B.opfull[s]);
      if (++slot == ns) { slot = 0; sph ^= 1; }
    }
  } else if (warp < DPC_TCP_ISSUER) {
    const int c = tid - DPC_TCP_NPROD, m = c & 127, h = c >> 7;
    const int sl = m >> 6, rx = m & 63;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    auto drain = [&](int j, int tile) {
      const int s = j & 1, k = j >> 1;
      const size_t base = (size_t)tile * (2 * V * V);
      dpc_mbar_wait(&B.done2[s], k & 1);
      dpc_tc_fence_after();
      float r[32];
      dpc_tc_ld32(tmem + DPC_TCP_D2(s) + (uint32_t)(h * 32) + lane_addr, r);
      dpc_tc_wait_ld();
      dpc_tc_fence_before();
      float* dst = a.out + base + (size_t)sl * V * V + (size_t)(32 * h) * V + rx;
#pragma unroll
      for (int q = 0; q < 32; ++q) dst[q * V] = r[q];
      __syncwarp();
      if (lane == 0) dpc_red_release_cta_shared(&G_stored, 1u);
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
      if (prev_tile >= 0) drain(i - 1, prev_tile);
      prev_tile = tile;
    }
    if (prev_tile >= 0) drain(i - 1, prev_tile);
  } else if (warp == DPC_TCP_ISSUER) {
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
  } else {
    // signaller: publish this CTA's finished x/y tiles to the other CTAs (see dpc_smooth_tcp.cuh, "cross-CTA hand-over")
    if (lane == 0) {
      int i = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
        while (dpc_ld_acquire_cta_shared(&G_stored) < 8u * (unsigned)(i + 1)) __nanosleep(64);
        __threadfence();
        atomicAdd(sample_cnt + (tile >> 5), 1u);
      }
    }
    __syncwarp();
  }
  // ------------------------------------------------------------------------------------------- this CTA's pipeline has drained
  if (warp <= DPC_TCP_ISSUER) {
    dpc_tc_fence_before();
    dpc_named_bar(2, DPC_TCP_THREADS);
    dpc_tc_fence_after();
  }
  dpc_kt_mark(DPC_KT_XY_F, 3);
  dpc_kt_mark(DPC_KT_Z_F, 0);
  dpc_kt_mark(DPC_KT_Z_F, 1);
  // =========================================================================================== phase B: depth pass + projection
  if (warp < 4) {
    const int m = tid;
    int i = 0, slot = 0, sph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
      const int s = i & 1, k = i >> 1;
      dpc_mbar_wait(&B2.sfull[slot], sph);
      if (k >= 1) dpc_mbar_wait(&B2.done[s], (k - 1) & 1);
      const float* stg = reinterpret_cast<const float*>(sm + DPC_TCP_S_OFF + (uint32_t)slot * DPC_TCP_SLOT) + m;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[32];
#pragma unroll
        for (int z = 0; z < 32; ++z) v[z] = stg[(32 * h + z) * 128];
        dpc_tcp_put_a(tmem, s, h, v);
      }
      dpc_tcp_warp_arrive(&B2.sfree[slot]);
      dpc_tc_wait_st();
      dpc_tcp_warp_arrive(&B2.opfull[s]);
      if (++slot == DPC_TCP_NS) { slot = 0; sph ^= 1; }
    }
  } else if (warp < DPC_TCP_ISSUER) {
    const int c = tid - DPC_TCP_NPROD, m = c & 127, ch = c >> 7;
    const DpcDrc D = dpc_drc_consts(az.mode, az.eps);
    int i = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
      const int s = i & 1, k = i >> 1;
      const int b = tile >> 5, y = (tile & 31) * 2 + (m >> 6), x = m & 63;
      const float sc = HAS_S ? az.scale[b] : 1.0f;
      dpc_mbar_wait(&B2.done[s], k & 1);
      dpc_tc_fence_after();
      float r[32];
      dpc_tc_ld32(tmem + DPC_TCP_D1(s) + (uint32_t)(ch * 32) + ((uint32_t)((warp & 3) * 32) << 16), r);
      dpc_tc_wait_ld();
      dpc_tcp_warp_arrive(&B2.accfree[s]);
      const size_t ray = ((size_t)b * Vz * V + y) * V + x;
      float* vout = az.vox_out + ray + (size_t)(32 * ch) * V * V;
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
      if (HAS_S && az.mask2_out) az.mask2_out[(((size_t)b * V + y) * V + x) * 2 + ch] = mw;
      if (MODE != DPC_PROJ_NONE) {
        const float Sh = fmaf(T0, S1, S0), Th = T0 * T1;
        if (ch == 1) comb[m] = (MODE == DPC_PROJ_MAX) ? make_float2(mx, 0.f) : make_float2(Th, Sh);
        dpc_named_bar(1, 256);
        if (ch == 0) {
          const float2 hi = comb[m];
          const float out = (MODE == DPC_PROJ_MAX) ? fmaxf(mx, hi.x) : fmaf(Th, hi.y, Sh);
          const int yo = az.flip_y ? (V - 1 - y) : y;
          az.proj[((size_t)b * V + yo) * V + x] = out;
        }
        dpc_named_bar(1, 256);
      }
    }
  } else if (warp == DPC_TCP_ISSUER) {
    if (dpc_elect_one()) {
      int ready_b = -1;       // samples <= ready_b are known to be complete
      auto fill = [&](int slot, int tile) {
        const int b = tile >> 5;
        if (b > ready_b) {
          while (dpc_ld_acquire_gpu(sample_cnt + b) < 32u) __nanosleep(64);
          asm volatile("fence.proxy.async.global;" ::: "memory");     // other SMs' generic-proxy stores -> this TMA read
          ready_b = b;
        }
        dpc_mbar_expect_tx(&B2.sfull[slot], DPC_TCP_SLOT);
        dpc_tma_load_2d(sm + DPC_TCP_S_OFF + (uint32_t)slot * DPC_TCP_SLOT, &zmap, (tile & 31) * 128, (tile >> 5) * 64, &B2.sfull[slot]);
      };
      for (int j = 0; j < DPC_TCP_NS; ++j) { const int t = (int)blockIdx.x + j * step; if (t < ntiles) fill(j, t); }
      int i = 0, slot = 0, sph = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += step, ++i) {
        const int s = i & 1, k = i >> 1;
        dpc_mbar_wait(&B2.opfull[s], k & 1);
        if (k >= 1) dpc_mbar_wait(&B2.accfree[s], (k - 1) & 1);
        dpc_tc_fence_after();
        dpc_tc_issue_ts(tmem + DPC_TCP_AHI(s), tmem + DPC_TCP_ALO(s), sbase + DPC_TCP_T_OFF, tmem + DPC_TCP_D1(s), &B2.done[s]);
        const int nt = tile + DPC_TCP_NS * step;
        if (nt < ntiles) { dpc_mbar_wait(&B2.sfree[slot], sph); fill(slot, nt); }
        if (++slot == DPC_TCP_NS) { slot = 0; sph ^= 1; }
      }
    }
    __syncwarp();
  }
  dpc_tc_fence_before();
  __syncthreads();
  dpc_kt_mark(DPC_KT_Z_F, 3);
  if (warp == 0) { __syncwarp(); dpc_tc_dealloc(tmem, 512); }
}

// Host side: in place on `grid` (raw -> x/y-smoothed), then the depth pass from it.  taps: one tap vector for all three
// axes (K taps, pad pl); az: the depth pass's arguments with az.in == grid.  sample_cnt[B]: ZERO on entry.
template <int MODE, bool HAS_S>
static inline int dpc_tcp_fwd_xyz_go(float* grid, const float* taps, int K, int pl, uint32_t* mask1_out, const float* host_taps,
                                     const DpcConvZArgs& az, unsigned* sample_cnt, void* stream) {
  const int64_t nslices = (int64_t)az.B * 64;
  if ((((uintptr_t)grid) & 15u) != 0 || !mask1_out || !sample_cnt) return DPC_ERR_ARG;
  DpcConvXY64Args a;
  a.in = grid; a.out = grid; a.taps_x = taps; a.taps_y = taps; a.clip_in = 1; a.mask_out = mask1_out; a.mask_in = nullptr;
  a.nslices = (int)nslices; a.rev = 0; a.zero_ptr = nullptr; a.dbg = 0;
  for (int i = 0; i < 24; ++i) { a.ht.px[i] = make_float2(0.f, 0.f); a.ht.dy[i] = make_float2(0.f, 0.f); }
  const int ntiles = (int)(nslices / 2), grid_x = ntiles < dpc_tc_sm_count() ? ntiles : dpc_tc_sm_count();
  CUtensorMap xymap, zmap;
  if (dpc_tc_make_xymap(&xymap, grid, nslices) != DPC_OK || dpc_tc_make_zmap(&zmap, grid, az.B) != DPC_OK) return DPC_ERR_CUDA;
  dpc_tcp_host_taps_next = host_taps;
  const DpcTcpTaps ht = dpc_tcp_take_host_taps(taps ? K : 0);
  if (cudaFuncSetAttribute(dpc_tcp_fwd_xyz_kernel<MODE, HAS_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, DPC_TCP_SMEM_BYTES) != cudaSuccess)
    return DPC_ERR_CUDA;
  DPC_LAUNCH((dpc_tcp_fwd_xyz_kernel<MODE, HAS_S>), dim3(grid_x), dim3(DPC_FXZ_THREADS), (size_t)DPC_TCP_SMEM_BYTES, stream, a, xymap, K, pl,
             ntiles, ht, az, zmap, sample_cnt);
  return DPC_OK;
}
static inline int dpc_tcp_fwd_xyz_launch(float* grid, const float* taps, int K, int pl, uint32_t* mask1_out, const float* host_taps,
                                         const DpcConvZArgs& az, unsigned* sample_cnt, void* stream) {
  const bool hs = az.scale != nullptr;
  switch (az.mode) {
    case DPC_PROJ_DRC: return hs ? dpc_tcp_fwd_xyz_go<DPC_PROJ_DRC, true>(grid, taps, K, pl, mask1_out, host_taps, az, sample_cnt, stream)
                                 : dpc_tcp_fwd_xyz_go<DPC_PROJ_DRC, false>(grid, taps, K, pl, mask1_out, host_taps, az, sample_cnt, stream);
    case DPC_PROJ_MAX: return hs ? dpc_tcp_fwd_xyz_go<DPC_PROJ_MAX, true>(grid, taps, K, pl, mask1_out, host_taps, az, sample_cnt, stream)
                                 : dpc_tcp_fwd_xyz_go<DPC_PROJ_MAX, false>(grid, taps, K, pl, mask1_out, host_taps, az, sample_cnt, stream);
    default: return hs ? dpc_tcp_fwd_xyz_go<DPC_PROJ_DRC_PROD, true>(grid, taps, K, pl, mask1_out, host_taps, az, sample_cnt, stream)
                       : dpc_tcp_fwd_xyz_go<DPC_PROJ_DRC_PROD, false>(grid, taps, K, pl, mask1_out, host_taps, az, sample_cnt, stream);
  }
}

#endif  // !DPC_EMU
