// Lab build only: micro-benchmark of the splat backward's memory access pattern, WITHOUT the kernel around it.
// Every thread gathers the 2 x 2 rows (z, y) of a uniformly random cell of its sample's V^3 grid, the way
// dpc_splat_bwd_kernel does, adds what it read and stores one float.  What the splat backward pays for its gathers is
// bounded below by the time of this kernel at the same grid / CTA shape; the variants separate the candidates for the
// bound (wavefronts of the SM's L1 miss path, sectors requested from L2, round trips in a dependent chain, occupancy).
#pragma once
#include "dpc_common.cuh"
#include "dpc_splat.cuh"

#ifndef DPC_EMU
struct DpcGatherBenchArgs {
  const float* grid;   // [B, V, V, V]
  float* out;          // [B, N]
  int N, V;
  int ppt;             // points per thread, one after the other
  int share_log2;      // lanes in groups of 2^share_log2 read the SAME cell (fewer distinct lines per instruction)
  unsigned seed;
  unsigned zero;       // 0 at run time (opaque to the compiler): builds the dependent chain of variant 3
  int cg;              // loads through ld.global.cg
};

DPC_DEV float4 dpc_gb_ld4(const float* p, int cg) {
  return cg ? __ldcg(reinterpret_cast<const float4*>(p)) : __ldg(reinterpret_cast<const float4*>(p));
}
DPC_DEV float dpc_gb_ld1_pred(const float* p, bool pred) {
  float v = 0.0f;
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
               : "+f"(v) : "l"(p), "r"((unsigned)pred));
  return v;
}

// VAR 0: four 16-byte loads (the aligned 4-voxel group of ix in each row) + four scalar loads PREDICATED on the
//        straddling lanes (ix % 4 == 3), all independent: the access pattern of the product kernel, ideal issue
//     1: the four 16-byte loads only          2: eight scalar loads (round 1's first kernel)
//     3: the four 16-byte loads as a DEPENDENT chain (each address waits for the previous result)
//     4: four 16-byte + four scalar loads, all lanes (dpc_gather_corners: 8 wavefronts per lane)
//     5: two 16-byte loads                    6: one 16-byte load
//     7: the product pattern through cp.async (LDGSTS: global -> shared without destination registers), read back from
//        shared memory after cp.async.wait_group 0
template <int VAR>
__global__ void dpc_gather_bench_kernel(DpcGatherBenchArgs a) {
  __shared__ __align__(16) float4 sq[VAR == 7 ? 4 : 1][VAR == 7 ? 256 : 1];
  __shared__ float se[VAR == 7 ? 4 : 1][VAR == 7 ? 256 : 1];
  const int b = blockIdx.y, V = a.V;
  const float* dv = a.grid + (size_t)b * V * V * V;
  float accum = 0.0f;
  for (int r = 0; r < a.ppt; ++r) {
    const int i = (blockIdx.x * a.ppt + r) * blockDim.x + threadIdx.x;
    if (i >= a.N) break;
    const uint32_t h = dpc_mix32(((uint32_t)(i >> a.share_log2) * 0x9E3779B9u) ^ ((uint32_t)b * 0x85EBCA6Bu) ^ a.seed);
    const int ix = (int)(((h & 0x3ffu) * (uint32_t)(V - 1)) >> 10);
    const int iy = (int)((((h >> 10) & 0x3ffu) * (uint32_t)(V - 1)) >> 10);
    const int iz = (int)((((h >> 20) & 0x3ffu) * (uint32_t)(V - 1)) >> 10);
    const int o4 = ix & 3;
    const float* g0 = dv + (iz * V + iy) * V + (ix - o4);
    if (VAR == 7) {
      const int tl = threadIdx.x & 255;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float* rp = g0 + ((c >> 1) * V + (c & 1)) * V;
        dpc_cp_async16(&sq[VAR == 7 ? c : 0][VAR == 7 ? tl : 0], rp, true);
        dpc_cp_async4(&se[VAR == 7 ? c : 0][VAR == 7 ? tl : 0], o4 == 3 ? rp + 4 : dv, o4 == 3);
      }
      dpc_cp_async_commit();
      dpc_cp_async_wait<0>();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 q = sq[VAR == 7 ? c : 0][VAR == 7 ? tl : 0];
        accum += (o4 == 0 ? q.x : (o4 == 1 ? q.y : (o4 == 2 ? q.z : q.w))) + q.w + se[VAR == 7 ? c : 0][VAR == 7 ? tl : 0];
      }
    } else if (VAR == 2) {
      float s = 0.0f;
#pragma unroll
      for (int c = 0; c < 8; ++c) s += __ldg(dv + ((iz + (c >> 2)) * V + iy + ((c >> 1) & 1)) * V + ix + (c & 1));
      accum += s;
    } else if (VAR == 3) {
      float4 q = dpc_gb_ld4(g0, a.cg);
      unsigned dep = __float_as_uint(q.x) & a.zero;
      q = dpc_gb_ld4(g0 + V + dep, a.cg);
      accum += q.y;
      dep = __float_as_uint(q.x) & a.zero;
      q = dpc_gb_ld4(g0 + V * V + dep, a.cg);
      accum += q.z;
      dep = __float_as_uint(q.x) & a.zero;
      q = dpc_gb_ld4(g0 + V * V + V + dep, a.cg);
      accum += q.w;
    } else {
      constexpr int NQ = VAR == 5 ? 2 : (VAR == 6 ? 1 : 4);
      float4 q[NQ];
#pragma unroll
      for (int c = 0; c < NQ; ++c) q[c] = dpc_gb_ld4(g0 + ((c >> 1) * V + (c & 1)) * V, a.cg);
      float e[4] = {0.f, 0.f, 0.f, 0.f};
      if (VAR == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) e[c] = dpc_gb_ld1_pred(g0 + 4 + ((c >> 1) * V + (c & 1)) * V, o4 == 3);
      } else if (VAR == 4) {
#pragma unroll
        for (int c = 0; c < 4; ++c) e[c] = __ldg(g0 + 4 + ((c >> 1) * V + (c & 1)) * V);
      }
#pragma unroll
      for (int c = 0; c < NQ; ++c) accum += (o4 == 0 ? q[c].x : (o4 == 1 ? q[c].y : (o4 == 2 ? q[c].z : q[c].w))) + q[c].w;
#pragma unroll
      for (int c = 0; c < 4; ++c) accum += e[c];
    }
  }
  const int i0 = blockIdx.x * a.ppt * blockDim.x + threadIdx.x;
  if (i0 < a.N) a.out[(size_t)b * a.N + i0] = accum;
}
#endif
