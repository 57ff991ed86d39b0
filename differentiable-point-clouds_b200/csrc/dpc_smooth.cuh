// K2 / K3, generic variants (any V <= 128, any tap count <= 63, any pad): separable Gaussian
// smoothing and the depth-axis projection.  The shape-specialised FFMA2 kernels for the
// benchmark shapes live in dpc_smooth_fast.cuh; these are the always-available path and the
// in-library cross-check for them.
//
//   conv_xy : one CTA per depth slice: [clip] -> x-correlation -> y-correlation -> [* mask]
//             (point_cloud.py:240 + the first two tf.nn.conv3d of :141-142)
//   conv_z  : one CTA per (sample, group of image rows), one THREAD per ray: depth
//             correlation (third conv3d) -> * scale -> clip (point_cloud.py:249-253) -> DRC /
//             max projection along the ray held by that thread (drc.py:47-123, point_cloud.py:265)
#pragma once
#include "dpc_common.cuh"

#define DPC_CONV_THREADS 256

// tap j of a K-tap filter: NULL taps = the identity filter (K = 1, used when the caller passes no
// smoothing kernel); rev = read the taps back to front (the transposed correlation of the backward).
DPC_DEV float dpc_tap(const float* taps, int K, int j, int rev) {
  return taps ? taps[rev ? (K - 1 - j) : j] : 1.0f;
}

// depth taps as FFMA2 operands, (t[j], t[j]), carried in the launch parameters (dpc_smooth_fast.cuh)
struct DpcTapsZ { float2 dz[24]; };

struct DpcConvXYArgs {
  const float* in; float* out;
  const float* taps_x; int Kx; int plx;
  const float* taps_y; int Ky; int ply;
  int B, Vz, V; int clip_in;
  uint32_t* mask_out; const uint32_t* mask_in;
  int rev;            // taps are read back to front
  float* zero_ptr;    // == in when the input is to be overwritten with zeros once it has been read
};

#ifndef DPC_EMU
__global__ void __launch_bounds__(DPC_CONV_THREADS)
#else
static void
#endif
dpc_conv_xy_kernel(DpcConvXYArgs a) {
  DPC_DYN_SMEM(float, sm);
  const int V = a.V, VV = V * V;
  float* A = sm;
  float* Bf = sm + VV;
  float* tx = sm + 2 * VV;
  float* ty = tx + DPC_MAX_TAPS + 1;
  const int tid = threadIdx.x;
  const size_t slice = (size_t)blockIdx.x * VV;
  if (tid < a.Kx) tx[tid] = dpc_tap(a.taps_x, a.Kx, tid, a.rev);
  if (tid < a.Ky) ty[tid] = dpc_tap(a.taps_y, a.Ky, tid, a.rev);
  dpc_grid_dep_sync();
  const int rounds = (VV + DPC_CONV_THREADS - 1) / DPC_CONV_THREADS;
  for (int r = 0; r < rounds; ++r) {
    const int i = r * DPC_CONV_THREADS + tid;
    float v = (i < VV) ? a.in[slice + i] : 0.0f;
    if (a.mask_out) {
      const unsigned bits = __ballot_sync(DPC_FULL, (i < VV) && (v >= 0.0f) && (v <= 1.0f));
      if ((tid & 31) == 0 && i < VV) a.mask_out[(slice + i) >> 5] = bits;
    }
    if (a.clip_in) v = dpc_clip01(v);
    if (i < VV) A[i] = v;
    if (a.zero_ptr && i < VV) a.zero_ptr[slice + i] = 0.0f;
  }
  __syncthreads();
  for (int i = tid; i < VV; i += DPC_CONV_THREADS) {
    const int y = i / V, x = i - y * V;
    float acc = 0.0f;
    for (int j = 0; j < a.Kx; ++j) {
      const int xx = x + j - a.plx;
      if (xx >= 0 && xx < V) acc = fmaf(tx[j], A[y * V + xx], acc);
    }
    Bf[i] = acc;
  }
  __syncthreads();
  for (int r = 0; r < rounds; ++r) {
    const int i = r * DPC_CONV_THREADS + tid;
    if (i >= VV) continue;
    const int y = i / V, x = i - y * V;
    float acc = 0.0f;
    for (int j = 0; j < a.Ky; ++j) {
      const int yy = y + j - a.ply;
      if (yy >= 0 && yy < V) acc = fmaf(ty[j], Bf[yy * V + x], acc);
    }
    if (a.mask_in) {
      const uint32_t wbits = a.mask_in[(slice + i) >> 5];
      if (!((wbits >> ((slice + i) & 31)) & 1u)) acc = 0.0f;
    }
    a.out[slice + i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
struct DpcConvZArgs {
  const float* in; const float* taps; int K; int pl; int rev;
  const float* scale; int mode; float eps; float cam_dist; float max_depth; int flip_y;
  int B, Vz, V, TY;
  float* vox_out; uint32_t* mask2_out; float* proj; float* probs; float* depth;
  DpcTapsZ ht; int use_ht;   // fast kernels: host-provided taps in the launch parameters (dpc_smooth_fast.cuh)
};

// DRC constants shared by forward and backward
struct DpcDrc {
  float lo, hi;     // clamp bounds for u (logsum) or (-inf, +inf)
  float c0, cZ;     // factors on the first and on the terminal event: e^eps (logsum quirk) or 1
  bool clampu;
};
DPC_DEV DpcDrc dpc_drc_consts(int mode, float eps) {
  DpcDrc d;
  if (mode == DPC_PROJ_DRC) {
    d.clampu = true; d.lo = eps; d.hi = (float)(1.0 - (double)eps);
    d.c0 = expf(eps); d.cZ = d.c0;
  } else {
    d.clampu = false; d.lo = 0.f; d.hi = 1.f; d.c0 = 1.f; d.cZ = 1.f;
  }
  return d;
}
DPC_DEV float dpc_psi(int i, int Vz, float cam_dist) {
  return __fadd_rn(__fsub_rn(__fdiv_rn((float)i, (float)Vz), 0.5f), cam_dist);
}

#ifndef DPC_EMU
__global__ void __launch_bounds__(128)
#else
static void
#endif
dpc_conv_z_fwd_kernel(DpcConvZArgs a) {
  DPC_DYN_SMEM(float, sm);
  const int V = a.V, Vz = a.Vz, TY = a.TY;
  const int b = blockIdx.y, y0 = blockIdx.x * TY;
  const int rows = min(TY, V - y0);
  const int RW = TY * V;              // tile row length (one z level)
  float* tile = sm;                   // [Vz][TY*V]
  float* taps = sm + (size_t)Vz * RW;
  const int tid = threadIdx.x;
  if (tid < a.K) taps[tid] = dpc_tap(a.taps, a.K, tid, a.rev);
  dpc_grid_dep_sync();
  const float* src = a.in + ((size_t)b * Vz * V + y0) * V;
  const int rowlen = rows * V;
  for (int z = 0; z < Vz; ++z)
    for (int i = tid; i < rowlen; i += blockDim.x) tile[z * RW + i] = src[(size_t)z * V * V + i];
  __syncthreads();
  if (tid >= rowlen) return;
  const int ty = tid / V, x = tid - ty * V, y = y0 + ty;
  const float* col = tile + tid;
  const bool has_s = a.scale != nullptr;
  const float s = has_s ? a.scale[b] : 1.0f;
  const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
  const int yo = a.flip_y ? (V - 1 - y) : y;
  const size_t ray = ((size_t)b * V + yo) * V + x;         // index into [B,V,V] outputs
  const size_t plane = (size_t)a.B * V * V;                // stride between drc_probs levels
  float T = 1.0f, proj = 0.0f, dep = 0.0f, mx = -INFINITY;
  uint32_t mbits = 0;
  const int nwords = (Vz + 31) >> 5;
  uint32_t* mrow = a.mask2_out ? a.mask2_out + (((size_t)b * V + y) * V + x) * nwords : nullptr;
  float* vout = a.vox_out + ((size_t)b * Vz * V + y) * V + x;
  for (int z = 0; z < Vz; ++z) {
    float acc = 0.0f;
    for (int j = 0; j < a.K; ++j) {
      const int zz = z + j - a.pl;
      if (zz >= 0 && zz < Vz) acc = fmaf(taps[j], col[zz * RW], acc);
    }
    float v = acc;
    if (has_s) {
      const float t = __fmul_rn(acc, s);
      if (t >= 0.0f && t <= 1.0f) mbits |= 1u << (z & 31);
      v = dpc_clip01(t);
    }
    if (mrow && ((z & 31) == 31 || z == Vz - 1)) { mrow[z >> 5] = mbits; mbits = 0; }
    vout[(size_t)z * V * V] = v;
    if (a.mode == DPC_PROJ_MAX) {
      mx = fmaxf(mx, v);
    } else if (a.mode != DPC_PROJ_NONE) {
      const float u = D.clampu ? fminf(fmaxf(v, D.lo), D.hi) : v;
      const float p = (z == 0 ? D.c0 : 1.0f) * u * T;
      T *= (1.0f - u);
      proj += p;
      if (a.probs) a.probs[(size_t)z * plane + ray] = p;
      if (a.depth) dep = fmaf(p, dpc_psi(z, Vz, a.cam_dist), dep);
    }
  }
  if (a.mode == DPC_PROJ_MAX) {
    a.proj[ray] = mx;
  } else if (a.mode != DPC_PROJ_NONE) {
    const float pZ = D.cZ * T;
    a.proj[ray] = proj;
    if (a.probs) a.probs[(size_t)Vz * plane + ray] = pZ;
    if (a.depth) a.depth[ray] = fmaf(pZ, a.max_depth, dep);
  }
}

struct DpcConvZBwdArgs {
  const float* vox; const uint32_t* mask2; const float* scale;
  const float* taps; int K; int pl; int rev;   // taps of the TRANSPOSED correlation (rev: given forward taps)
  int mode; float eps; float cam_dist; float max_depth; int flip_y;
  int B, Vz, V, TY;
  const float* g_proj; const float* g_vox; const float* g_probs; const float* g_depth;
  float* d_in; float* d_scale;
  DpcTapsZ ht; int use_ht;   // fast kernels: host-provided taps in the launch parameters
  // tcgen05 pipeline only: dL/dscale as per-warp partial sums [B * 32 tiles * 8 warps] (folded by the splat backward)
  // instead of atomics, and the small accumulation targets of the splat backward zeroed by CTA 0 -- together they
  // remove the zeroing launch in front of the backward (every stream node costs ~3 us)
  float* d_scale_part;
  DpcZero4Args zero;
};

#ifndef DPC_EMU
__global__ void __launch_bounds__(128)
#else
static void
#endif
dpc_conv_z_bwd_kernel(DpcConvZBwdArgs a) {
  DPC_DYN_SMEM(float, sm);
  __shared__ float red[4];
  const int V = a.V, Vz = a.Vz, TY = a.TY;
  const int b = blockIdx.y, y0 = blockIdx.x * TY;
  const int rows = min(TY, V - y0);
  const int RW = TY * V;
  float* tileU = sm;                         // forward's voxels
  float* tileD = sm + (size_t)Vz * RW;       // T_k, then dL/d(smoothed)
  float* taps = tileD + (size_t)Vz * RW;
  const int tid = threadIdx.x;
  if (tid < a.K) taps[tid] = dpc_tap(a.taps, a.K, tid, a.rev);
  dpc_grid_dep_sync();
  const float* src = a.vox + ((size_t)b * Vz * V + y0) * V;
  const int rowlen = rows * V;
  for (int z = 0; z < Vz; ++z)
    for (int i = tid; i < rowlen; i += blockDim.x) tileU[z * RW + i] = src[(size_t)z * V * V + i];
  __syncthreads();
  float ds = 0.0f;
  const bool active = tid < rowlen;
  if (active) {
    const int ty = tid / V, x = tid - ty * V, y = y0 + ty;
    float* cu = tileU + tid;
    float* cd = tileD + tid;
    const bool has_s = a.scale != nullptr;
    const float s = has_s ? a.scale[b] : 1.0f;
    const DpcDrc D = dpc_drc_consts(a.mode, a.eps);
    const int yo = a.flip_y ? (V - 1 - y) : y;
    const size_t ray = ((size_t)b * V + yo) * V + x;
    const size_t plane = (size_t)a.B * V * V;
    const float gp = a.g_proj ? a.g_proj[ray] : 0.0f;
    const float gd = a.g_depth ? a.g_depth[ray] : 0.0f;
    const float* gv = a.g_vox ? a.g_vox + ((size_t)b * Vz * V + y) * V + x : nullptr;
    const int nwords = (Vz + 31) >> 5;
    const uint32_t* mrow = (a.mask2 && has_s) ? a.mask2 + (((size_t)b * V + y) * V + x) * nwords : nullptr;

    if (a.mode == DPC_PROJ_MAX) {
      float mx = -INFINITY;
      for (int z = 0; z < Vz; ++z) mx = fmaxf(mx, cu[z * RW]);
      int cnt = 0;
      for (int z = 0; z < Vz; ++z) cnt += (cu[z * RW] == mx) ? 1 : 0;
      const float share = gp / (float)cnt;   // TF _MaxGrad: ties share the gradient equally
      for (int z = 0; z < Vz; ++z) cd[z * RW] = (cu[z * RW] == mx) ? share : 0.0f;
    } else if (a.mode == DPC_PROJ_NONE) {
      for (int z = 0; z < Vz; ++z) cd[z * RW] = 0.0f;
    } else {
      // forward sweep: T_k = prod_{j<k} (1-u_j)
      float T = 1.0f;
      for (int z = 0; z < Vz; ++z) {
        const float v = cu[z * RW];
        const float u = D.clampu ? fminf(fmaxf(v, D.lo), D.hi) : v;
        cd[z * RW] = T;
        T *= (1.0f - u);
      }
      // reverse sweep: Q_{k-1} = G_k c_k u_k + (1-u_k) Q_k, Q_{Z-1} = G_Z c_Z;
      // dL/du_k = T_k (G_k c_k - Q_k)
      const float gZ = (a.g_probs ? a.g_probs[(size_t)Vz * plane + ray] : 0.0f) + gd * a.max_depth;
      float Q = gZ * D.cZ;
      for (int z = Vz - 1; z >= 0; --z) {
        const float v = cu[z * RW];
        const float u = D.clampu ? fminf(fmaxf(v, D.lo), D.hi) : v;
        float G = gp + gd * dpc_psi(z, Vz, a.cam_dist);
        if (a.g_probs) G += a.g_probs[(size_t)z * plane + ray];
        const float Gc = G * (z == 0 ? D.c0 : 1.0f);
        float du = cd[z * RW] * (Gc - Q);
        Q = fmaf(1.0f - u, Q, Gc * u);
        if (D.clampu && !(v >= D.lo && v <= D.hi)) du = 0.0f;   // clip_by_value passes lo <= v <= hi
        cd[z * RW] = du;
      }
    }
    // + direct gradient on the voxels output, then back through clip(.*scale) (point_cloud.py:249-253)
    uint32_t mbits = 0;
    for (int z = 0; z < Vz; ++z) {
      float dv = cd[z * RW] + (gv ? gv[(size_t)z * V * V] : 0.0f);
      if (has_s) {
        if (mrow) { if ((z & 31) == 0) mbits = mrow[z >> 5]; }
        else mbits = 0xffffffffu;
        const bool pass = (mbits >> (z & 31)) & 1u;
        dv = pass ? dv : 0.0f;
        // smoothed value = voxels / scale wherever the clip passed (scale != 0)
        ds = fmaf(dv, (s != 0.0f) ? cu[z * RW] / s : 0.0f, ds);
        dv *= s;
      }
      cd[z * RW] = dv;
    }
    // transposed depth correlation (reversed taps) straight to global
    float* dout = a.d_in + ((size_t)b * Vz * V + y) * V + x;
    for (int z = 0; z < Vz; ++z) {
      float acc = 0.0f;
      for (int j = 0; j < a.K; ++j) {
        const int zz = z + j - a.pl;
        if (zz >= 0 && zz < Vz) acc = fmaf(taps[j], cd[zz * RW], acc);
      }
      dout[(size_t)z * V * V] = acc;
    }
  }
  if (a.d_scale) {
    const float v = dpc_warp_sum(ds);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      float t = 0.0f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
      atomicAdd(a.d_scale + b, t);
    }
  }
}

// ------------------------------------------------------------------------------ N1: dL/d(taps), hence dL/d(sigma)
// out[j] += sum over voxels of g[pos] * a[pos + (j - pad_lo) along `axis`]   (a taken as 0 outside the grid):
// the gradient of a zero-padded correlation  out = corr(a, taps)  w.r.t. its taps, given g = dL/d(out).  sigma is a
// pure function of the step in the reference (model_pc.py:35-40) and nothing there consumes dL/dsigma, so this is an
// optional output off the hot path (the composed route of util/point_cloud.py uses it when the taps require a gradient).
// axis: 0 = depth, 1 = y, 2 = x.  One CTA = DPC_TAPCORR_CHUNK consecutive voxels; per tap: thread partial -> warp
// shuffle -> shared memory -> one atomicAdd per CTA and tap.
#define DPC_TAPCORR_THREADS 256
#define DPC_TAPCORR_CHUNK 4096
#ifndef DPC_EMU
__global__ void __launch_bounds__(DPC_TAPCORR_THREADS)
#else
static void
#endif
dpc_tap_corr_kernel(const float* a, const float* g, int axis, long long nvox, int Vz, int V, int K, int pad_lo, float* out) {
  __shared__ float part[DPC_TAPCORR_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long first = (long long)blockIdx.x * DPC_TAPCORR_CHUNK;
  const int len = axis == 0 ? Vz : V;
  const long long stride = axis == 0 ? (long long)V * V : (axis == 1 ? V : 1);
  dpc_grid_dep_wait();
  for (int j = 0; j < K; ++j) {
    const int off = j - pad_lo;
    float s = 0.0f;
    for (int i = tid; i < DPC_TAPCORR_CHUNK; i += DPC_TAPCORR_THREADS) {
      const long long pos = first + i;
      if (pos >= nvox) break;
      const int c = (int)((pos / stride) % len) + off;        // coordinate along the axis of the shifted read
      if (c >= 0 && c < len) s = fmaf(g[pos], a[pos + (long long)off * stride], s);
    }
    s = dpc_warp_sum(s);
    if (lane == 0) part[warp] = s;
    __syncthreads();
    if (tid == 0) {
      float t = 0.0f;
      for (int w = 0; w < DPC_TAPCORR_THREADS / 32; ++w) t += part[w];
      atomicAdd(out + j, t);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ f-3: K3 for the colour grid
// project_volume_rgb_integral (drc.py:126-136): proj_rgb[b,y,x,c] = sum_{z<Vz} p[z,b,y,x] * rgb[b,z,y,x,c] + p[Vz,b,y,x]
// (a white background for the "ray escapes" event).  p [Vz+1,B,V,V] are the (already row-flipped) termination
// probabilities, rgb [B,Vz,V,V,3] the (already row-flipped) colour grid.  One thread per ray; the reference materialises
// p * concat(rgb, ones) as a [Vz+1,B,V,V,3] tensor (three grids) before the reduction.
#ifndef DPC_EMU
__global__ void __launch_bounds__(128)
#else
static void
#endif
dpc_project_rgb_fwd_kernel(const float* p, const float* rgb, int B, int Vz, int V, float* out) {
  const long long rays = (long long)B * V * V;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  dpc_grid_dep_sync();
  if (r >= rays) return;
  const long long b = r / ((long long)V * V), yx = r - b * V * V;
  const float* pr = p + r;                                    // level stride = rays
  const float* cr = rgb + ((b * Vz) * (long long)V * V + yx) * 3;   // level stride = V*V*3
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int z = 0; z < Vz; ++z) {
    const float w = pr[(long long)z * rays];
    const float* c = cr + (long long)z * V * V * 3;
    a0 = fmaf(w, c[0], a0); a1 = fmaf(w, c[1], a1); a2 = fmaf(w, c[2], a2);
  }
  const float bg = pr[(long long)Vz * rays];
  out[r * 3 + 0] = a0 + bg; out[r * 3 + 1] = a1 + bg; out[r * 3 + 2] = a2 + bg;
}

// backward: d_p[z] = sum_c g[c] * rgb[z][c]  (z < Vz),  d_p[Vz] = sum_c g[c];  d_rgb[z][c] = g[c] * p[z]
#ifndef DPC_EMU
__global__ void __launch_bounds__(128)
#else
static void
#endif
dpc_project_rgb_bwd_kernel(const float* p, const float* rgb, const float* g, int B, int Vz, int V, float* d_p, float* d_rgb) {
  const long long rays = (long long)B * V * V;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  dpc_grid_dep_sync();
  if (r >= rays) return;
  const long long b = r / ((long long)V * V), yx = r - b * V * V;
  const float g0 = g[r * 3 + 0], g1 = g[r * 3 + 1], g2 = g[r * 3 + 2];
  const long long coff = ((b * Vz) * (long long)V * V + yx) * 3;
  for (int z = 0; z < Vz; ++z) {
    const long long ci = coff + (long long)z * V * V * 3;
    if (d_p) d_p[r + (long long)z * rays] = g0 * rgb[ci] + g1 * rgb[ci + 1] + g2 * rgb[ci + 2];
    if (d_rgb) {
      const float w = p[r + (long long)z * rays];
      d_rgb[ci] = g0 * w; d_rgb[ci + 1] = g1 * w; d_rgb[ci + 2] = g2 * w;
    }
  }
  if (d_p) d_p[r + (long long)Vz * rays] = g0 + g1 + g2;
}
