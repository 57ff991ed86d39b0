// C-ABI of the projection path (include/dpc_b200.h): argument checks and kernel launches.
// No torch types, no host synchronisation; process-wide state = the experiment knobs only (thread-local hand-overs
// between an entry point and its launchers are consumed within the call).
#include "dpc_common.cuh"
#include "dpc_math.cuh"
#include "dpc_splat.cuh"
#include "dpc_smooth.cuh"
#include "dpc_smooth_fast.cuh"
#include "dpc_smooth_tc.cuh"
#ifdef DPC_EXPERIMENTS
#include "dpc_smooth_fused.cuh"
#include "dpc_fused_bwd.cuh"
#include "dpc_gather_bench.cuh"
#endif
#include "dpc_chamfer.cuh"
#include "dpc_loss.cuh"

static thread_local int g_last_cuda_error = 0;

static int dpc_check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return DPC_ERR_CUDA; }
  return DPC_OK;
}
#define DPC_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { g_last_cuda_error = (int)e__; return DPC_ERR_CUDA; } } while (0)
#define DPC_TRY(call) do { int r__ = (call); if (r__ != DPC_OK) return r__; } while (0)

// Experiment knobs: mutable in the lab build only (benchmark sweeps; not thread-safe); compile-time constants in the
// product build, where only the diagnostics switches below can change.
static int g_diag_stage = 0;     // key 3: CUDA events around every stage of the fused path
static DPC_KNOB_T g_tune[24] = {4, 1, 0, 0, 0, 0, 0, 0, 2, 0, 0, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0};   // [0] points/thread splat fwd, [1] splat bwd
// [10] 1 = the raw grid is zeroed by dpc_zero_kernel and the forward splat runs its transform ahead of the grid
//      dependency (default 0 = cudaMemsetAsync + wait-first splat: the reductions run ~4 us faster behind the driver's
//      memset than behind a store kernel, profiles/r01_k_step_timeline.txt); [11] 1 = 16-byte red.v4 / gathers in the
//      splats; [12] per-kernel timeline (dpc_kt); [13] 1 = keep the zeroing launch + dL/dscale atomics in the fused
//      backward (0 = folded partials, no launch); [14] 1 = the backward splat transforms ahead of its grid dependency;
//      [4] 1 = the gathers of the splat backward run INSIDE the x/y pass of the backward (dpc_fused_bwd.cuh: gather warps
//      next to the pipeline warps, per-sample completion counters; needs p->tr_pc), 0 (default) = the splat backward is its
//      own kernel behind that pass.  Measured (profiles/r02_f_fused_gather.md): correct, but the gathers and the pipeline
//      contend for the SM's load/store path -- 34 + 8 us against 15 + 20 -- so it stays an experiment; [16] its debug
//      flags, [17] 1 = 768-thread variant (10 gather warps);
//      [15] 1 = the fused path smooths x/y IN PLACE and runs the backward in the same grid (two 32 MiB grids per step
//      instead of three)
// measured after the 16-byte gathers (profiles/r01_m_*): forward 4 / 2 / 1 points per thread = 10.2 / 10.7 / 13.9 us,
// backward 21.5 / 28.5 / 20.6 us -> forward 4 (250 CTAs), backward 1 (1000 CTAs)
// SMs the pipelined splat kernels size their grids for (CPU emulation: 4, so that its tests walk several tiles per warp)
static int splat_sm_count() {
#ifndef DPC_EMU
  return dpc_tc_sm_count();
#else
  return 4;
#endif
}
static int tune_ppt(int which) { int v = g_tune[which]; return (v == 1 || v == 2 || v == 4) ? v : 4; }

// ---- optional stage instrumentation of the fused path (dpc_debug_set(3, 1)): CUDA events are
// recorded on the caller's stream around every stage so bench.py can attribute the step time to
// the very kernels the fused entry points launch.  Off by default (no events, no overhead).
#ifndef DPC_EMU
static cudaEvent_t g_ev[8];
static bool g_ev_ready = false;
static void stage_mark(int i, void* stream) {
  if (!g_diag_stage) return;
  if (!g_ev_ready) { for (int k = 0; k < 8; ++k) cudaEventCreate(&g_ev[k]); g_ev_ready = true; }
  cudaEventRecord(g_ev[i], (cudaStream_t)stream);
}
#else
static void stage_mark(int, void*) {}
#endif


static bool shape_ok(int B, int Vz, int V) {
  return B >= 1 && B <= 65535 && V >= 1 && V <= DPC_MAX_V && Vz >= 1 && Vz <= DPC_MAX_V;
}

extern "C" {

int dpc_abi_version(void) { return 3; }
/* ms between the stage marks of the last instrumented forward+backward (synchronises):
 * out[0..5] = splat_fwd(+memset), conv_xy_fwd, conv_z_fwd, conv_z_bwd(+memsets), conv_xy_bwd, splat_bwd */
int dpc_debug_stage_ms(float* out6) {
#ifndef DPC_EMU
  if (!out6 || !g_ev_ready) return DPC_ERR_ARG;
  if (cudaEventSynchronize(g_ev[7]) != cudaSuccess) return DPC_ERR_CUDA;
  const int pairs[6][2] = {{0, 1}, {1, 2}, {2, 3}, {4, 5}, {5, 6}, {6, 7}};
  for (int i = 0; i < 6; ++i)
    if (cudaEventElapsedTime(&out6[i], g_ev[pairs[i][0]], g_ev[pairs[i][1]]) != cudaSuccess) return DPC_ERR_CUDA;
  return DPC_OK;
#else
  (void)out6;
  return DPC_ERR_ARG;
#endif
}

int dpc_debug_set(int key, int value) {
  if (key < 0 || key >= 24) return DPC_ERR_ARG;
  if (key == 3) { g_diag_stage = value; return DPC_OK; }
#ifndef DPC_EMU
  if (key == 8) { dpc_tc_enable = value; return DPC_OK; }      // kernel family of the 64^3 smoothing passes (tests)
  if (key == 12) {     // (re)arm the per-kernel timeline: minima to ~0, maxima to 0
    unsigned long long init[64];
    for (int i = 0; i < 64; ++i) init[i] = ((i & 3) < 2) ? ~0ull : 0ull;
    cudaMemcpyToSymbol(dpc_kt, init, sizeof(init));
    int v = value; cudaMemcpyToSymbol(dpc_kt_on, &v, sizeof(int));
    return DPC_OK;
  }
#endif
#ifdef DPC_EXPERIMENTS
  g_tune[key] = value;
  if (key == 5) dpc_ignore_host_taps = value ? 1 : 0;
  if (key == 6) dpc_z_tile_cpasync = value ? 1 : 0;
  if (key == 7) dpc_xy_dbg = value;
#ifndef DPC_EMU
  if (key == 2) { dpc_tcp_pdrain = (value == 1) ? 1 : 0; dpc_tcp_ns3 = (value == 2) ? 1 : 0; }
  if (key == 9) { int v = value; cudaMemcpyToSymbol(dpc_tcp_trace_on, &v, sizeof(int)); }
#endif
  return DPC_OK;
#else
  (void)value;
  return DPC_ERR_ARG;          // an experiment knob: only the lab build (-DDPC_EXPERIMENTS) has it
#endif
}
int dpc_is_lab_build(void) {
#ifdef DPC_EXPERIMENTS
  return 1;
#else
  return 0;
#endif
}
int dpc_last_cuda_error(void) { return g_last_cuda_error; }
#ifdef DPC_EXPERIMENTS
/* diagnostics: copy the 16 x 16 pipeline trace of CTA 0 (clock64 values) to host memory (synchronises) */
int dpc_debug_trace_read(long long* host_out) {
#ifndef DPC_EMU
  if (!host_out) return DPC_ERR_NULL;
  if (cudaMemcpyFromSymbol(host_out, dpc_tcp_trace, sizeof(long long) * 256) != cudaSuccess) return DPC_ERR_CUDA;
  return cudaMemcpyFromSymbol(host_out + 256, dpc_tcp_cta_ns, sizeof(long long) * 480) == cudaSuccess ? DPC_OK : DPC_ERR_CUDA;
#else
  return DPC_ERR_ARG;
#endif
}
#endif  // DPC_EXPERIMENTS
/* diagnostics: the per-kernel timeline (16 kernels x {first entry, first/last CTA past its dependency, last exit}, ns) */
int dpc_debug_ktrace_read(unsigned long long* host_out) {
#ifndef DPC_EMU
  if (!host_out) return DPC_ERR_NULL;
  return cudaMemcpyFromSymbol(host_out, dpc_kt, sizeof(unsigned long long) * 64) == cudaSuccess ? DPC_OK : DPC_ERR_CUDA;
#else
  (void)host_out;
  return DPC_ERR_ARG;
#endif
}
#ifdef DPC_EXPERIMENTS
/* diagnostics: per-CTA phase stamps of the splat kernels (2 kernels x 512 CTAs x 8 slots, ns) */
int dpc_debug_phase_read(unsigned long long* host_out) {
#ifndef DPC_EMU
  if (!host_out) return DPC_ERR_NULL;
  return cudaMemcpyFromSymbol(host_out, dpc_ph, sizeof(unsigned long long) * 2 * 512 * 8) == cudaSuccess ? DPC_OK : DPC_ERR_CUDA;
#else
  (void)host_out;
  return DPC_ERR_ARG;
#endif
}
/* diagnostics: tcgen05.mma micro-benchmark (see dpc_tc_mma_bench_kernel); out = 3 int64 per CTA (device memory) */
int dpc_debug_mma_bench(long long* out, int nctas, int threads, int reps, int nmma, int spin, int M, int N, void* stream) {
#ifndef DPC_EMU
  if (!out || nctas < 1 || threads < 32 || threads > 512 || nmma < 1) return DPC_ERR_ARG;
  if (cudaFuncSetAttribute(dpc_tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608) != cudaSuccess) return DPC_ERR_CUDA;
  dpc_tc_mma_bench_kernel<<<nctas, threads, 196608, (cudaStream_t)stream>>>(out, reps, nmma, spin, M, N);
  return dpc_check_launch();
#else
  return DPC_ERR_ARG;
#endif
}
/* diagnostics: the splat backward's gather pattern alone (dpc_gather_bench.cuh); smem_pad bytes of dynamic shared memory
 * per CTA limit the resident CTAs per SM */
int dpc_debug_gather_bench(const float* grid, float* out, int B, int N, int V, int variant, int threads, int ppt,
                           int share_log2, int smem_pad, int cg, void* stream) {
#ifndef DPC_EMU
  if (!grid || !out || B < 1 || N < 1 || V < 8 || (V & 3) || threads < 32 || threads > 1024 || ppt < 1 || smem_pad < 0 ||
      smem_pad > 200 * 1024 || share_log2 < 0 || share_log2 > 5)
    return DPC_ERR_ARG;
  DpcGatherBenchArgs a{grid, out, N, V, ppt, share_log2, 0x1234567u, 0u, cg};
  const dim3 g((unsigned)((N + threads * ppt - 1) / (threads * ppt)), (unsigned)B);
  void (*k)(DpcGatherBenchArgs) = nullptr;
  switch (variant) {
    case 0: k = dpc_gather_bench_kernel<0>; break;
    case 1: k = dpc_gather_bench_kernel<1>; break;
    case 2: k = dpc_gather_bench_kernel<2>; break;
    case 3: k = dpc_gather_bench_kernel<3>; break;
    case 4: k = dpc_gather_bench_kernel<4>; break;
    case 5: k = dpc_gather_bench_kernel<5>; break;
    case 6: k = dpc_gather_bench_kernel<6>; break;
    case 7: if (threads > 256) return DPC_ERR_ARG; k = dpc_gather_bench_kernel<7>; break;
    default: return DPC_ERR_ARG;
  }
  if (smem_pad > 48 * 1024 &&
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pad) != cudaSuccess) return DPC_ERR_CUDA;
  k<<<g, threads, smem_pad, (cudaStream_t)stream>>>(a);
  return dpc_check_launch();
#else
  return DPC_ERR_ARG;
#endif
}
#endif  // DPC_EXPERIMENTS
int dpc_is_cuda_build(void) {
#ifdef DPC_EMU
  return 0;
#else
  return 1;
#endif
}
const char* dpc_error_string(int code) {
  switch (code) {
    case DPC_OK: return "ok";
    case DPC_ERR_NULL: return "required pointer is NULL";
    case DPC_ERR_SHAPE: return "unsupported shape (B, N, V, Vz or tap count out of range)";
    case DPC_ERR_ARG: return "inconsistent arguments";
    case DPC_ERR_CUDA: return "CUDA error (see dpc_last_cuda_error)";
    case DPC_ERR_WORKSPACE: return "workspace too small or misaligned";
    default: return "unknown error";
  }
}

}  // extern "C"

static int splat_fwd_launch(const float* pc, const float* pose, int pose_kind, const float* trans,
                            const float* focal, float focal_const, float cam_dist, const float* rgb,
                            int B, int N, int Vz, int V,
                            float* tr_pc, float* vox, float* vox_rgb, int32_t* idx_out, uint8_t* valid_out,
                            const int32_t* sel, int N_src, void* stream, int early = 0, unsigned* zero_u32 = nullptr, int n_zero = 0,
                            bool* zeroes_grid = nullptr) {
  // zeroes_grid != NULL: the caller has NOT zeroed `vox`; *zeroes_grid comes back true when the splat kernel does it
  // itself (dpc_splat_fwd_warp_kernel<.., true>, cooperative launch), false when the caller has to (then nothing was launched)
  // early (lab build): the stream predecessor is a grid-zeroing KERNEL; the splat then transforms ahead of its dependency
  if (!pc) return DPC_ERR_NULL;
  if (sel && (N_src < N || rgb)) return DPC_ERR_ARG;
  if (pose_kind != DPC_POSE_NONE && !pose) return DPC_ERR_NULL;
  if (pose_kind < DPC_POSE_NONE || pose_kind > DPC_POSE_MATRIX) return DPC_ERR_ARG;
  if (trans && pose_kind != DPC_POSE_QUAT) return DPC_ERR_ARG;  // reference: tf.slice rank error
  if (rgb && !vox_rgb && vox) return DPC_ERR_NULL;
  if (!shape_ok(B, Vz, V) || N < 1) return DPC_ERR_SHAPE;
  DpcSplatArgs a;
  a.pc = pc; a.pose = pose; a.trans = trans; a.focal = (pose_kind == DPC_POSE_QUAT) ? focal : nullptr; a.rgb = rgb;
  a.pose_kind = pose_kind; a.focal_const = focal_const; a.cam_dist = cam_dist;
  a.B = B; a.N = N; a.Vz = Vz; a.V = V;
  a.tr_pc = tr_pc; a.vox = vox; a.vox_rgb = vox_rgb; a.idx_out = idx_out; a.valid_out = valid_out;
  a.early = sel ? 0 : early;
  a.red4 = g_tune[11] ? 1 : 0;
  a.sel = sel; a.N_src = N_src;
  a.zero_u32 = zero_u32; a.n_zero = n_zero;
  // software-pipelined form (dpc_splat_fwd_warp_kernel), same grid sizing as the backward's; lab build: knob 0 = 1 / 2 /
  // 8 selects the tile-per-CTA kernel with 1 / 2 / 4 points per thread, knob 19 = warps per SM the grid is sized for
  if (g_tune[0] == 4 && !rgb && !zero_u32) {
    const int tiles = (N + 31) / 32;
    const int per_sm = g_tune[19] > 0 ? g_tune[19] : 28;
    const long long cap = (long long)splat_sm_count() * per_sm;
    int k = (int)(((long long)B * tiles + cap - 1) / cap);
    int wps = (tiles + k - 1) / k;
    dim3 g((wps + DPC_SPLAT_WPC - 1) / DPC_SPLAT_WPC, B);
    const dim3 blk(32 * DPC_SPLAT_WPC);
    if (zeroes_grid) {
      // in-kernel zeroing needs every CTA resident (grid-wide barrier): cooperative launch, grid shrunk until it fits.
      // Lab build only (knob 10 = 3): measured slower than the driver's memset in front of the kernel (zeros + barrier
      // take 10 us inside the kernel, the memset 5 us: step 98.5 vs 96.4 us; profiles/r02_w_timeline_coop_zero.txt)
      *zeroes_grid = false;
#if defined(DPC_EXPERIMENTS) && !defined(DPC_EMU)
      void (*kz)(DpcSplatArgs) = dpc_splat_fwd_warp_kernel<7, true, false>;
      int per = 0;
      if (g_tune[10] == 3 && !sel && vox && (V & 3) == 0 && ((((uintptr_t)vox) & 15u) == 0) &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kz, 32 * DPC_SPLAT_WPC, 0) == cudaSuccess && per > 0) {
        const long long fit = (long long)per * dpc_tc_sm_count();
        while ((long long)g.x * B > fit && wps > 1) { ++k; wps = (tiles + k - 1) / k; g.x = (wps + DPC_SPLAT_WPC - 1) / DPC_SPLAT_WPC; }
        if ((long long)g.x * B <= fit) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = g; cfg.blockDim = blk; cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeCooperative;
          attr[0].val.cooperative = 1;
          cfg.attrs = attr; cfg.numAttrs = 1;
          a.early = 0;
          if (cudaLaunchKernelEx(&cfg, kz, a) == cudaSuccess) { *zeroes_grid = true; return dpc_check_launch(); }
          cudaGetLastError();      // not launchable cooperatively here: the caller zeroes, the plain kernel follows
        }
      }
#endif
      return DPC_OK;
    }
    if (sel) { DPC_LAUNCH((dpc_splat_fwd_warp_kernel<7, false, true>), g, blk, 0, stream, a); }
    else if (per_sm > 24) { DPC_LAUNCH((dpc_splat_fwd_warp_kernel<7, false, false>), g, blk, 0, stream, a); }
    else { DPC_LAUNCH((dpc_splat_fwd_warp_kernel<6, false, false>), g, blk, 0, stream, a); }
    return dpc_check_launch();
  }
  if (zeroes_grid) { *zeroes_grid = false; return DPC_OK; }
  const int ppt = tune_ppt(0), tile = DPC_SPLAT_THREADS * ppt;
  dim3 grid((N + tile - 1) / tile, B);
  if (ppt == 4) { DPC_LAUNCH(dpc_splat_fwd_kernel<4>, grid, dim3(DPC_SPLAT_THREADS), 0, stream, a); }
  else if (ppt == 2) { DPC_LAUNCH(dpc_splat_fwd_kernel<2>, grid, dim3(DPC_SPLAT_THREADS), 0, stream, a); }
  else { DPC_LAUNCH(dpc_splat_fwd_kernel<1>, grid, dim3(DPC_SPLAT_THREADS), 0, stream, a); }
  return dpc_check_launch();
}

extern "C" int dpc_splat_fwd(const float* pc, const float* pose, int pose_kind, const float* trans,
                             const float* focal, float focal_const, float cam_dist, const float* rgb,
                             int B, int N, int Vz, int V,
                             float* tr_pc, float* vox, float* vox_rgb, int32_t* idx_out, uint8_t* valid_out,
                             void* stream) {
  return splat_fwd_launch(pc, pose, pose_kind, trans, focal, focal_const, cam_dist, rgb, B, N, Vz, V, tr_pc, vox, vox_rgb,
                          idx_out, valid_out, nullptr, 0, stream);
}

// the public entry point plus the fused backward's extra: fold the depth pass's dL/dscale partials
static int splat_bwd_launch(const float* pc, const float* pose, int pose_kind, const float* trans,
                            const float* focal, float focal_const, float cam_dist, const float* rgb,
                            int rgb_stop_grad, int B, int N, int Vz, int V,
                            const float* d_vox, const float* d_vox_rgb, const float* d_tr_pc_in,
                            float* d_pc, float* d_pose, float* d_trans, float* d_focal, float* d_rgb,
                            const float* d_scale_part, int n_part, float* d_scale_out, void* stream,
                            const int32_t* sel = nullptr, int N_src = 0, const float* tr_pc = nullptr) {
  if (!pc) return DPC_ERR_NULL;
  if (sel && (N_src < N || rgb)) return DPC_ERR_ARG;
  if (pose_kind != DPC_POSE_NONE && !pose) return DPC_ERR_NULL;
  if (pose_kind < DPC_POSE_NONE || pose_kind > DPC_POSE_MATRIX) return DPC_ERR_ARG;
  if (trans && pose_kind != DPC_POSE_QUAT) return DPC_ERR_ARG;
  if (d_vox_rgb && !rgb) return DPC_ERR_NULL;
  if (!shape_ok(B, Vz, V) || N < 1) return DPC_ERR_SHAPE;
  DpcSplatBwdArgs a;
  a.pc = pc; a.pose = pose; a.trans = trans; a.focal = (pose_kind == DPC_POSE_QUAT) ? focal : nullptr; a.rgb = rgb;
  a.pose_kind = pose_kind; a.focal_const = focal_const; a.cam_dist = cam_dist; a.rgb_stop_grad = rgb_stop_grad;
  a.B = B; a.N = N; a.Vz = Vz; a.V = V;
  a.d_vox = d_vox; a.d_vox_rgb = d_vox_rgb; a.d_tr_pc_in = d_tr_pc_in;
  a.d_pc = d_pc; a.d_pose = d_pose; a.d_trans = d_trans; a.d_focal = d_focal; a.d_rgb = d_rgb;
  a.early = g_tune[14] ? 1 : 0;
  a.gather4 = g_tune[11] ? 1 : 0;
  a.gather_cg = g_tune[22] ? 1 : 0;
  a.stagger_ns = g_tune[23];
  a.d_scale_part = d_scale_part; a.n_part = n_part; a.d_scale_out = d_scale_out;
  a.sel = sel; a.N_src = N_src; a.tr_pc = tr_pc;
  if (sel && d_pc) DPC_CUDA(cudaMemsetAsync(d_pc, 0, (size_t)B * N_src * 12, (cudaStream_t)stream));      // dropped points: zero gradient
  const int ppt = tune_ppt(1), tile = DPC_SPLAT_THREADS * ppt;
  dim3 grid((N + tile - 1) / tile, B);
  // Default whenever the forward's tr_pc is at hand (the fused path passes it): the software-pipelined form,
  // dpc_splat_bwd_warp_kernel -- independent warps, every warp resident at once, k tiles of 32 points per warp with k the
  // smallest count for which the grid fits 28 warps per SM (13.6-14.0 us against 20.0 us of the tile-per-CTA kernel at
  // B=32, N=8000; 24 warps per SM, i.e. k = 3: 15.1 us; profiles/r02_w_splat_bwd_warp.md).  Lab build: knob 20 = 6
  // forces the tile-per-CTA kernel, knob 19 = warps per SM the grid is sized for.
  if ((g_tune[20] == 0 || g_tune[20] == 4) && tr_pc && d_vox && !rgb && !d_vox_rgb && !d_rgb && (V & 3) == 0 &&
      ((((uintptr_t)d_vox) & 15u) == 0)) {
    const int tiles = (N + 31) / 32;
    const int per_sm = g_tune[19] > 0 ? g_tune[19] : 28;
    const long long cap = (long long)splat_sm_count() * per_sm;
    const int k = (int)(((long long)B * tiles + cap - 1) / cap);
    const int wps = (tiles + k - 1) / k;
    void (*kw)(DpcSplatBwdArgs) = sel ? dpc_splat_bwd_warp_kernel<7, true>
                                  : (per_sm > 24 ? dpc_splat_bwd_warp_kernel<7, false> : dpc_splat_bwd_warp_kernel<6, false>);
#ifndef DPC_EMU
    // 28 x 7 KB of static shared memory per SM need the large carve-out (per device, so set on every call like the
    // dynamic-shared-memory limits of the other launchers; a host-side attribute, legal during stream capture)
    if (cudaFuncSetAttribute(kw, cudaFuncAttributePreferredSharedMemoryCarveout, 100) != cudaSuccess) return DPC_ERR_CUDA;
#endif
    DPC_LAUNCH(kw, dim3((wps + DPC_SPLAT_WPC - 1) / DPC_SPLAT_WPC, B), dim3(32 * DPC_SPLAT_WPC), 0, stream, a);
    return dpc_check_launch();
  }
#if !defined(DPC_EMU) && defined(DPC_EXPERIMENTS)
  if (ppt == 1 && g_tune[20] != 6 && (g_tune[18] || g_tune[19] || g_tune[20])) {
    // occupancy / decorrelation experiments (knobs 18: 128-thread CTAs, 19: compiled for 75 % occupancy, 20: independent gathers)
    const int sel3 = (g_tune[18] ? 4 : 0) | (g_tune[19] ? 2 : 0) | (g_tune[20] ? 1 : 0);
    dim3 g128((N + 127) / 128, B);
    if (g_tune[20] == 3) {        // predicated independent gathers (dpc_gather_corners_pred)
      if (g_tune[18]) { DPC_LAUNCH((dpc_splat_bwd_kernel<1, 128, 0, 3>), g128, dim3(128), 0, stream, a); }
      else { DPC_LAUNCH((dpc_splat_bwd_kernel<1, 256, 0, 3>), grid, dim3(256), 0, stream, a); }
      return dpc_check_launch();
    }
    if (g_tune[20] == 2) {        // single guarded gather path
      if (g_tune[18]) { DPC_LAUNCH((dpc_splat_bwd_kernel<1, 128, 0, 2>), g128, dim3(128), 0, stream, a); }
      else { DPC_LAUNCH((dpc_splat_bwd_kernel<1, 256, 0, 2>), grid, dim3(256), 0, stream, a); }
      return dpc_check_launch();
    }
    switch (sel3) {
      case 1: DPC_LAUNCH((dpc_splat_bwd_kernel<1, 256, 0, 1>), grid, dim3(256), 0, stream, a); break;
      case 2: DPC_LAUNCH((dpc_splat_bwd_kernel<1, 256, 6, 0>), grid, dim3(256), 0, stream, a); break;
      case 3: DPC_LAUNCH((dpc_splat_bwd_kernel<1, 256, 6, 1>), grid, dim3(256), 0, stream, a); break;
      case 4: DPC_LAUNCH((dpc_splat_bwd_kernel<1, 128, 0, 0>), g128, dim3(128), 0, stream, a); break;
      case 5: DPC_LAUNCH((dpc_splat_bwd_kernel<1, 128, 0, 1>), g128, dim3(128), 0, stream, a); break;
      case 6: DPC_LAUNCH((dpc_splat_bwd_kernel<1, 128, 12, 0>), g128, dim3(128), 0, stream, a); break;
      default: DPC_LAUNCH((dpc_splat_bwd_kernel<1, 128, 12, 1>), g128, dim3(128), 0, stream, a); break;
    }
    return dpc_check_launch();
  }
#endif
  if (ppt == 4) { DPC_LAUNCH((dpc_splat_bwd_kernel<4, DPC_SPLAT_THREADS>), grid, dim3(DPC_SPLAT_THREADS), 0, stream, a); }
  else if (ppt == 2) { DPC_LAUNCH((dpc_splat_bwd_kernel<2, DPC_SPLAT_THREADS>), grid, dim3(DPC_SPLAT_THREADS), 0, stream, a); }
  else {
    // one point per thread in 128-thread CTAs: 20.2 us against 21.1 us with 256-thread CTAs at B=32, N=8000
    // (profiles/r02_j_splat_bwd_occupancy.md; compiling for more resident CTAs or issuing the gathers independently is slower)
    dim3 g128((N + 127) / 128, B);
    DPC_LAUNCH((dpc_splat_bwd_kernel<1, 128>), g128, dim3(128), 0, stream, a);
  }
  return dpc_check_launch();
}

extern "C" int dpc_splat_bwd(const float* pc, const float* pose, int pose_kind, const float* trans,
                             const float* focal, float focal_const, float cam_dist, const float* rgb,
                             int rgb_stop_grad, int B, int N, int Vz, int V,
                             const float* d_vox, const float* d_vox_rgb, const float* d_tr_pc_in,
                             float* d_pc, float* d_pose, float* d_trans, float* d_focal, float* d_rgb,
                             void* stream) {
  return splat_bwd_launch(pc, pose, pose_kind, trans, focal, focal_const, cam_dist, rgb, rgb_stop_grad, B, N, Vz, V,
                          d_vox, d_vox_rgb, d_tr_pc_in, d_pc, d_pose, d_trans, d_focal, d_rgb, nullptr, 0, nullptr, stream);
}

// (internal launchers below have C++ linkage)

// ---- internal launchers: like the public entry points plus `rev` (read the taps back to front:
// the transposed correlation, so the backward needs no reversed copy of the taps), NULL taps =
// identity filter, and `zero_in` (conv_xy hands its input back all-zero once it has been read).
static int launch_conv_xy(const float* in, float* out, const float* taps_x, int Kx, int pad_lo_x,
                          const float* taps_y, int Ky, int pad_lo_y, int B, int Vz, int V,
                          int clip_in, uint32_t* mask_bits_out, const uint32_t* mask_bits_in,
                          int rev, int zero_in, void* stream, const float* hx = nullptr, const float* hy = nullptr) {
  if (!in || !out) return DPC_ERR_NULL;
  if (!shape_ok(B, Vz, V) || Kx < 1 || Ky < 1 || Kx > DPC_MAX_TAPS || Ky > DPC_MAX_TAPS) return DPC_ERR_SHAPE;
  if (pad_lo_x < 0 || pad_lo_x >= Kx || pad_lo_y < 0 || pad_lo_y >= Ky) return DPC_ERR_ARG;
  if ((mask_bits_out || mask_bits_in) && ((V * V) % 32 != 0)) return DPC_ERR_SHAPE;
  if ((int64_t)B * Vz > 2147483647LL) return DPC_ERR_SHAPE;
  float* zero_ptr = zero_in ? const_cast<float*>(in) : nullptr;
#ifndef DPC_EMU
  // tensor-core path (dpc_smooth_tc.cuh): 64^3 grids, any tap count, same taps along x and y
  if (dpc_tc_conv_xy_supported(V, Kx, pad_lo_x, Ky, pad_lo_y, taps_x, taps_y, (int64_t)B * Vz, zero_ptr)) {
    dpc_tcp_host_taps_next = hx;
    DPC_TRY(dpc_tc_conv_xy_launch(in, out, taps_x, Kx, pad_lo_x, (int64_t)B * Vz, clip_in, mask_bits_out, mask_bits_in, rev, stream));
    return dpc_check_launch();
  }
#endif
  if (taps_x && taps_y && dpc_conv_xy_fast_supported(V, Kx, pad_lo_x, Ky, pad_lo_y) && ((int64_t)B * Vz * V * V) % (V == 128 ? 16384 : 4096) == 0) {
    DPC_TRY(dpc_conv_xy_fast_launch(in, out, taps_x, taps_y, Kx, B, Vz, V, clip_in, mask_bits_out, mask_bits_in,
                                    rev, zero_ptr, hx, hy, stream));
    return dpc_check_launch();
  }
  DpcConvXYArgs a;
  a.in = in; a.out = out; a.taps_x = taps_x; a.Kx = Kx; a.plx = pad_lo_x;
  a.taps_y = taps_y; a.Ky = Ky; a.ply = pad_lo_y; a.B = B; a.Vz = Vz; a.V = V; a.clip_in = clip_in;
  a.mask_out = mask_bits_out; a.mask_in = mask_bits_in; a.rev = rev; a.zero_ptr = zero_ptr;
  const size_t smem = (size_t)(2 * V * V + 2 * (DPC_MAX_TAPS + 1)) * sizeof(float);
#ifndef DPC_EMU
  if (smem > 48 * 1024) DPC_CUDA(cudaFuncSetAttribute(dpc_conv_xy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
  DPC_LAUNCH(dpc_conv_xy_kernel, dim3(B * Vz), dim3(DPC_CONV_THREADS), smem, stream, a);
  return dpc_check_launch();
}

static int conv_z_ty(int V) {
  int ty = 128 / V;
  return ty < 1 ? 1 : ty;
}

static int launch_conv_z_fwd(const float* in, const float* taps_z, int Kz, int pad_lo_z,
                             const float* scale, int mode, float clip_eps, float cam_dist, float max_depth,
                             int flip_y, int B, int Vz, int V,
                             float* vox_out, uint32_t* mask2_out, float* proj, float* drc_probs,
                             float* proj_depth, void* stream, const float* hz = nullptr) {
  if (!in || !vox_out) return DPC_ERR_NULL;
  if (mode < DPC_PROJ_NONE || mode > DPC_PROJ_DRC_PROD) return DPC_ERR_ARG;
  if (mode != DPC_PROJ_NONE && !proj) return DPC_ERR_NULL;
  if ((drc_probs || proj_depth) && (mode == DPC_PROJ_NONE || mode == DPC_PROJ_MAX)) return DPC_ERR_ARG;
  if (!shape_ok(B, Vz, V) || Kz < 1 || Kz > DPC_MAX_TAPS) return DPC_ERR_SHAPE;
  if (pad_lo_z < 0 || pad_lo_z >= Kz) return DPC_ERR_ARG;
#ifndef DPC_EMU
  if (dpc_tc_conv_z_supported(V, Vz, Kz, drc_probs != nullptr || proj_depth != nullptr)) {
    DpcConvZArgs a;
    dpc_set_taps_z(&a.ht, &a.use_ht, nullptr, 0, 0);
    a.in = in; a.taps = taps_z; a.K = Kz; a.pl = pad_lo_z; a.rev = 0; a.scale = scale; a.mode = mode; a.eps = clip_eps;
    a.cam_dist = cam_dist; a.max_depth = max_depth; a.flip_y = flip_y; a.B = B; a.Vz = Vz; a.V = V; a.TY = 2;
    a.vox_out = vox_out; a.mask2_out = mask2_out; a.proj = proj; a.probs = nullptr; a.depth = nullptr;
    dpc_tcp_host_taps_next = hz;
    DPC_TRY(dpc_tc_conv_z_fwd_launch(a, stream));
    return dpc_check_launch();
  }
#endif
  if (taps_z && dpc_conv_z_fast_supported(V, Vz, Kz, pad_lo_z, drc_probs != nullptr || proj_depth != nullptr)) {
    DPC_TRY(dpc_conv_z_fwd_fast_launch(in, taps_z, Kz, scale, mode, clip_eps, cam_dist, max_depth, flip_y, B, Vz, V,
                                       vox_out, mask2_out, proj, drc_probs, proj_depth, hz, stream));
    return dpc_check_launch();
  }
  DpcConvZArgs a;
  dpc_set_taps_z(&a.ht, &a.use_ht, nullptr, 0, 0);
  a.in = in; a.taps = taps_z; a.K = Kz; a.pl = pad_lo_z; a.rev = 0; a.scale = scale; a.mode = mode; a.eps = clip_eps;
  a.cam_dist = cam_dist; a.max_depth = max_depth; a.flip_y = flip_y; a.B = B; a.Vz = Vz; a.V = V; a.TY = conv_z_ty(V);
  a.vox_out = vox_out; a.mask2_out = mask2_out; a.proj = proj; a.probs = drc_probs; a.depth = proj_depth;
  const size_t smem = ((size_t)Vz * a.TY * V + DPC_MAX_TAPS + 1) * sizeof(float);
#ifndef DPC_EMU
  if (smem > 48 * 1024) DPC_CUDA(cudaFuncSetAttribute(dpc_conv_z_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
  dim3 grid((V + a.TY - 1) / a.TY, B);
  DPC_LAUNCH(dpc_conv_z_fwd_kernel, grid, dim3(128), smem, stream, a);
  return dpc_check_launch();
}

static int launch_conv_z_bwd(const float* vox, const uint32_t* mask2, const float* scale,
                             const float* taps, int Kz, int pad_lo, int rev,
                             int mode, float clip_eps, float cam_dist, float max_depth, int flip_y,
                             int B, int Vz, int V,
                             const float* g_proj, const float* g_vox, const float* g_probs, const float* g_depth,
                             float* d_in, float* d_scale, void* stream, const float* hz = nullptr,
                             float* d_scale_part = nullptr, const DpcZero4Args* zero = nullptr) {
  if (!vox || !d_in) return DPC_ERR_NULL;
  if (mode < DPC_PROJ_NONE || mode > DPC_PROJ_DRC_PROD) return DPC_ERR_ARG;
  if ((g_probs || g_depth) && (mode == DPC_PROJ_NONE || mode == DPC_PROJ_MAX)) return DPC_ERR_ARG;
  if (!shape_ok(B, Vz, V) || Kz < 1 || Kz > DPC_MAX_TAPS) return DPC_ERR_SHAPE;
  if (pad_lo < 0 || pad_lo >= Kz) return DPC_ERR_ARG;
  const bool lean_case = (mode == DPC_PROJ_DRC) && scale && mask2 && g_proj && !g_vox;
#ifndef DPC_EMU
  // the persistent tcgen05 pipeline also has the max-projection form of the silhouette-only backward
  const bool lean_tcp = (mode == DPC_PROJ_DRC || mode == DPC_PROJ_MAX) && scale && mask2 && g_proj && !g_vox && dpc_tc_level() == 2;
  if ((lean_case || lean_tcp) && dpc_tc_conv_z_supported(V, Vz, Kz, g_probs != nullptr || g_depth != nullptr)) {
    DpcConvZBwdArgs a = {};
    dpc_set_taps_z(&a.ht, &a.use_ht, nullptr, 0, 0);
    a.vox = vox; a.mask2 = mask2; a.scale = scale; a.taps = taps; a.K = Kz; a.pl = pad_lo; a.rev = rev;
    a.mode = mode; a.eps = clip_eps; a.cam_dist = cam_dist; a.max_depth = max_depth; a.flip_y = flip_y;
    a.B = B; a.Vz = Vz; a.V = V; a.TY = 2;
    a.g_proj = g_proj; a.g_vox = nullptr; a.g_probs = nullptr; a.g_depth = nullptr; a.d_in = d_in; a.d_scale = d_scale;
    if (d_scale_part && dpc_tc_level() == 2) { a.d_scale_part = d_scale_part; if (zero) a.zero = *zero; }
    dpc_tcp_host_taps_next = hz;
    DPC_TRY(dpc_tc_conv_z_bwd_lean_launch(a, stream));
    return dpc_check_launch();
  }
#endif
  if (taps && dpc_conv_z_fast_supported(V, Vz, Kz, pad_lo, g_probs != nullptr || g_depth != nullptr) &&
      (lean_case || dpc_conv_z_bwd_fast_general_ok(V))) {
    DPC_TRY(dpc_conv_z_bwd_fast_launch(vox, mask2, scale, taps, Kz, mode, clip_eps, cam_dist, max_depth, flip_y,
                                       B, Vz, V, g_proj, g_vox, g_probs, g_depth, d_in, d_scale, rev, hz, stream));
    return dpc_check_launch();
  }
  DpcConvZBwdArgs a = {};
  dpc_set_taps_z(&a.ht, &a.use_ht, nullptr, 0, 0);
  a.vox = vox; a.mask2 = mask2; a.scale = scale; a.taps = taps; a.K = Kz; a.pl = pad_lo; a.rev = rev;
  a.mode = mode; a.eps = clip_eps; a.cam_dist = cam_dist; a.max_depth = max_depth; a.flip_y = flip_y;
  a.B = B; a.Vz = Vz; a.V = V; a.TY = conv_z_ty(V);
  a.g_proj = g_proj; a.g_vox = g_vox; a.g_probs = g_probs; a.g_depth = g_depth; a.d_in = d_in; a.d_scale = d_scale;
  const size_t smem = ((size_t)2 * Vz * a.TY * V + DPC_MAX_TAPS + 1) * sizeof(float);
#ifndef DPC_EMU
  if (smem > 48 * 1024) DPC_CUDA(cudaFuncSetAttribute(dpc_conv_z_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
  dim3 grid((V + a.TY - 1) / a.TY, B);
  DPC_LAUNCH(dpc_conv_z_bwd_kernel, grid, dim3(128), smem, stream, a);
  return dpc_check_launch();
}

extern "C" {

int dpc_conv_xy(const float* in, float* out, const float* taps_x, int Kx, int pad_lo_x,
                const float* taps_y, int Ky, int pad_lo_y, int B, int Vz, int V,
                int clip_in, uint32_t* mask_bits_out, const uint32_t* mask_bits_in, void* stream) {
  if (!taps_x || !taps_y) return DPC_ERR_NULL;
  return launch_conv_xy(in, out, taps_x, Kx, pad_lo_x, taps_y, Ky, pad_lo_y, B, Vz, V, clip_in, mask_bits_out,
                        mask_bits_in, 0, 0, stream);
}

int dpc_conv_z_fwd(const float* in, const float* taps_z, int Kz, int pad_lo_z,
                   const float* scale, int mode, float clip_eps, float cam_dist, float max_depth,
                   int flip_y, int B, int Vz, int V,
                   float* vox_out, uint32_t* mask2_out, float* proj, float* drc_probs,
                   float* proj_depth, void* stream) {
  if (!taps_z) return DPC_ERR_NULL;
  return launch_conv_z_fwd(in, taps_z, Kz, pad_lo_z, scale, mode, clip_eps, cam_dist, max_depth, flip_y, B, Vz, V,
                           vox_out, mask2_out, proj, drc_probs, proj_depth, stream);
}

int dpc_conv_z_bwd(const float* vox, const uint32_t* mask2, const float* scale,
                   const float* taps_z_rev, int Kz, int pad_lo_z_rev,
                   int mode, float clip_eps, float cam_dist, float max_depth, int flip_y,
                   int B, int Vz, int V,
                   const float* g_proj, const float* g_vox, const float* g_probs, const float* g_depth,
                   float* d_in, float* d_scale, void* stream) {
  if (!taps_z_rev) return DPC_ERR_NULL;
  return launch_conv_z_bwd(vox, mask2, scale, taps_z_rev, Kz, pad_lo_z_rev, 0, mode, clip_eps, cam_dist, max_depth,
                           flip_y, B, Vz, V, g_proj, g_vox, g_probs, g_depth, d_in, d_scale, stream);
}

// ------------------------------------------------------------------------------------ fused path
static inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

struct DpcScratch { float* raw; float* tmp; float* part; unsigned* cnt; float* d_tr; int64_t total; };
struct DpcSaved { uint32_t* mask1; uint32_t* mask2; int64_t total; };

static DpcScratch scratch_layout(const dpc_project_params* p, void* base) {
  const int64_t g = (int64_t)p->B * p->Vz * p->V * p->V;
  DpcScratch w;
  char* c = (char*)base;
  w.raw = (float*)c;
  w.tmp = (float*)(c + align256(g * 4));
  w.part = (float*)(c + 2 * align256(g * 4));      // dL/dscale partials of the depth-pass backward: [B, 32 tiles x 8 warps]
  w.cnt = (unsigned*)(c + 2 * align256(g * 4) + align256((int64_t)p->B * 256 * 4));   // per-sample completion counters (knob 4)
  w.total = 2 * align256(g * 4) + align256((int64_t)p->B * 256 * 4) + align256((int64_t)p->B * 4);
  w.d_tr = (float*)(c + w.total);                   // dL/d(tr_pc) of every point, from the fused x/y + gather kernel to the chain rule
  w.total += align256((int64_t)p->B * p->N * 12);
  return w;
}

static DpcSaved saved_layout(const dpc_project_params* p, void* base) {
  const int64_t g = (int64_t)p->B * p->Vz * p->V * p->V;
  const int64_t nw = (p->Vz + 31) / 32;
  DpcSaved w;
  char* c = (char*)base;
  w.mask1 = (uint32_t*)c;
  const int64_t m1 = align256((g / 32 + 1) * 4);
  w.mask2 = (uint32_t*)(c + m1);
  w.total = m1 + align256((int64_t)p->B * p->V * p->V * nw * 4);
  return w;
}

static int params_ok(const dpc_project_params* p) {
  if (!p) return DPC_ERR_NULL;
  if (!shape_ok(p->B, p->Vz, p->V) || p->N < 1) return DPC_ERR_SHAPE;
  if ((p->V * p->V) % 32 != 0) return DPC_ERR_SHAPE;
  if (p->K < 0 || p->K > DPC_MAX_TAPS || p->Kz < 0 || p->Kz > DPC_MAX_TAPS) return DPC_ERR_SHAPE;
  if ((p->K == 0) != (p->Kz == 0)) return DPC_ERR_ARG;
  if (p->mode < DPC_PROJ_DRC || p->mode > DPC_PROJ_DRC_PROD) return DPC_ERR_ARG;
  if (p->pose_kind < DPC_POSE_NONE || p->pose_kind > DPC_POSE_MATRIX) return DPC_ERR_ARG;
  return DPC_OK;
}

int64_t dpc_project_fast_scratch_bytes(const dpc_project_params* p) {
  if (params_ok(p) != DPC_OK) return -1;
  return scratch_layout(p, nullptr).total;
}

int64_t dpc_project_fast_saved_bytes(const dpc_project_params* p) {
  if (params_ok(p) != DPC_OK) return -1;
  return saved_layout(p, nullptr).total;
}

int dpc_project_fast_fwd(const dpc_project_params* p,
                         const float* pc, const float* pose, const float* trans, const float* focal,
                         const float* scale, const float* taps_xy, const float* taps_z,
                         float* tr_pc, float* voxels, float* proj, float* drc_probs, float* proj_depth,
                         void* scratch, int64_t scratch_bytes, void* saved, int64_t saved_bytes, void* stream) {
  DPC_TRY(params_ok(p));
  if (!pc || !voxels || !proj || !scratch || !saved) return DPC_ERR_NULL;
  if (p->K > 0 && (!taps_xy || !taps_z)) return DPC_ERR_NULL;
  if ((((uintptr_t)scratch) & 15) != 0 || (((uintptr_t)saved) & 15) != 0) return DPC_ERR_WORKSPACE;
  DpcScratch w = scratch_layout(p, scratch);
  DpcSaved sv = saved_layout(p, saved);
  if (scratch_bytes < w.total || saved_bytes < sv.total) return DPC_ERR_WORKSPACE;
  const int K = p->K > 0 ? p->K : 1, Kz = p->Kz > 0 ? p->Kz : 1;
  const float* tx = p->K > 0 ? taps_xy : nullptr;   // NULL taps = identity filter (kernel=None)
  const float* tz = p->Kz > 0 ? taps_z : nullptr;
  const int64_t g = (int64_t)p->B * p->Vz * p->V * p->V;
  stage_mark(0, stream);
  const float* hxy = (tx && p->taps_xy_host) ? p->taps_xy_host : nullptr;   // host copies of the taps (optional)
  const float* hz = (tz && p->taps_z_host) ? p->taps_z_host : nullptr;
  // cudaMemsetAsync beat a hand-written float4 zero kernel here (23.6 vs 28.4 us for memset + splat,
  // gpurun round 7), so the driver's memset stays.
  // Round 8: a zeroing KERNEL again, now as the first half of a PDL pair -- it waits for everything older, lets the
  // splat start, and the splat stages + transforms + writes tr_pc while the grid is being zeroed (knob 10).
  int splat_early = 0;
  if (!(p->flags & DPC_FLAG_SCRATCH_RAW_ZERO)) {
#if !defined(DPC_EMU) && defined(DPC_EXPERIMENTS)
    if (g_tune[10] == 2 && ((size_t)g * 4) % 16384 == 0) {
      DPC_LAUNCH(dpc_zero_bulk_kernel, dim3(148 * 2), dim3(128), 0, stream, (unsigned char*)w.raw, (size_t)g * 4);
      DPC_TRY(dpc_check_launch());
      splat_early = 1;
    } else
#endif
#ifdef DPC_EXPERIMENTS
    if (g_tune[10] && g_tune[10] != 3) {
      const size_t n4 = (size_t)g / 4;
#ifdef DPC_EMU
      const int zgrid = 2;
#else
      const int zgrid = 148 * 4;
#endif
      DPC_LAUNCH(dpc_zero_kernel, dim3(zgrid), dim3(256), 0, stream, (float4*)w.raw, n4, w.raw + n4 * 4, (int)(g - (int64_t)n4 * 4));
      DPC_TRY(dpc_check_launch());
      splat_early = 1;
    } else
#endif
    {
      // lab knob 10 = 3: the splat kernel zeroes the grid itself (cooperative launch, dpc_splat_fwd_warp_kernel<.., true>)
      bool done = false;
      if (g_tune[10] == 3 && !(p->sel) && !g_tune[21]) {
        DPC_TRY(splat_fwd_launch(pc, pose, p->pose_kind, trans, focal, p->focal_const, p->cam_dist, nullptr,
                                 p->B, p->N, p->Vz, p->V, tr_pc, w.raw, nullptr, nullptr, nullptr, nullptr, 0, stream, 0,
                                 nullptr, 0, &done));
      }
      if (done) splat_early = 2;      // the splat has been launched already
      else DPC_CUDA(cudaMemsetAsync(w.raw, 0, (size_t)g * 4, (cudaStream_t)stream));
    }
  }
  if (p->sel && p->N_src < p->N) return DPC_ERR_ARG;
  // x/y + depth pass as ONE persistent kernel (dpc_smooth_fused.cuh) when the three passes share one tap vector
  bool fused_xyz = false;
#if !defined(DPC_EMU) && defined(DPC_EXPERIMENTS)
  // lab knob 21; measured (profiles/r02_m_fused_fwd.md): correct, no gain -- the boundary it removes is worth ~1.4 us and the
  // in-kernel signalling + phase barrier cost more
  fused_xyz = g_tune[21] && !splat_early && dpc_tc_level() == 2 && p->V == 64 && p->Vz == 64 && K == Kz && tx == tz && hxy == hz &&
              !(p->flags & DPC_FLAG_SCRATCH_RAW_ZERO) && g_tune[15] && !drc_probs && !proj_depth;
#endif
  if (splat_early != 2)
    DPC_TRY(splat_fwd_launch(pc, pose, p->pose_kind, trans, focal, p->focal_const, p->cam_dist, nullptr,
                             p->B, p->N, p->Vz, p->V, tr_pc, w.raw, nullptr, nullptr, nullptr, p->sel, p->N_src, stream, splat_early,
                             fused_xyz ? w.cnt : nullptr, p->B));
  stage_mark(1, stream);
  // clip + x/y smoothing.  With DPC_FLAG_SCRATCH_RAW_ZERO the pass also hands the raw grid back
  // all-zero (so the next forward needs no memset).  Measured on B200 (profiles/r01_d): not a win --
  // the memset doubles as an L2 warm-up for the splat's reductions, which otherwise miss to HBM --
  // so the default path keeps the memset and leaves the flag to callers that want it.
  // In place unless the caller wants the raw grid handed back zeroed: every x/y tile is a pair of whole depth slices,
  // read completely before it is written, so the pass can overwrite its input -- one 32 MiB grid less per step in L2.
#if !defined(DPC_EMU) && defined(DPC_EXPERIMENTS)
  if (fused_xyz) {
    DpcConvZArgs az;
    dpc_set_taps_z(&az.ht, &az.use_ht, nullptr, 0, 0);
    az.in = w.raw; az.taps = tz; az.K = Kz; az.pl = (Kz - 1) / 2; az.rev = 0; az.scale = scale; az.mode = p->mode; az.eps = p->clip_eps;
    az.cam_dist = p->cam_dist; az.max_depth = p->max_depth; az.flip_y = 1; az.B = p->B; az.Vz = p->Vz; az.V = p->V; az.TY = 2;
    az.vox_out = voxels; az.mask2_out = scale ? sv.mask2 : nullptr; az.proj = proj; az.probs = nullptr; az.depth = nullptr;
    DPC_TRY(dpc_tcp_fwd_xyz_launch(w.raw, tx, K, (K - 1) / 2, sv.mask1, hxy, az, w.cnt, stream));
    DPC_TRY(dpc_check_launch());
    stage_mark(2, stream);
    stage_mark(3, stream);
    return DPC_OK;
  }
#endif
  float* xy_out = ((p->flags & DPC_FLAG_SCRATCH_RAW_ZERO) || !g_tune[15]) ? w.tmp : w.raw;
  DPC_TRY(launch_conv_xy(w.raw, xy_out, tx, K, (K - 1) / 2, tx, K, (K - 1) / 2,
                         p->B, p->Vz, p->V, /*clip_in=*/1, sv.mask1, nullptr, /*rev=*/0,
                         /*zero_in=*/(p->flags & DPC_FLAG_SCRATCH_RAW_ZERO) ? 1 : 0, stream, hxy, hxy));
  stage_mark(2, stream);
  const bool want_probs = (p->mode != DPC_PROJ_MAX);
  DPC_TRY(launch_conv_z_fwd(xy_out, tz, Kz, (Kz - 1) / 2, scale, p->mode, p->clip_eps, p->cam_dist,
                            p->max_depth, /*flip_y=*/1, p->B, p->Vz, p->V, voxels, scale ? sv.mask2 : nullptr, proj,
                            want_probs ? drc_probs : nullptr, want_probs ? proj_depth : nullptr, stream, hz));
  stage_mark(3, stream);
  return DPC_OK;
}

int dpc_project_fast_bwd(const dpc_project_params* p,
                         const float* pc, const float* pose, const float* trans, const float* focal,
                         const float* scale, const float* taps_xy, const float* taps_z,
                         const float* voxels,
                         const float* g_proj, const float* g_voxels, const float* g_tr_pc,
                         const float* g_probs, const float* g_depth,
                         float* d_pc, float* d_pose, float* d_trans, float* d_focal, float* d_scale,
                         void* scratch, int64_t scratch_bytes, const void* saved, int64_t saved_bytes, void* stream) {
  DPC_TRY(params_ok(p));
  if (!pc || !voxels || !scratch || !saved) return DPC_ERR_NULL;
  if (p->K > 0 && (!taps_xy || !taps_z)) return DPC_ERR_NULL;
  if ((((uintptr_t)scratch) & 15) != 0 || (((uintptr_t)saved) & 15) != 0) return DPC_ERR_WORKSPACE;
  DpcScratch w = scratch_layout(p, scratch);
  DpcSaved sv = saved_layout(p, const_cast<void*>(saved));
  if (scratch_bytes < w.total || saved_bytes < sv.total) return DPC_ERR_WORKSPACE;
  const int K = p->K > 0 ? p->K : 1, Kz = p->Kz > 0 ? p->Kz : 1;
  const float* tx = p->K > 0 ? taps_xy : nullptr;
  const float* tz = p->Kz > 0 ? taps_z : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  (void)st;
  const bool any_grid_grad = g_proj || g_voxels || g_probs || g_depth;
  DpcZero4Args z;
  z.p[0] = d_pose; z.n[0] = p->B * (p->pose_kind == DPC_POSE_MATRIX ? 16 : 4);
  z.p[1] = d_trans; z.n[1] = p->B * 3;
  z.p[2] = d_focal; z.n[2] = p->B;
  z.p[3] = d_scale; z.n[3] = p->B;
  // Training case on the tcgen05 pipelines: no zeroing launch at all.  The depth-pass backward (first kernel) zeroes
  // the splat backward's targets and leaves dL/dscale as per-warp partials, which the splat backward (last kernel)
  // folds.  Otherwise: one launch zeroes the (up to) four accumulation targets.
  bool fold_scale = false;
#ifndef DPC_EMU
  fold_scale = g_tune[13] == 0 && any_grid_grad && scale && d_scale && g_proj && !g_voxels && !g_probs && !g_depth &&
               (p->mode == DPC_PROJ_DRC || p->mode == DPC_PROJ_MAX) && dpc_tc_level() == 2 &&
               dpc_tc_conv_z_supported(p->V, p->Vz, Kz, false);
#endif
  // knob 4 (default): the splat backward runs inside the x/y pass (gather warps next to the pipeline warps) and starts
  // on a sample as soon as that pass has published it (per-sample counters, zeroed by the depth pass like the other targets)
  bool fused_gather = false;
#if !defined(DPC_EMU) && defined(DPC_EXPERIMENTS)
  fused_gather = fold_scale && p->tr_pc && g_tune[4] && g_tune[15] && !(p->flags & DPC_FLAG_SCRATCH_RAW_ZERO);
#endif
  if (fold_scale) {
    z.p[3] = fused_gather ? (float*)w.cnt : nullptr; z.n[3] = p->B;     // all-zero bits either way
  } else if (d_pose || d_trans || d_focal || d_scale) {
    DPC_LAUNCH(dpc_zero4_kernel, dim3(1), dim3(256), 0, stream, z);
    DPC_TRY(dpc_check_launch());
  }
  const float* d_raw = nullptr;
  stage_mark(4, stream);
  const float* hxy = (tx && p->taps_xy_host) ? p->taps_xy_host : nullptr;
  const float* hz = (tz && p->taps_z_host) ? p->taps_z_host : nullptr;
  if (any_grid_grad) {
    // voxels/proj -> dL/d(xy-smoothed) in G -> dL/d(raw) in G again (per-slice in place; the saved clip mask
    // applied).  G is the raw grid's storage (dead after the forward, and the forward zeroes it again) unless the
    // caller keeps `raw` all-zero between calls (DPC_FLAG_SCRATCH_RAW_ZERO): then G = tmp.
    float* G = ((p->flags & DPC_FLAG_SCRATCH_RAW_ZERO) || !g_tune[15]) ? w.tmp : w.raw;
    DPC_TRY(launch_conv_z_bwd(voxels, scale ? sv.mask2 : nullptr, scale, tz, Kz, Kz - 1 - (Kz - 1) / 2, /*rev=*/1,
                              p->mode, p->clip_eps, p->cam_dist, p->max_depth, /*flip_y=*/1, p->B, p->Vz, p->V,
                              g_proj, g_voxels, g_probs, g_depth, G, d_scale, stream, hz,
                              fold_scale ? w.part : nullptr, fold_scale ? &z : nullptr));
    stage_mark(5, stream);
#if !defined(DPC_EMU) && defined(DPC_EXPERIMENTS)
    if (fused_gather) {
      // x/y pass + the gathers in one kernel (dL/d(tr_pc) of every point into scratch), then the chain rule through the
      // camera: the splat backward without a grid to read (d_vox = NULL, d_tr_pc_in = the gathered gradients)
      DpcXYGatherArgs ga;
      ga.tr_pc = p->tr_pc; ga.d_raw = G; ga.g_tr_pc = g_tr_pc; ga.d_tr = w.d_tr; ga.sample_cnt = w.cnt;
      ga.B = p->B; ga.N = p->N; ga.dbg = g_tune[16];
      DPC_TRY(dpc_tcp_conv_xy_gather_launch(G, tx, K, K - 1 - (K - 1) / 2, (int64_t)p->B * p->Vz, sv.mask1, /*rev=*/1, hxy, ga,
                                            g_tune[17], stream));
      DPC_TRY(dpc_check_launch());
      stage_mark(6, stream);
      if (g_tune[17] == 2)       // pricing experiment: pipeline + signalling only, the whole splat backward afterwards
        DPC_TRY(splat_bwd_launch(pc, pose, p->pose_kind, trans, focal, p->focal_const, p->cam_dist, nullptr, 0,
                                 p->B, p->N, p->Vz, p->V, G, nullptr, g_tr_pc, d_pc, d_pose, d_trans, d_focal, nullptr,
                                 w.part, 256, d_scale, stream, p->sel, p->N_src));
      else
      DPC_TRY(splat_bwd_launch(pc, pose, p->pose_kind, trans, focal, p->focal_const, p->cam_dist, nullptr, 0,
                               p->B, p->N, p->Vz, p->V, nullptr, nullptr, w.d_tr, d_pc, d_pose, d_trans, d_focal, nullptr,
                               w.part, 256, d_scale, stream, p->sel, p->N_src));
      stage_mark(7, stream);
      return DPC_OK;
    }
#endif
    DPC_TRY(launch_conv_xy(G, G, tx, K, K - 1 - (K - 1) / 2, tx, K, K - 1 - (K - 1) / 2,
                           p->B, p->Vz, p->V, /*clip_in=*/0, nullptr, sv.mask1, /*rev=*/1, /*zero_in=*/0, stream, hxy, hxy));
    d_raw = G;
  }
  stage_mark(6, stream);
  DPC_TRY(splat_bwd_launch(pc, pose, p->pose_kind, trans, focal, p->focal_const, p->cam_dist, nullptr, 0,
                           p->B, p->N, p->Vz, p->V, d_raw, nullptr, g_tr_pc, d_pc, d_pose, d_trans, d_focal, nullptr,
                           fold_scale ? w.part : nullptr, 256, fold_scale ? d_scale : nullptr, stream, p->sel, p->N_src,
                           p->tr_pc));
  stage_mark(7, stream);
  return DPC_OK;
}

int dpc_project_rgb_fwd(const float* probs, const float* rgb, int B, int Vz, int V, float* proj_rgb, void* stream) {
  if (!probs || !rgb || !proj_rgb) return DPC_ERR_NULL;
  if (!shape_ok(B, Vz, V)) return DPC_ERR_SHAPE;
  const long long rays = (long long)B * V * V;
  DPC_LAUNCH(dpc_project_rgb_fwd_kernel, dim3((unsigned)((rays + 127) / 128)), dim3(128), 0, stream, probs, rgb, B, Vz, V, proj_rgb);
  return dpc_check_launch();
}

int dpc_project_rgb_bwd(const float* probs, const float* rgb, const float* g_proj_rgb, int B, int Vz, int V,
                        float* d_probs, float* d_rgb, void* stream) {
  if (!probs || !rgb || !g_proj_rgb || (!d_probs && !d_rgb)) return DPC_ERR_NULL;
  if (!shape_ok(B, Vz, V)) return DPC_ERR_SHAPE;
  const long long rays = (long long)B * V * V;
  DPC_LAUNCH(dpc_project_rgb_bwd_kernel, dim3((unsigned)((rays + 127) / 128)), dim3(128), 0, stream, probs, rgb, g_proj_rgb, B, Vz, V,
             d_probs, d_rgb);
  return dpc_check_launch();
}

int dpc_tap_corr(const float* a, const float* g, int axis, int B, int Vz, int V, int K, int pad_lo, float* out, void* stream) {
  if (!a || !g || !out) return DPC_ERR_NULL;
  if (!shape_ok(B, Vz, V) || K < 1 || K > DPC_MAX_TAPS) return DPC_ERR_SHAPE;
  if (axis < 0 || axis > 2 || pad_lo < 0 || pad_lo >= K) return DPC_ERR_ARG;
  const long long nvox = (long long)B * Vz * V * V;
  DPC_CUDA(cudaMemsetAsync(out, 0, (size_t)K * 4, (cudaStream_t)stream));
  const long long ctas = (nvox + DPC_TAPCORR_CHUNK - 1) / DPC_TAPCORR_CHUNK;
  if (ctas > 2147483647LL) return DPC_ERR_SHAPE;
  DPC_LAUNCH(dpc_tap_corr_kernel, dim3((unsigned)ctas), dim3(DPC_TAPCORR_THREADS), 0, stream, a, g, axis, nvox, Vz, V, K, pad_lo, out);
  return dpc_check_launch();
}

int dpc_dropout_indices(unsigned long long seed, unsigned long long draw, const unsigned long long* state,
                        int B, int N, int n_keep, int32_t* sel, void* stream) {
  if (!sel) return DPC_ERR_NULL;
  if (B < 1 || B > 65535 || N < 1 || n_keep < 0 || n_keep > N) return DPC_ERR_SHAPE;
  if (n_keep == 0) return DPC_OK;
  DPC_LAUNCH(dpc_dropout_indices_kernel, dim3((n_keep + 255) / 256, B), dim3(256), 0, stream, seed, draw, state, N, n_keep, sel);
  return dpc_check_launch();
}

int dpc_gather_points(const float* in, const int64_t* sel, int B, int N, int n_keep, int C, float* out, void* stream) {
  if (!in || !sel || !out) return DPC_ERR_NULL;
  if (B < 1 || B > 65535 || N < 1 || n_keep < 0 || n_keep > N || C < 1) return DPC_ERR_SHAPE;
  if (n_keep == 0) return DPC_OK;
  dim3 grid((n_keep * C + 255) / 256, B);
  DPC_LAUNCH(dpc_gather_kernel, grid, dim3(256), 0, stream, in, sel, N, n_keep, C, out);
  return dpc_check_launch();
}

int dpc_gather_points_bwd(const float* g_out, const int64_t* sel, int B, int N, int n_keep, int C, float* g_in, void* stream) {
  if (!g_out || !sel || !g_in) return DPC_ERR_NULL;
  if (B < 1 || B > 65535 || N < 1 || n_keep < 0 || n_keep > N || C < 1) return DPC_ERR_SHAPE;
  DPC_CUDA(cudaMemsetAsync(g_in, 0, (size_t)B * N * C * 4, (cudaStream_t)stream));
  if (n_keep == 0) return DPC_OK;
  dim3 grid((n_keep * C + 255) / 256, B);
  DPC_LAUNCH(dpc_gather_bwd_kernel, grid, dim3(256), 0, stream, g_out, sel, N, n_keep, C, g_in);
  return dpc_check_launch();
}

}  // extern "C"

// ------------------------------------------------------------------------------------ f-4: nearest neighbours
template <typename T>
static int nn_launch(const T* vs, int ns, const T* vt, int nt, T* proj, T* min_dist, int32_t* idx,
                     void* workspace, int64_t workspace_bytes, void* stream) {
  if (!vs || !vt || !workspace) return DPC_ERR_NULL;
  if (ns < 1 || nt < 1) return DPC_ERR_SHAPE;
  const int splits = dpc_nn_splits(ns, nt);
  const int64_t need = (int64_t)splits * ns * (int64_t)(sizeof(T) + sizeof(int32_t));
  if (workspace_bytes < need || (((uintptr_t)workspace) & 7) != 0) return DPC_ERR_WORKSPACE;
  T* part_s = (T*)workspace;
  int32_t* part_i = (int32_t*)(part_s + (size_t)splits * ns);
  int chunk = (nt + splits - 1) / splits;
  chunk = ((chunk + DPC_NN_TILE - 1) / DPC_NN_TILE) * DPC_NN_TILE;
  const int bx = (ns + DPC_NN_THREADS - 1) / DPC_NN_THREADS;
  DPC_LAUNCH(dpc_nn_partial_kernel<T>, dim3(bx, splits), dim3(DPC_NN_THREADS), 0, stream, vs, ns, vt, nt, chunk, part_s, part_i);
  DPC_TRY(dpc_check_launch());
  DPC_LAUNCH(dpc_nn_final_kernel<T>, dim3(bx), dim3(DPC_NN_THREADS), 0, stream, vt, ns, splits, (const T*)part_s,
             (const int32_t*)part_i, proj, min_dist, idx);
  return dpc_check_launch();
}

extern "C" {

int64_t dpc_proj_l2_loss_workspace_bytes(void) { return (int64_t)(DPC_LOSS_MAX_CTAS + 1) * 4; }

int dpc_proj_l2_loss(const float* pred, const float* gt, int64_t n, float inv_count, float* loss, float* g_pred,
                     void* workspace, int64_t workspace_bytes, void* stream) {
  if (!pred || !gt || (!loss && !g_pred)) return DPC_ERR_NULL;
  if (loss && !workspace) return DPC_ERR_NULL;
  if (n < 1) return DPC_ERR_SHAPE;
  if (loss && (workspace_bytes < dpc_proj_l2_loss_workspace_bytes() || (((uintptr_t)workspace) & 3) != 0)) return DPC_ERR_WORKSPACE;
  int64_t ctas = (n / 4 + DPC_LOSS_THREADS - 1) / DPC_LOSS_THREADS;      // one float4 per thread where that is enough
  if (ctas < 1) ctas = 1;
  if (ctas > 148 * 2) ctas = 148 * 2;
  float* partial = (float*)workspace;
  unsigned* counter = (unsigned*)(partial + DPC_LOSS_MAX_CTAS);
  DPC_LAUNCH(dpc_proj_l2_loss_kernel, dim3((unsigned)ctas), dim3(DPC_LOSS_THREADS), 0, stream, pred, gt, (long long)n, inv_count,
             loss, g_pred, partial, counter);
  return dpc_check_launch();
}

int64_t dpc_point_cloud_distance_workspace_bytes(int ns, int nt, int elem_bytes) {
  if (ns < 1 || nt < 1 || (elem_bytes != 4 && elem_bytes != 8)) return -1;
  return (int64_t)dpc_nn_splits(ns, nt) * ns * (int64_t)(elem_bytes + 4);
}

int dpc_point_cloud_distance_f32(const float* vs, int ns, const float* vt, int nt, float* proj, float* min_dist,
                                 int32_t* idx, void* workspace, int64_t workspace_bytes, void* stream) {
  return nn_launch<float>(vs, ns, vt, nt, proj, min_dist, idx, workspace, workspace_bytes, stream);
}

int dpc_point_cloud_distance_f64(const double* vs, int ns, const double* vt, int nt, double* proj, double* min_dist,
                                 int32_t* idx, void* workspace, int64_t workspace_bytes, void* stream) {
  return nn_launch<double>(vs, ns, vt, nt, proj, min_dist, idx, workspace, workspace_bytes, stream);
}

}  // extern "C"
