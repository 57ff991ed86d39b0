/* dpc_b200.h -- C-ABI of the B200-native differentiable point-cloud projection path.
 *
 * Drop-in boundary for the hot path of eldar/differentiable-point-clouds:
 *   dpc/util/point_cloud.py:157-216  pc_perspective_transform      -> dpc_splat_fwd (pose stage)
 *   dpc/util/point_cloud.py:60-136   pointcloud2voxels3d_fast      -> dpc_splat_fwd / dpc_splat_bwd
 *   dpc/util/point_cloud.py:139-145  smoothen_voxels3d             -> dpc_conv_xy + dpc_conv_z_fwd
 *   dpc/util/drc.py:47-123,139-153   drc_projection / depth        -> dpc_conv_z_fwd / dpc_conv_z_bwd (Kz = 1)
 *   dpc/util/point_cloud.py:229-290  pointcloud_project_fast       -> dpc_project_fast_fwd / _bwd
 *
 * The reference has no FFI of its own (it is Python calling TensorFlow ops); these are the
 * entry points a binding for this path would need -- see INTEGRATION.md for the ctypes stub.
 *
 * Conventions
 *   - plain pointers to DEVICE memory (fp32 unless stated), C-contiguous, reference axis order:
 *       point clouds [B,N,3]; grids [B,Vz,V,V]; silhouettes [B,V,V]; drc_probs [Vz+1,B,V,V].
 *     tr_pc channel 0 is depth and indexes the Vz axis (point_cloud.py:215).
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work (no host sync),
 *     keeps no static mutable state and is re-entrant.
 *   - every function returns 0 on success or a negative DPC_ERR_* code; nothing throws.
 *     dpc_last_cuda_error() reports the CUDA error behind DPC_ERR_CUDA for the calling thread.
 *   - nullable arguments are marked; outputs are written by the callee, never read first unless
 *     stated ("accumulates").
 */
#ifndef DPC_B200_H_
#define DPC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPC_OK 0
#define DPC_ERR_NULL (-1)     /* a required pointer is NULL                         */
#define DPC_ERR_SHAPE (-2)    /* B/N/V/Vz/K out of the supported range              */
#define DPC_ERR_ARG (-3)      /* inconsistent flags (e.g. translation with a matrix pose) */
#define DPC_ERR_CUDA (-4)     /* a CUDA call failed; see dpc_last_cuda_error()      */
#define DPC_ERR_WORKSPACE (-5)/* workspace too small                                */

#define DPC_POSE_NONE (-1)   /* points are already in camera space (pointcloud2voxels3d_fast alone) */
#define DPC_POSE_QUAT 0      /* [B,4] unnormalised quaternion, w first (cfg.pose_quaternion) */
#define DPC_POSE_MATRIX 1    /* [B,4,4] extrinsic; intrinsic diag(1,f,f,1) applied inside (camera.py:5-13) */

#define DPC_PROJ_NONE (-1)      /* no projection (smoothen_voxels3d alone)            */
#define DPC_PROJ_DRC 0          /* ray termination sum, drc_logsum quirks (drc.py:47-123) */
#define DPC_PROJ_MAX 1          /* tf.reduce_max over depth (point_cloud.py:265)      */
#define DPC_PROJ_DRC_PROD 2     /* drc_logsum=false variant (cumprod, unity 1)        */

#define DPC_MAX_TAPS 63
#define DPC_MAX_V 128

int dpc_abi_version(void);
const char* dpc_error_string(int code);
int dpc_last_cuda_error(void);
/* Diagnostics switches of the PRODUCT library (process-wide, not thread-safe, off by default; they change no result):
 *   3      1 = record CUDA events around every stage of the fused forward / backward (dpc_debug_stage_ms)
 *   8      kernel family of the 64^3 smoothing passes: 0 FFMA2 (CUDA cores), 1 single-tile tcgen05, 2 tcgen05 pipelines
 *          (default; falls back to 1 when the driver has no tensor-map encoder).  Exists so that the three families can
 *          be tested behind the same ABI (tests/test_gpu_families.py).
 *   12     1 = arm the per-kernel timeline (dpc_debug_ktrace_read)
 * Every other key is an EXPERIMENT knob and exists only in the lab build of the same sources (-DDPC_EXPERIMENTS,
 * libdpc_b200_lab.so; dpc_is_lab_build() == 1); the product build returns DPC_ERR_ARG for them, compiles their defaults
 * in as constants and does not contain the experimental kernels.  Lab keys (every value gives the same results --
 * tests/test_gpu_parity.py::test_splat_variants_full_shape):
 *   0 / 1  points per thread of the tile-per-CTA forward / backward splat kernel (1|2|4; defaults 4 / 1).  Key 0 also
 *          selects the forward kernel: 4 (default) = the software-pipelined dpc_splat_fwd_warp_kernel, 8 = the
 *          tile-per-CTA kernel with 4 points per thread, 1 / 2 = that kernel with 1 / 2
 *   2      1 = the producer warps of the x/y pipeline store the finished tiles; 2 = 3-slot staging ring (default 0)
 *   4      1 = the gathers of the splat backward run inside the x/y pass of the backward (csrc/dpc_fused_bwd.cuh; needs
 *          dpc_project_params.tr_pc; default 0: measured slower, profiles/r02_f_fused_gather.md); 16 = its debug flags,
 *          17 = 1: its 768-thread variant
 *   5      1 = ignore host copies of the taps (vector-register FFMA2 kernels / device taps in the pipelines)
 *   6      1 = cp.async tile load in the FFMA2 depth kernels
 *   7      conv_xy (FFMA2) diagnostics: 1 memory path only, 2 arithmetic only, 3 empty CTAs, 4/5 skip the x / y correlation
 *   9      1 = per-tile / per-CTA trace of the pipelines (dpc_debug_trace_read)
 *   10     zeroing of the raw grid in the fused forward: 0 cudaMemsetAsync + wait-first splat (default), 1 store kernel,
 *          2 TMA bulk-store kernel -- 1 and 2 as PDL primaries of a splat that transforms ahead of its grid dependency;
 *          3 = the forward splat kernel zeroes the grid itself (cooperative launch, grid barrier before the reductions)
 *   11     1 = 16-byte reductions / gathers in the splats (default), 0 = 8-byte / scalar
 *   13     1 = keep the zeroing launch + dL/dscale atomics in the fused backward (default 0: folded partial sums)
 *   14     1 = the backward splat stages + transforms ahead of its grid dependency (default)
 *   15     1 = x/y pass in place, backward in the raw grid's storage: two grids per step (default)
 *   18-20  tile-per-CTA splat backward: 128-thread CTAs / compiled for 75 % occupancy / gather style (1 = independent
 *          un-guarded loads, 2 = one guarded path, 3 = predicated independent loads); 22 = 1: gathers through
 *          ld.global.cg (profiles/r02_j_splat_bwd_occupancy.md).  Key 20 also selects the kernel: 0 (default) or 4 = the
 *          software-pipelined dpc_splat_bwd_warp_kernel whenever dpc_project_params.tr_pc is given, 6 = the tile-per-CTA
 *          kernel; with the pipelined kernels key 19 = the number of warps per SM the grid is sized for (default 28)
 *          (profiles/r02_w_splat_bwd_warp.md)
 *   21     1 = x/y pass + depth pass of the forward as ONE persistent kernel (csrc/dpc_smooth_fused.cuh; default 0:
 *          measured no faster, profiles/r02_m_fused_fwd.md) */
int dpc_debug_set(int key, int value);
int dpc_is_lab_build(void);
/* diagnostics: after dpc_debug_set(12, 1), every kernel of the fused path folds %globaltimer (ns) into
 * out[4*k + {0: first CTA entry, 1: first CTA past its grid dependency, 2: last CTA past it, 3: last CTA exit}],
 * k = 0 zero, 1 splat fwd, 2 x/y fwd, 3 depth fwd, 4 zero4, 5 depth bwd, 6 x/y bwd, 7 splat bwd; 64 uint64 to host. */
int dpc_debug_ktrace_read(unsigned long long* host_out);
/* key 3 = 1: record CUDA events around every stage of the fused forward/backward; this returns the
 * six stage durations (ms) of the last instrumented step (synchronises on the last event). */
int dpc_debug_stage_ms(float* out6);
#ifdef DPC_EXPERIMENTS
/* lab build only.  phase_read: thread 0 of every CTA (first 512) of the splat kernels stamps its phases
 * (out[((which * 512 + cta) * 8 + slot]); mma_bench: tcgen05.mma issue / latency micro-benchmark (3 int64 per CTA in
 * device memory); trace_read: per-tile hand-offs of CTA 0 and entry / set-up / exit of every CTA of the pipelines. */
int dpc_debug_phase_read(unsigned long long* host_out);
int dpc_debug_mma_bench(long long* out, int nctas, int threads, int reps, int nmma, int spin, int M, int N, void* stream);
int dpc_debug_trace_read(long long* host_out);
/* gather_bench: the splat backward's access pattern alone (B x N threads each gather the 2 x 2 rows of a random cell of a
 * [B,V,V,V] grid); variant 0 = the product kernel's loads issued ideally, 1-6 see csrc/dpc_gather_bench.cuh;
 * scripts/gather_bench.py */
int dpc_debug_gather_bench(const float* grid, float* out, int B, int N, int V, int variant, int threads, int ppt,
                           int share_log2, int smem_pad, int cg, void* stream);
#endif
/* compiled for sm_100a?  1 = real CUDA build, 0 = the CPU emulation build used by tests/emu */
int dpc_is_cuda_build(void);

/* ---- K1: camera transform + validity + trilinear splat, one fused kernel -----------------
 * pose: [B,4] or [B,4,4] per pose_kind (NULL for DPC_POSE_NONE).  trans: [B,3] or NULL (quaternion
 * pose only).  focal: [B] or NULL (then focal_const; the matrix pose always uses focal_const).
 * rgb: [B,N,3] or NULL.  Outputs: tr_pc [B,N,3] (nullable), vox [B,Vz,V,V] (nullable = transform
 * only; ACCUMULATES: caller zeroes it, dpc_project_fast_fwd does so itself), vox_rgb [B,Vz,V,V,3]
 * (required iff rgb; accumulates), idx_out int32 [B,N,3] = floor((p+0.5)*(S-1)) (nullable; the
 * bit-exact parity gate), valid_out uint8 [B,N] (nullable). */
int dpc_splat_fwd(const float* pc, const float* pose, int pose_kind, const float* trans,
                  const float* focal, float focal_const, float cam_dist, const float* rgb,
                  int B, int N, int Vz, int V,
                  float* tr_pc, float* vox, float* vox_rgb, int32_t* idx_out, uint8_t* valid_out,
                  void* stream);

/* ---- K1b: backward of K1.  d_vox [B,Vz,V,V] (nullable), d_vox_rgb (nullable), d_tr_pc_in
 * [B,N,3] = gradient arriving directly at the tr_pc output (nullable).  rgb_stop_grad mirrors
 * cfg.pc_rgb_stop_points_gradient.  Outputs (each nullable): d_pc [B,N,3]; d_pose [B,4]|[B,4,4],
 * d_trans [B,3], d_focal [B] ACCUMULATE (caller zeroes); d_rgb [B,N,3]. */
int dpc_splat_bwd(const float* pc, const float* pose, int pose_kind, const float* trans,
                  const float* focal, float focal_const, float cam_dist, const float* rgb,
                  int rgb_stop_grad, int B, int N, int Vz, int V,
                  const float* d_vox, const float* d_vox_rgb, const float* d_tr_pc_in,
                  float* d_pc, float* d_pose, float* d_trans, float* d_focal, float* d_rgb,
                  void* stream);

/* ---- K2a: per depth-slice: [clip to [0,1]] -> correlate along x (taps_x) -> along y (taps_y)
 * -> [multiply by a saved pass mask].  Zero padding, pad_lo taps to the low side (TF "SAME":
 * (K-1)/2).  The backward of a correlation with (taps, pad_lo) is this same call with reversed
 * taps and pad_lo' = K-1-pad_lo.
 *   clip_in != 0      : input is clipped to [0,1] first (point_cloud.py:240)
 *   mask_bits_out     : uint32 [B*Vz*V*V/32] (nullable) bit = (0 <= in <= 1), the clip's pass mask
 *   mask_bits_in      : (nullable) output is multiplied by the saved mask (backward of the clip) */
int dpc_conv_xy(const float* in, float* out, const float* taps_x, int Kx, int pad_lo_x,
                const float* taps_y, int Ky, int pad_lo_y, int B, int Vz, int V,
                int clip_in, uint32_t* mask_bits_out, const uint32_t* mask_bits_in, void* stream);

/* ---- K2b+K3 forward: correlate along depth (taps_z) -> [* scale[b], clip to [0,1]] ->
 * projection along depth, one kernel (the depth axis is resident per ray).
 *   scale [B] nullable.  mode: DPC_PROJ_*.  flip_y: write proj / drc_probs with the image-row
 *   axis reversed (point_cloud.py:270,273); the standalone drc_projection uses flip_y = 0.
 *   Outputs: vox_out [B,Vz,V,V]; mask2_out uint32 [B*Vz*V*V/32]... stored per ray as
 *   ceil(Vz/32) words: [B,V,V,ceil(Vz/32)] (nullable) bit z = (0 <= scale*smoothed <= 1);
 *   proj [B,V,V] (nullable iff mode NONE); drc_probs [Vz+1,B,V,V] (nullable);
 *   proj_depth [B,V,V] (nullable; needs DRC mode). */
int dpc_conv_z_fwd(const float* in, const float* taps_z, int Kz, int pad_lo_z,
                   const float* scale, int mode, float clip_eps, float cam_dist, float max_depth,
                   int flip_y, int B, int Vz, int V,
                   float* vox_out, uint32_t* mask2_out, float* proj, float* drc_probs,
                   float* proj_depth, void* stream);

/* ---- K3b+K2b backward: from the gradients at proj / voxels / drc_probs / proj_depth (each
 * nullable) back through projection, clip, scale and the depth correlation.
 *   vox [B,Vz,V,V] is the forward's vox_out; mask2 its mask2_out (NULL = no scale/clip stage);
 *   taps_z_rev / pad_lo_z_rev are the reversed taps (see dpc_conv_xy).
 *   Outputs: d_in [B,Vz,V,V] gradient w.r.t. the forward's `in`; d_scale [B] ACCUMULATES (nullable). */
int dpc_conv_z_bwd(const float* vox, const uint32_t* mask2, const float* scale,
                   const float* taps_z_rev, int Kz, int pad_lo_z_rev,
                   int mode, float clip_eps, float cam_dist, float max_depth, int flip_y,
                   int B, int Vz, int V,
                   const float* g_proj, const float* g_vox, const float* g_probs, const float* g_depth,
                   float* d_in, float* d_scale, void* stream);

/* ---- the fused pipeline of pointcloud_project_fast (no rgb) -------------------------------
 * Two caller-owned device buffers (16-byte aligned):
 *   scratch (dpc_project_fast_scratch_bytes): raw grid + one intermediate grid.  Carries NO state
 *     between calls and may be shared by any number of forward/backward calls on one stream.  If
 *     DPC_FLAG_SCRATCH_RAW_ZERO is set the caller guarantees the first half (the raw grid) is
 *     all-zero on entry and the forward skips its memset; every forward leaves it all-zero again
 *     (the smoothing pass zeroes each slice as it consumes it) and the backward never writes it.
 *   saved (dpc_project_fast_saved_bytes): the two clip masks as bit planes; written by the forward,
 *     read by the backward of the same call -- keep it until then.
 * taps: fp32 device arrays of K / Kz taps (K = Kz = 0: no smoothing kernel, taps ignored); the
 * backward takes the SAME taps (it reads them back to front).
 * taps_xy_host / taps_z_host (optional, may be NULL): the SAME tap values in HOST memory, read
 *   during the call (not retained).  The reference derives sigma from the global step
 *   (model_pc.py:35-40), so a training loop knows the taps on the host.  Given them, the smoothing
 *   kernels receive the taps as launch parameters (constant bank -> uniform registers), need half
 *   the registers and run 5-6 CTAs per SM instead of 3-4 (~10 % faster forward+backward at 64^3,
 *   K = 21).  The caller guarantees host and device taps hold identical values. */
#define DPC_FLAG_SCRATCH_RAW_ZERO 1

typedef struct {
  int B, N, Vz, V;
  int pose_kind;        /* DPC_POSE_* */
  int mode;             /* DPC_PROJ_DRC | DPC_PROJ_MAX | DPC_PROJ_DRC_PROD */
  int K, Kz;            /* tap counts along x/y and along depth; 0 = no smoothing */
  float focal_const, cam_dist, clip_eps, max_depth;
  int flags;            /* DPC_FLAG_* */
  const float* taps_xy_host;   /* optional host copy of taps_xy (K floats), or NULL */
  const float* taps_z_host;    /* optional host copy of taps_z (Kz floats), or NULL */
  const float* tr_pc;          /* backward only, optional (device, [B,N,3]): the tr_pc the forward of this call wrote.
                                  With it the splat backward runs in its software-pipelined form: the cell of a point is
                                  known without redoing the camera transform, so the gathers of the next 32 points are
                                  prefetched while the current ones are computed (10 us instead of 20 us at B=32,
                                  N=8000).  NULL = the tile-per-CTA kernel, which recomputes the transform first.  The
                                  results are the same either way, and they do not depend on the CONTENT of tr_pc: the
                                  transform is recomputed for the chain rule anyway, and a point whose recomputed cell is
                                  not the one its gathers were aimed at (a stale tr_pc) fetches its corners again */
  const int32_t* sel;          /* optional (device, [B,N] int32): point dropout consumed by the splat's load stage
                                  (point_cloud.py:293-319).  pc and d_pc are then [B,N_src,3]; point i of sample b is
                                  pc[b, sel[b*N + i]] (indices of a sample distinct), tr_pc stays [B,N,3]; dropped points
                                  are never read and get a zero gradient.  NULL = pc is [B,N,3] */
  int N_src;                   /* points per sample in pc when sel != NULL */
} dpc_project_params;

int64_t dpc_project_fast_scratch_bytes(const dpc_project_params* p);
int64_t dpc_project_fast_saved_bytes(const dpc_project_params* p);

int dpc_project_fast_fwd(const dpc_project_params* p,
                         const float* pc, const float* pose, const float* trans, const float* focal,
                         const float* scale, const float* taps_xy, const float* taps_z,
                         float* tr_pc, float* voxels, float* proj, float* drc_probs, float* proj_depth,
                         void* scratch, int64_t scratch_bytes, void* saved, int64_t saved_bytes, void* stream);

int dpc_project_fast_bwd(const dpc_project_params* p,
                         const float* pc, const float* pose, const float* trans, const float* focal,
                         const float* scale, const float* taps_xy, const float* taps_z,
                         const float* voxels,
                         const float* g_proj, const float* g_voxels, const float* g_tr_pc,
                         const float* g_probs, const float* g_depth,
                         float* d_pc, float* d_pose, float* d_trans, float* d_focal, float* d_scale,
                         void* scratch, int64_t scratch_bytes, const void* saved, int64_t saved_bytes, void* stream);

/* ---- f-3: K3 of the colour grid.  Replaces project_volume_rgb_integral (drc.py:126-136):
 *   proj_rgb[b,y,x,c] = sum_{z<Vz} probs[z,b,y,x] * rgb[b,z,y,x,c] + probs[Vz,b,y,x]      (white background)
 * probs [Vz+1,B,V,V] (drc_probs as pointcloud_project_fast returns them), rgb [B,Vz,V,V,3] (voxels_rgb as returned),
 * proj_rgb [B,V,V,3].  Backward: d_probs [Vz+1,B,V,V] and d_rgb [B,Vz,V,V,3] (either may be NULL), fully written. */
int dpc_project_rgb_fwd(const float* probs, const float* rgb, int B, int Vz, int V, float* proj_rgb, void* stream);
int dpc_project_rgb_bwd(const float* probs, const float* rgb, const float* g_proj_rgb, int B, int Vz, int V,
                        float* d_probs, float* d_rgb, void* stream);

/* ---- N1: gradient of a zero-padded correlation w.r.t. its taps, out[j] = sum_pos g[pos] * a[pos + (j - pad_lo) along
 * axis] (axis 0 = depth, 1 = y, 2 = x; a = the pass's input, g = the gradient at its output, both [B,Vz,V,V]).  With the
 * closed form of d taps / d sigma (gauss_kernel.py:5-11) this gives dL/dsigma of smoothen_voxels3d
 * (point_cloud.py:139-145); out [K] is zeroed by the callee.  Optional, off the hot path: sigma is a function of the
 * step count in the reference and nothing consumes its gradient. */
int dpc_tap_corr(const float* a, const float* g, int axis, int B, int Vz, int V, int K, int pad_lo, float* out, void* stream);

/* ---- f-2: the subsets of pc_point_dropout (point_cloud.py:296-311: np.random.choice(N, n_keep, replace=False) per
 * sample, on the host, through tf.py_func) drawn on the device: sel[b, i] = pi_b(i), i < n_keep, with pi_b a pseudo-random
 * permutation of [0, N) (Feistel network keyed by Philox4x32-10 of (b, draw) under `seed`).  Distinct by construction, no
 * sort.  state (nullable, device, 2 x uint64 {seed, draw}): read on the device instead of the two arguments, so that a
 * captured CUDA graph draws anew on every replay.  sel feeds dpc_project_params.sel. */
int dpc_dropout_indices(unsigned long long seed, unsigned long long draw, const unsigned long long* state,
                        int B, int N, int n_keep, int32_t* sel, void* stream);

/* ---- f-2, materialising form: out[b,i,:] = in[b, sel[b,i], :] (point_cloud.py:312-318) and its
 * backward (scatter-add).  sel int64 [B,n_keep]. */
int dpc_gather_points(const float* in, const int64_t* sel, int B, int N, int n_keep, int C,
                      float* out, void* stream);
int dpc_gather_points_bwd(const float* g_out, const int64_t* sel, int B, int N, int n_keep, int C,
                          float* g_in /* zeroed by callee */, void* stream);

/* ---- f-1 (loss row): the silhouette loss of the training step and its gradient in one pass.  Replaces
 * `tf.nn.l2_loss(gt - pred) / num_samples` (models/model_pc.py:414-415) and its autodiff:
 *   *loss = sum((gt - pred)^2) / 2 * inv_count,   g_pred[i] = (pred[i] - gt[i]) * inv_count   (g_pred may be NULL).
 * pred / gt / g_pred: n floats.  loss may be NULL as well (gradient only: no reduction, workspace unused).  workspace: dpc_proj_l2_loss_workspace_bytes() bytes of device memory, 4-byte aligned,
 * ZEROED ONCE by the caller before its first use (per-CTA partial sums + a self-resetting completion counter; calls on
 * one stream may share it).  Deterministic (fixed summation order). */
int64_t dpc_proj_l2_loss_workspace_bytes(void);
int dpc_proj_l2_loss(const float* pred, const float* gt, int64_t n, float inv_count, float* loss, float* g_pred,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* ---- f-4: nearest-neighbour projection of the chamfer evaluation.  Replaces point_cloud_distance
 * (util/point_cloud_distance.py:26-39; driver run/eval_chamfer.py:18-34): for every source point vs[i] the closest
 * target: min_dist[i] = min_j sqrt(sum((vt[j] - vs[i])^2)), idx[i] = the FIRST j attaining it (tf.argmin over the
 * rounded distances), proj[i,:] = vt[idx[i],:].  vs [ns,3], vt [nt,3] row-major, fp32 or fp64 (the evaluation feeds
 * fp64).  Any of proj / min_dist / idx may be NULL.  workspace: dpc_point_cloud_distance_workspace_bytes(ns, nt,
 * sizeof element) bytes of device memory, 8-byte aligned.  Nothing of size ns*nt is materialised, so the source
 * need not be cut into pc_eval_chamfer_num_parts pieces. */
int64_t dpc_point_cloud_distance_workspace_bytes(int ns, int nt, int elem_bytes);
int dpc_point_cloud_distance_f32(const float* vs, int ns, const float* vt, int nt, float* proj, float* min_dist,
                                 int32_t* idx, void* workspace, int64_t workspace_bytes, void* stream);
int dpc_point_cloud_distance_f64(const double* vs, int ns, const double* vt, int nt, double* proj, double* min_dist,
                                 int32_t* idx, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPC_B200_H_ */
