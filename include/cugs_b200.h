/* cugs_b200.h — C ABI of the B200-native (sm_100a) Gaussian-splatting rasterizer hot path.
 *
 * Drop-in boundary for the rasterizer of Artemarius/cuda-gaussian-splatting. Every entry point
 * below takes raw DEVICE pointers, sizes and an explicit cudaStream_t (passed as void* so that
 * this header needs no CUDA include); nothing here allocates caller-visible memory, throws, or
 * synchronises the device except where documented. Return value: 0 = success, < 0 = argument
 * error (CUGS_ERR_*), > 0 = cudaError_t of the failing runtime call / launch. After a non-zero
 * return cugs_b200_last_error(h) holds a message.
 *
 * Tensor layouts are the reference's (citations relative to /root/reference/src):
 *   positions [N,3], rotations [N,4] wxyz un-normalised, scales [N,3] log-space,
 *   opacities [N,1] logit-space, sh_coeffs [N,3,C] channel-major        (core/gaussian.hpp:24-40)
 *   means_2d [N,2], depths [N], cov_2d_inv [N,3] (a,b,c), radii [N] i32, tiles_touched [N] i32,
 *   rgb [N,3], opacities_act [N]                                  (rasterizer/projection.hpp:15-24)
 *   keys [P] u64 = tile_id<<32 | float_bits(depth), values [P] i32, tile_ranges [T,2] i32
 *                                                                    (rasterizer/sorting.hpp:19-24)
 *   color [H,W,3], final_T [H,W], n_contrib [H,W] i32                (rasterizer/forward.hpp:11-15)
 *   all f32 unless noted, contiguous, row-major, 16-byte aligned base pointers.
 *
 * Threading / streams: a handle belongs to one device and is NOT thread-safe (it owns the pinned
 * words through which the blocking calls return their counts -- a ring, one word per call, so frames
 * in flight on several streams and scans issued in between do not disturb each other -- the error
 * string and the optional stage-timing events); use one handle per host thread. Every kernel is launched on the stream passed in; the
 * only blocking calls are cugs_b200_render_plan and cugs_b200_scan with total_host != NULL (one
 * cudaStreamSynchronize each, the read the reference does at rasterizer/sorting.cu:146),
 * cugs_b200_densify_classify and cugs_b200_mcmc_relocate with counts_host != NULL (the counts the
 * reference reads with .item()), and cugs_b200_get_stage_ms. Two frames may be in flight on two streams through one handle as long as
 * each frame's render_finish is called before the next frame's render_plan (bench.py does this).
 * Limits: P < 2^30 pairs, N < 2^31 Gaussians, at most 48 K tiles (7680x4096) on the fused path
 * (CUGS_ERR_UNSUPPORTED otherwise); the stage functions have no tile limit.
 */
#ifndef CUGS_B200_H_
#define CUGS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUGS_B200_ABI_VERSION 7 /* bump on ANY signature change: callers compare it with cugs_b200_abi_version() */
#define CUGS_TILE 16 /* rasterizer/sorting.hpp:16 kTileSize */

enum {
    CUGS_OK = 0,
    CUGS_ERR_INVALID_ARG = -1,   /* null pointer, negative size, bad SH degree / coefficient count */
    CUGS_ERR_WORKSPACE = -2,     /* workspace too small (see *_workspace_bytes) */
    CUGS_ERR_CAPACITY = -3,      /* P exceeds the pair capacity the caller provided */
    CUGS_ERR_UNSUPPORTED = -4,   /* e.g. P >= 2^30 (look-back counters are 30-bit) */
    CUGS_ERR_NOT_BLACKWELL = -5  /* device is not sm_100 */
};

typedef struct cugs_handle cugs_handle_t;

/* One camera view + render settings. Replaces CameraInfo (core/types.hpp:78-109: width, height,
 * intrinsics, w2c rotation/translation) and RenderSettings (rasterizer/rasterizer.hpp:17-21).
 * view is the ROW-MAJOR 4x4 world->camera matrix the reference launchers build
 * (rasterizer/projection.cu:227-233); cam_center = -R^T t (core/types.hpp:98-100). */
typedef struct cugs_view {
    int32_t width, height;
    float fx, fy, cx, cy;
    float view[16];
    float cam_center[3];
    float bg[3];
    int32_t active_sh_degree; /* 0..3, already min(settings, model) as rasterizer.cpp:60 */
    int32_t num_coeffs;       /* C = allocated coefficients per channel, >= (deg+1)^2 */
    float scale_modifier;
} cugs_view_t;

/* ---- handle ------------------------------------------------------------------------------ */
int cugs_b200_create(int device, cugs_handle_t** out);
void cugs_b200_destroy(cugs_handle_t* h);
const char* cugs_b200_last_error(const cugs_handle_t* h);
int cugs_b200_abi_version(void);
int cugs_b200_sm_count(const cugs_handle_t* h);
/* SM count, maximum SM clock (kHz) and L2 size (bytes) of the handle's device, from the CUDA runtime: the
 * denominators of the FP32 / MUFU rooflines the benchmark reports for the blend kernels. */
int cugs_b200_device_info(const cugs_handle_t* h, int* sm_count, int* clock_khz, int* l2_bytes);
/* number of CUDA kernels this handle has launched so far (memsets / copies not counted; a replayed CUDA
 * graph of the trainer counts the kernel nodes it contains) */
uint64_t cugs_b200_launch_count(const cugs_handle_t* h);

/* ---- stage 1: preprocess forward ---------------------------------------------------------
 * Replaces project_gaussians (rasterizer/projection.cu:195-289): k_project_gaussians (:55-189),
 * the libtorch view-direction ops (:273-280), evaluate_sh_cuda (core/sh.cu:81-123) and
 * clamp_min(0) (:284), fused into one launch. Every output element is written (zeros where the
 * reference leaves its torch::zeros untouched), so outputs may be uninitialised.
 * packed (optional, may be NULL): [N,12] f32 private record {x,y,a,b | c,thr,op,r | g,b,hx,hy}
 * consumed by the blend kernels. depth_minmax (optional): 2 x u32 device words, pre-set by the
 * caller to {0xFFFFFFFF, 0}; receives min / max of float_bits(depth) over Gaussians with
 * tiles_touched > 0 (used to trim the sort's key bits). */
int cugs_b200_preprocess_fwd(cugs_handle_t* h, void* stream, int64_t n, const cugs_view_t* v,
                             const float* positions, const float* rotations, const float* scales,
                             const float* opacities, const float* sh_coeffs, float* means_2d,
                             float* depths, float* cov_2d_inv, int32_t* radii,
                             int32_t* tiles_touched, float* rgb, float* opacities_act,
                             float* packed, uint32_t* depth_minmax);

/* Stage functions evaluate_sh_cuda (core/sh.hpp:29) / evaluate_sh_backward_cuda
 * (core/sh_backward.hpp:25): directions [N,3] are given explicitly; rgb is NOT clamped. */
int cugs_b200_sh_forward(cugs_handle_t* h, void* stream, int64_t n, int degree, int num_coeffs,
                         const float* sh_coeffs, const float* directions, float* rgb);
int cugs_b200_sh_backward(cugs_handle_t* h, void* stream, int64_t n, int degree, int num_coeffs,
                          const float* sh_coeffs, const float* directions, const float* dL_drgb,
                          float* dL_dsh);

/* ---- stage 2: device-wide scan + duplicateWithKeys ---------------------------------------
 * Replaces cumsum/.item()/slice-copy (rasterizer/sorting.cu:145-152) and k_fill_sort_pairs
 * (:30-72). scan: offsets = exclusive int32 prefix sum; the total P is written to *total_dev
 * (device, int64, may be NULL) and, if total_host != NULL, copied to host with ONE stream
 * synchronisation (the same blocking read the reference does at sorting.cu:146).
 * scan_temp: cugs_b200_scan_temp_bytes(n) bytes of device scratch. */
size_t cugs_b200_scan_temp_bytes(int64_t n);
int cugs_b200_scan(cugs_handle_t* h, void* stream, int64_t n, const int32_t* tiles_touched,
                   int32_t* offsets, int64_t* total_dev, int64_t* total_host, void* scan_temp,
                   size_t scan_temp_bytes);
/* keys/values [P]: every slot in [0,P) is written, including the zero key/value filler of the
 * tile-count quirk (slots reserved by tiles_touched but not emitted). */
int cugs_b200_duplicate_with_keys(cugs_handle_t* h, void* stream, int64_t n, int width, int height,
                                  const float* means_2d, const float* depths, const int32_t* radii,
                                  const int32_t* tiles_touched, const int32_t* offsets, int64_t p,
                                  uint64_t* keys, int32_t* values);

/* ---- stage 3: onesweep radix sort ---------------------------------------------------------
 * Replaces cub::DeviceRadixSort::SortPairs (rasterizer/sorting.cu:190-211). Stable ascending
 * LSD sort of (u64 key, u32 value) pairs over the key bits in [0,depth_bits) U [32,32+tile_bits)
 * only (all other bits must be equal across keys, which holds for tile|depth keys when
 * depth_bits covers the highest differing depth bit and tile_bits = ceil(log2(num_tiles))).
 * depth_bits = 32, tile_bits = 32 sorts on all 64 bits like the reference call.
 * keys_in/values_in are clobbered (used as the ping-pong buffer). Result in keys_out/values_out. */
size_t cugs_b200_sort_temp_bytes(int64_t p);
int cugs_b200_sort_pairs(cugs_handle_t* h, void* stream, int64_t p, int depth_bits, int tile_bits,
                         uint64_t* keys_in, int32_t* values_in, uint64_t* keys_out,
                         int32_t* values_out, void* temp, size_t temp_bytes);

/* The packed sort of the fused render path, exposed as a stage function (tests, callers that build
 * their own binning). Stable ascending LSD sort of n packed 8-byte elements by the low key_bits bits of
 * their HIGH word (the low word is the payload, e.g. tile_id << 32 | gaussian index), 1-4 onesweep passes
 * of at most 8 bits (cugs_b200_sort_packed_passes). elts_a holds the input and is clobbered; the result
 * is in elts_b when the number of passes is odd, else in elts_a -- unless out32_last != NULL, in which
 * case the last pass writes only the low words there and the elements are not materialised.
 * tile_ranges (optional, needs num_tiles <= 2^key_bits): [num_tiles,2] i32 ranges [start,end) of each key
 * value in the sorted order, {0,0} for absent keys -- what k_compute_tile_ranges (rasterizer/sorting.cu:
 * 82-109) derives from the sorted keys, taken here from the key histogram of the same read.
 * n_dev (optional, device int64): the element count is min(*n_dev, n) read ON THE DEVICE and n is only the
 * capacity of the buffers; every launch is sized on n, so no host round trip is needed for the count. */
int cugs_b200_sort_packed_passes(int key_bits);
size_t cugs_b200_sort_packed_temp_bytes(int64_t n, int key_bits, int num_tiles);
int cugs_b200_sort_packed(cugs_handle_t* h, void* stream, int64_t n, int key_bits, uint64_t* elts_a,
                          uint64_t* elts_b, int32_t* out32_last, int num_tiles, int32_t* tile_ranges,
                          void* temp, size_t temp_bytes, const int64_t* n_dev);

/* ---- stage 4: tile ranges (rasterizer/sorting.cu:82-109); zero-fills tile_ranges first ----- */
int cugs_b200_tile_ranges(cugs_handle_t* h, void* stream, int64_t p, const uint64_t* keys_sorted,
                          int num_tiles, int32_t* tile_ranges);

/* ---- stage 5: blend forward (rasterize_forward, rasterizer/forward.cu:180-240, :48-174) ----
 * packed may be NULL, in which case the records are gathered from the four public arrays. */
int cugs_b200_blend_fwd(cugs_handle_t* h, void* stream, const cugs_view_t* v,
                        const int32_t* tile_ranges, const int32_t* gaussian_idx,
                        const float* means_2d, const float* cov_2d_inv, const float* rgb,
                        const float* opacities_act, const float* packed, float* color,
                        float* final_T, int32_t* n_contrib);

/* ---- stage 6: backward --------------------------------------------------------------------
 * blend_bwd replaces rasterize_backward (rasterizer/backward.cu:239-306, :31-233). The four
 * outputs are zero-filled by this call and then accumulated (warp-reduced, one vector
 * reduction per warp and Gaussian). grad_acc: [N,12] f32 scratch. */
int cugs_b200_blend_bwd(cugs_handle_t* h, void* stream, int64_t n, const cugs_view_t* v,
                        const int32_t* tile_ranges, const int32_t* gaussian_idx,
                        const float* means_2d, const float* cov_2d_inv, const float* rgb,
                        const float* opacities_act, const float* packed, const float* dL_dcolor,
                        const float* final_T, const int32_t* n_contrib, float* dL_drgb,
                        float* dL_dopacity_act, float* dL_dmeans_2d, float* dL_dcov_2d_inv,
                        float* grad_acc);
/* Measurement only (never on a hot path): counts the (pixel, Gaussian) evaluations of the REFERENCE's
 * traversal of a frame -- counts4_dev (device, 4 x u64) = {forward alpha-rejected, forward contributing,
 * backward alpha-rejected, backward contributing}, the work units E_fwd / E_bwd of the blend kernels'
 * FP32 / MUFU roofline (rasterizer/forward.cu:121-157, rasterizer/backward.cu:117-145). */
int cugs_b200_count_evaluations(cugs_handle_t* h, void* stream, const cugs_view_t* v,
                                const int32_t* tile_ranges, const int32_t* gaussian_idx,
                                const float* means_2d, const float* cov_2d_inv,
                                const float* opacities_act, const int32_t* n_contrib,
                                uint64_t* counts4_dev);
/* preprocess_bwd replaces project_backward (rasterizer/projection_backward.cu:253-344):
 * k_project_backward (:26-247) + directions + evaluate_sh_backward_cuda (core/sh_backward.cu),
 * one launch. rgb = forward output (its sign is the ReLU gate). All outputs fully written.
 * stats (optional, all three or none): densification accumulators updated in place as
 * DensificationController::accumulate_gradients does (optimizer/densification.cpp:59-88). */
int cugs_b200_preprocess_bwd(cugs_handle_t* h, void* stream, int64_t n, const cugs_view_t* v,
                             const float* positions, const float* rotations, const float* scales,
                             const float* opacities, const float* sh_coeffs, const int32_t* radii,
                             const float* rgb, const float* dL_dmeans_2d,
                             const float* dL_dcov_2d_inv, const float* dL_drgb,
                             const float* dL_dopacity_act, float* dL_dpositions,
                             float* dL_drotations, float* dL_dscales, float* dL_dopacities,
                             float* dL_dsh_coeffs, float* grad_accum, float* grad_count,
                             float* max_radii);

/* ---- fused entry points: cugs::render / cugs::render_backward ------------------------------
 * render_plan = preprocess + scan, returns P on the host (one sync, as the reference).
 * render_finish = duplicateWithKeys + sort + tile ranges + blend; gaussian_idx must hold P
 * entries. workspace: cugs_b200_render_workspace_bytes(n, p_capacity) bytes; the SAME workspace
 * must be passed to plan, finish and render_backward of one frame (it carries the packed
 * records). render_backward = blend_bwd + preprocess_bwd (rasterizer/rasterizer.cpp:115-186).
 * flags = 0: the five parameter gradients are overwritten (reference behaviour);
 * flags & CUGS_BWD_ACCUMULATE: they are added to what the buffers hold (view-batched training: the
 * gradient of a batch of views is summed in place, no separate axpy pass).
 * flags & CUGS_BWD_SPARSE_ROWS (needs touch_mask): the caller guarantees that every gradient row
 * whose touch_mask entry is 0 on entry is all zero (buffers allocated with zeros and only ever
 * written by this call / cugs_b200_scatter_grad_rows with the matching mask); rows that are not
 * touched are then neither read nor written (overwrite mode still zeroes rows touched before).
 * dL_dmeans_2d is always
 * overwritten (it is a per-view quantity, optimizer/densification.cpp:77).
 * touch_mask (optional, [N] i32): 1 where this view gave the Gaussian a non-zero 2-D gradient, else
 * 0 (OR-ed into the previous content when accumulate = 1). Rows with mask 0 have all-zero parameter
 * gradients; the view-parallel gradient exchange below only moves the other rows.
 * Render-only frames (evaluation / viewer callers, training/metrics.cpp:131, viewer/viewer.cpp:645-669,
 * which consume color / final_T / n_contrib only): pass NULL for ALL FOUR of depths, cov_2d_inv, rgb
 * and opacities_act to render_plan and render_finish; those backward-only arrays are then not
 * written (the blend reads the packed records of the workspace). Such a frame cannot be passed to
 * render_backward. */
size_t cugs_b200_render_workspace_bytes(int64_t n, int64_t p_capacity);
/* The workspace has an N-sized head (lives from render_plan to render_backward: cugs_b200_render_workspace_bytes(n, 0))
 * and a P-sized tail that is only used inside render_finish. A caller that sizes its allocations per frame (the
 * libtorch wrapper) passes the tail as a SEPARATE block pair_scratch of cugs_b200_render_pair_scratch_bytes(p)
 * bytes, so that learning P does not force it to re-allocate and copy the head; pair_scratch = NULL means the
 * tail follows the head inside `workspace` (cugs_b200_render_workspace_bytes(n, p) bytes). */
size_t cugs_b200_render_pair_scratch_bytes(int64_t p_capacity);
/* render_forward = render_plan + render_finish WITHOUT the host round trip for P: gaussian_idx and the
 * workspace are sized for p_capacity pairs, every launch is sized on that capacity and the kernels read
 * the frame's pair count on the device. Nothing blocks, so a whole training step can be queued ahead or
 * captured in a CUDA graph (the reference blocks once per frame, rasterizer/sorting.cu:146).
 * status_dev (optional, DEVICE memory, 2 x int64): receives {P, P > p_capacity}. An overflowed frame is
 * safe (pairs beyond the capacity are dropped, no kernel runs past a buffer) but its image and gradients
 * are incomplete: the caller reads the status whenever convenient (e.g. once per step), and re-runs the
 * frame with a larger capacity. Entries of gaussian_idx beyond P are not written. */
int cugs_b200_render_forward(cugs_handle_t* h, void* stream, int64_t n, int64_t p_capacity,
                             const cugs_view_t* v, const float* positions, const float* rotations,
                             const float* scales, const float* opacities, const float* sh_coeffs,
                             float* means_2d, float* depths, float* cov_2d_inv, int32_t* radii,
                             float* rgb, float* opacities_act, int32_t* gaussian_idx,
                             int32_t* tile_ranges, float* color, float* final_T, int32_t* n_contrib,
                             void* workspace, size_t workspace_bytes, void* pair_scratch,
                             size_t pair_scratch_bytes, int64_t* status_dev);
int cugs_b200_render_plan(cugs_handle_t* h, void* stream, int64_t n, const cugs_view_t* v,
                          const float* positions, const float* rotations, const float* scales,
                          const float* opacities, const float* sh_coeffs, float* means_2d,
                          float* depths, float* cov_2d_inv, int32_t* radii, float* rgb,
                          float* opacities_act, void* workspace, size_t workspace_bytes,
                          int64_t* p_host);
int cugs_b200_render_finish(cugs_handle_t* h, void* stream, int64_t n, int64_t p,
                            const cugs_view_t* v, const float* means_2d, const float* depths,
                            const float* cov_2d_inv, const int32_t* radii, const float* rgb,
                            const float* opacities_act, int32_t* gaussian_idx, int32_t* tile_ranges,
                            float* color, float* final_T, int32_t* n_contrib, void* workspace,
                            size_t workspace_bytes, void* pair_scratch, size_t pair_scratch_bytes);
int cugs_b200_render_backward(cugs_handle_t* h, void* stream, int64_t n, const cugs_view_t* v,
                              const float* positions, const float* rotations, const float* scales,
                              const float* opacities, const float* sh_coeffs,
                              const float* means_2d, const float* cov_2d_inv, const int32_t* radii,
                              const float* rgb, const float* opacities_act,
                              const int32_t* gaussian_idx, const int32_t* tile_ranges,
                              const float* final_T, const int32_t* n_contrib,
                              const float* dL_dcolor, float* dL_dpositions, float* dL_drotations,
                              float* dL_dscales, float* dL_dopacities, float* dL_dsh_coeffs,
                              float* dL_dmeans_2d, float* grad_accum, float* grad_count,
                              float* max_radii, int32_t* touch_mask, int flags, void* workspace,
                              size_t workspace_bytes);

/* Optional per-stage device timing of the three fused entry points above: when enabled, CUDA
 * events are recorded on the caller's stream at the stage boundaries (a few microseconds per
 * frame); get_stage_ms synchronises on the last recorded event and returns the durations in ms
 * of {preprocess_fwd, scan, duplicate_with_keys, sort, tile_ranges, blend_fwd, blend_bwd,
 * preprocess_bwd} of the most recent frame (-1 for a stage that did not run). */
/* Sort plan of the most recent render_finish: number of onesweep passes and sorted key bits. */
int cugs_b200_last_sort_plan(const cugs_handle_t* h, int* passes, int* key_bits);
#define CUGS_BWD_ACCUMULATE 1
#define CUGS_BWD_SPARSE_ROWS 2
/* With SPARSE_ROWS the call can be split in two: STOP_AFTER_MASK runs the backward blend and the classification
 * pass -- after it touch_mask, dL_dmeans_2d and the three statistics are final -- and RESUME_AFTER_MASK (same
 * arguments, same workspace) runs the rest (the chain rule on the touched Gaussians). View-parallel training
 * starts the MAX all-reduce of the mask between the two, so it is hidden under the second part. */
#define CUGS_BWD_STOP_AFTER_MASK 4
#define CUGS_BWD_RESUME_AFTER_MASK 8
#define CUGS_NUM_STAGES 8
int cugs_b200_set_stage_timing(cugs_handle_t* h, int enable);
int cugs_b200_get_stage_ms(cugs_handle_t* h, float* ms8);

/* ---- stage 7: fused L1 + SSIM loss and fused multi-tensor Adam ------------------------------
 * loss: combined_loss (training/loss.cpp:131-135) value AND its gradient w.r.t. rendered (the
 * reference gets the gradient from libtorch autograd, training/trainer.cpp:214-217).
 * scalars3 (device) = {loss, l1, mean ssim}. workspace: cugs_b200_loss_workspace_bytes(w,h).
 * ssim_map (optional, may be NULL): [H,W] per-pixel SSIM averaged over the three channels, what
 * cugs::ssim returns (training/loss.cpp:88-124; consumed by training/metrics.cpp:41-46).
 * window_size: side of the gaussian window (sigma 1.5), odd, 3..33 (loss.hpp:33-44; loss.cpp:91-92 reject
 * even sizes and sizes < 3): 11 = the default every training caller uses (tuned kernels); any other size
 * runs generic fall-back kernels with the same formulation. */
size_t cugs_b200_loss_workspace_bytes(int width, int height);
int cugs_b200_loss_l1_ssim(cugs_handle_t* h, void* stream, int width, int height, float lambda,
                           int window_size, const float* rendered, const float* target, float* dL_dcolor,
                           float* scalars3, void* workspace, size_t workspace_bytes, float* ssim_map);
/* Adam: FusedAdam::step (optimizer/fused_adam.cu:140-164) + k_fused_adam (:44-76) for all five
 * groups in ONE launch. Group order positions, sh_coeffs, opacities, scales, rotations
 * (fused_adam.cu:94-97); counts[g] = number of floats. bc1/bc2 are computed by the caller in
 * double exactly as fused_adam.cu:145-149. grad_scale multiplies every gradient first
 * (1.0 = reference behaviour; 1/views for view-batched training). */
int cugs_b200_adam_step(cugs_handle_t* h, void* stream, float* const params[5],
                        const float* const grads[5], float* const m[5], float* const v[5],
                        const int64_t counts[5], const float lr[5], float beta1, float beta2,
                        float eps, float bc1, float bc2, float grad_scale);
/* MCMC per-step operations (optimizer/mcmc_densification.cpp; training/trainer.cpp:232-237, :251).
 * adam_step_mcmc = adam_step with the closed-form gradient of MCMCController::compute_regularization
 * (:167-186: lambda_opacity * mean(sigmoid(opacity)) + lambda_scale * mean(exp(scale))) added to the
 * opacity / scale gradients inside the same launch (after grad_scale).
 * mcmc_inject_noise = MCMCController::inject_noise (:144-161), in place on positions, using the
 * (already updated) scales and opacities; noise_lr = MCMCController::noise_lr(step) computed by the
 * caller; normals are Philox-4x32-10(seed; index, step) so that replicated ranks draw identical noise;
 * normals_out (optional [N,3]) returns the N(0,1) draws for testing. */
int cugs_b200_adam_step_mcmc(cugs_handle_t* h, void* stream, float* const params[5],
                             const float* const grads[5], float* const m[5], float* const v[5],
                             const int64_t counts[5], const float lr[5], float beta1, float beta2,
                             float eps, float bc1, float bc2, float grad_scale, float lambda_opacity,
                             float lambda_scale);
int cugs_b200_mcmc_inject_noise(cugs_handle_t* h, void* stream, int64_t n, float* positions,
                                const float* scales, const float* opacities, float noise_lr, float gate_k,
                                float gate_t, uint64_t seed, uint32_t step, float* normals_out);
/* DensificationController::accumulate_gradients (optimizer/densification.cpp:59-88). */
int cugs_b200_accumulate_stats(cugs_handle_t* h, void* stream, int64_t n,
                               const float* dL_dmeans_2d, const int32_t* radii, float* grad_accum,
                               float* grad_count, float* max_radii);

/* ---- model-resizing steps on a schedule ---------------------------------------
 * ADC densification, DensificationController::densify (optimizer/densification.cpp:94-329), as two
 * calls around the host policy:
 *   densify_classify  one pass over scales / opacities / the three accumulators -> flags[N] (u8, OR of
 *                     CUGS_DENSIFY_*: the clone mask :351-370, the split mask :372-399, and KEEP = passes
 *                     compute_keep_mask :401-444; a row with SPLIT is dropped whatever its KEEP bit,
 *                     :296-303) and the three counts {kept originals, clones, splits} on the host (one sync; the reference has
 *                     three .item() reads). The caller may clear CLONE / SPLIT bits afterwards (budget caps,
 *                     :122-139) before calling apply.
 *   densify_apply     stable stream compaction of the five parameter arrays (Adam group order: positions,
 *                     sh_coeffs, opacities, scales, rotations) into dst, n_out = kept + clones + 2 * splits
 *                     rows in the reference's order [kept originals | clones | first children | second
 *                     children]; children: scale - log(1.6), position + N(0,1) * exp(new scale) (:243-253),
 *                     normals = Philox-4x32-10(seed; source row, child), optionally returned in
 *                     split_normals_out [2 * splits, 3]. src_m/src_v/dst_m/dst_v (each NULL or five
 *                     pointers): Adam moments carried over for kept rows and zeroed for new rows (the
 *                     reference rebuilds the optimizer, i.e. zeroes all of them: pass src_* = NULL for that).
 * densify_apply blocks once at its end to compare n_out with the row counts it derived from the flags
 * (CUGS_ERR_INVALID_ARG on a mismatch; no row beyond n_out is ever written).
 * temp: cugs_b200_densify_temp_bytes(n) bytes of device scratch, the same buffer for both calls. */
#define CUGS_DENSIFY_KEEP 1
#define CUGS_DENSIFY_CLONE 2
#define CUGS_DENSIFY_SPLIT 4
typedef struct cugs_densify_config {
    float grad_threshold;     /* DensificationConfig::grad_threshold, densification.hpp:31 */
    float size_threshold;     /* percent_dense * scene_extent, densification.cpp:366 */
    float opacity_threshold;  /* densification.hpp:32 */
    int32_t apply_size_pruning; /* opacity_reset_every > 0 && step > opacity_reset_every, densification.cpp:416-417 */
    float max_screen_size;    /* (float)max_screen_size; <= 0 disables the screen-size test, :421 */
    float ws_threshold;       /* 0.1 * scene_extent, :438 */
} cugs_densify_config_t;
size_t cugs_b200_densify_temp_bytes(int64_t n);
int cugs_b200_densify_classify(cugs_handle_t* h, void* stream, int64_t n, const float* scales,
                               const float* opacities, const float* grad_accum, const float* grad_count,
                               const float* max_radii, const cugs_densify_config_t* cfg, uint8_t* flags,
                               int64_t* counts_host, void* temp, size_t temp_bytes);
int cugs_b200_densify_apply(cugs_handle_t* h, void* stream, int64_t n, int64_t n_out, int num_coeffs,
                            const uint8_t* flags, const float* const* src, float* const* dst,
                            const float* const* src_m, const float* const* src_v, float* const* dst_m,
                            float* const* dst_v, uint64_t seed, float* split_normals_out, void* temp,
                            size_t temp_bytes);
/* MCMCController::relocate (optimizer/mcmc_densification.cpp:56-138), in place: the first max_relocate
 * dead Gaussians (sigmoid(opacity) < dead_threshold, in index order) each take over an alive Gaussian
 * drawn with probability proportional to sigmoid(opacity) (with replacement): SH and rotation copied,
 * position + N(0,1) * scene_extent * 0.01, scale - log(10), opacity = logit(0.01). Nothing happens when
 * there is no dead or no alive Gaussian. Draws are Philox-4x32-10(seed; index, step): identical on every
 * rank of a view-parallel run. source_out (optional [N] i32): chosen source per Gaussian, -1 = untouched;
 * normals_out (optional [N,3]); counts_host (optional): {dead, relocated} -- non-NULL makes the call
 * blocking, as the reference's .item() is. temp: cugs_b200_mcmc_relocate_temp_bytes(n). */
size_t cugs_b200_mcmc_relocate_temp_bytes(int64_t n);
int cugs_b200_mcmc_relocate(cugs_handle_t* h, void* stream, int64_t n, int num_coeffs, float* positions,
                            float* sh_coeffs, float* opacities, float* scales, float* rotations,
                            float dead_threshold, int64_t max_relocate, float scene_extent, uint64_t seed,
                            uint32_t step, int32_t* source_out, float* normals_out, int64_t* counts_host,
                            void* temp, size_t temp_bytes);

/* ---- the training step in C++ ---------------------------------------------------
 * Trainer::train_step (training/trainer.cpp:178-316) without the Dataset, as host code of this library:
 * update_lr (:180) -> active SH degree (:183, training/lr_schedule.hpp:70-72) -> per view: render (:211) ->
 * fused L1+SSIM loss and gradient (:214-225) -> render_backward (:228) with the step's gradients summed in
 * place and accumulate_gradients (:269) fused in -> FusedAdam::step (:240-242) [+ the MCMC regulariser
 * :232-237 inside the Adam launch, + inject_noise :251]. No host synchronisation inside a step (the
 * reference blocks 3-6 times per iteration): frames go through cugs_b200_render_forward, losses and the
 * per-frame pair counts stay on the device and reach pinned memory at the end of the step. With use_graph the
 * step is captured into a CUDA graph on its second call and replayed afterwards (re-captured when the view
 * set or the active SH degree changes); per-step scalars (learning rates, bias corrections, noise scale,
 * step number) are refreshed by one small H2D copy in front of each replay. If a frame's pair count exceeds
 * p_capacity the update phase does nothing (status ok = 0): re-create the trainer with a larger capacity and
 * repeat the step. All tensors are caller-owned device memory in the reference's layouts; params / adam_m /
 * adam_v / grads are in the Adam group order positions, sh_coeffs, opacities, scales, rotations
 * (optimizer/fused_adam.cu:94-97). */
typedef struct cugs_train_config {
    float lambda_ssim;             /* TrainConfig::lambda_ssim, weight of the SSIM term (loss.hpp:52) */
    int32_t max_sh_degree;
    float background[3];
    float lr_position_init, lr_position_final; /* PositionLRConfig, training/lr_schedule.hpp:36-41 */
    int32_t lr_position_max_steps;
    float lr_sh_coeffs, lr_opacities, lr_scales, lr_rotations; /* AdamConfig, optimizer/adam.hpp:30-41 */
    float beta1, beta2, eps;
    int32_t accumulate_stats;      /* fuse DensificationController::accumulate_gradients (needs the stats pointers) */
    int32_t mcmc;                  /* MCMC mode: regulariser gradient inside Adam + position noise after it */
    float lambda_opacity, lambda_scale;
    float noise_lr_init, noise_lr_final;
    int32_t noise_lr_max_steps;
    float noise_gate_k, noise_gate_t;
    uint64_t noise_seed;
    int32_t frames_in_flight;      /* 1 or 2 (two streams: the front end of view v+1 under the backward of view v) */
    int32_t use_graph;             /* capture the step into a CUDA graph and replay it */
} cugs_train_config_t;

typedef struct cugs_train_tensors {
    float* params[5];
    float* adam_m[5];
    float* adam_v[5];
    float* grads[5];
    float* dL_dmeans_2d;           /* [N,2], per view (densification.cpp:77) */
    float* grad_accum;             /* optional, all three or none: the densification accumulators */
    float* grad_count;
    float* max_radii;
    int32_t* touch_mask;           /* optional [N] i32: sparse gradient rows (see CUGS_BWD_SPARSE_ROWS) */
} cugs_train_tensors_t;

typedef struct cugs_trainer cugs_trainer_t;
size_t cugs_b200_trainer_workspace_bytes(int64_t n, int num_coeffs, int width, int height,
                                         int64_t p_capacity, int frames_in_flight);
int cugs_b200_trainer_create(cugs_handle_t* h, int64_t n, int num_coeffs, int width, int height,
                             int64_t p_capacity, const cugs_train_config_t* cfg,
                             const cugs_train_tensors_t* tensors, void* workspace, size_t workspace_bytes,
                             cugs_trainer_t** out);
void cugs_b200_trainer_destroy(cugs_trainer_t* t);
/* The views this rank renders every step (cameras + device target images [H,W,3]); total_views_per_step =
 * views of ALL ranks (Adam's gradient scale is 1 / total). active_sh_degree, num_coeffs, bg and
 * scale_modifier of the views are set by the trainer. dL_dcolor_dev (optional array, entries may be NULL):
 * a view with a given dL/dcolor [H,W,3] skips the loss (forward + backward only; its loss scalars are 0).
 * targets_host_pinned (optional array, entries may be NULL): the view's target image lives in PINNED host memory
 * and is copied into targets_dev[v] inside every step, on a copy stream, under the rendering of that view
 * (the reference uploads the image synchronously every iteration, training/trainer.cpp:186-198). Such views need
 * distinct device buffers. */
int cugs_b200_trainer_set_views(cugs_trainer_t* t, int num_views, const cugs_view_t* views,
                                const float* const* targets_dev, const float* const* dL_dcolor_dev,
                                const float* const* targets_host_pinned, int total_views_per_step);
/* phases: 1 = views (render -> loss -> backward), 2 = update (Adam [+ regulariser] [+ noise]), 3 = both.
 * View-parallel training calls 1, exchanges the gradients, then 2. Nothing blocks. */
int cugs_b200_trainer_step(cugs_trainer_t* t, void* stream, int step, int phases);
/* Blocking read of the most recent step: scalars3 = {loss, l1, mean ssim} averaged over this rank's views,
 * status3 = {ok, largest pair count of the step's frames, views folded in}. */
int cugs_b200_trainer_result(cugs_trainer_t* t, void* stream, float scalars3[3], int64_t status3[3]);
/* FusedAdam::step_count_ (bias corrections): carried over when a trainer is re-created */
int cugs_b200_trainer_set_adam_steps(cugs_trainer_t* t, int64_t steps);
int64_t cugs_b200_trainer_adam_steps(const cugs_trainer_t* t);

/* ---- view-parallel gradient exchange (no reference counterpart: the reference is single-GPU) ----
 * Compaction of the gradient rows of the touched Gaussians around the all-reduce. touch: [N] i32
 * union mask (after a MAX all-reduce over the ranks); offsets: its exclusive scan (cugs_b200_scan);
 * m: number of touched Gaussians; grads: the five dense gradient arrays in Adam group order
 * (positions [N,3], sh_coeffs [N,3,C], opacities [N,1], scales [N,3], rotations [N,4]); compact:
 * cugs_b200_compact_grad_floats(m, C) floats, group-major, every group block starting at a multiple of
 * 4 floats; idx_scratch: m ints of device scratch (index list).
 * m_dev (optional, device int64 = the total the scan wrote): no host round trip for M -- `m` is then only the
 * row CAPACITY the compact layout is sized for (e.g. 1.25 x the previous step's M), the real count is read on
 * the device, the gather zero-fills the rows in between and status_dev (optional, 2 x int64) receives
 * {M, M > capacity}; rows beyond the capacity are dropped, so an overflowed exchange must be repeated.
 * scatter with touch = offsets = NULL reuses the index list the gather of the same exchange left in idx_scratch. */
int64_t cugs_b200_compact_grad_floats(int64_t m, int num_coeffs);
int cugs_b200_gather_grad_rows(cugs_handle_t* h, void* stream, int64_t n, int num_coeffs,
                               const int32_t* touch, const int32_t* offsets, int64_t m,
                               const float* const grads[5], float* compact, int32_t* idx_scratch,
                               const int64_t* m_dev, int64_t* status_dev);
int cugs_b200_scatter_grad_rows(cugs_handle_t* h, void* stream, int64_t n, int num_coeffs,
                                const int32_t* touch, const int32_t* offsets, int64_t m,
                                const float* compact, float* const grads[5], int32_t* idx_scratch,
                                const int64_t* m_dev);

/* Peer-to-peer variant of the exchange (csrc/grad_exchange.cu): every rank's gradient arena, [touch mask |
 * max_radii] buffer and statistics live in symmetric memory and each rank holds peer-mapped pointers to all of
 * them (`*_peers[world]`, rank order, own buffers included). p2p_reduce_masks: MAX of the 2 N int32 of max_buf and
 * SUM of the two statistics (NULL, NULL to skip them), each rank reducing its slice of [0, n) and writing the
 * result into every peer. p2p_reduce_rows: the five gradient rows of every Gaussian of the union mask
 * (idx = cugs_b200_build_touch_index of the reduced mask's scan, m_dev = that scan's device-side total) are summed
 * across the ranks in rank order and written back into every arena; grads_peers[p * 5 + k] = group k
 * (positions, sh_coeffs, opacities, scales, rotations) of rank p. No compact buffers, no host knowledge of M, the
 * same bits on every rank. The CALLER synchronises the ranks (symmetric-memory barrier) before the first call,
 * between the two, and after the second.
 * *_mc (optional): the MULTICAST mapping of the same symmetric buffers (NVSwitch / NVLS). When given, the kernels
 * use multimem.ld_reduce (the switch reduces the element over all GPUs) and multimem.st (the switch replicates the
 * store) instead of R unicast loads and R unicast stores -- about half the bytes per NVLink direction; the
 * summation order inside the switch is unspecified (every rank still receives the same bits). */
int cugs_b200_build_touch_index(cugs_handle_t* h, void* stream, int64_t n, const int32_t* touch,
                                const int32_t* offsets, int32_t* idx);
int cugs_b200_p2p_reduce_masks(cugs_handle_t* h, void* stream, int64_t n, int world, int rank,
                               int32_t* const* max_buf_peers, float* const* grad_accum_peers,
                               float* const* grad_count_peers, int32_t* max_buf_mc, float* grad_accum_mc,
                               float* grad_count_mc);
int cugs_b200_p2p_reduce_rows(cugs_handle_t* h, void* stream, int64_t n, int num_coeffs, int world, int rank,
                              const int32_t* idx, const int64_t* m_dev, float* const* grads_peers,
                              float* const* grads_mc);

#ifdef __cplusplus
}
#endif
#endif /* CUGS_B200_H_ */
