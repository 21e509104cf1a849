// trainer.cu — the training step in C++ on the C ABI.
//
// Replaces the body of Trainer::train_step (reference training/trainer.cpp:178-316) without the Dataset:
//   update_lr (:180) -> active SH degree (:183) -> for each view of the step: render (:211) -> L1+SSIM
//   loss and its gradient (:214-225, one fused pass instead of libtorch autograd) -> render_backward
//   (:228; gradients of the step's views summed in place, accumulate_gradients :269 fused into the same
//   launch) -> [MCMC: regulariser gradient :232-237 inside the Adam launch] -> FusedAdam::step (:240-242,
//   ONE multi-tensor launch) -> [MCMC: inject_noise :251].
//
// What is different from the reference's host code, by design:
//  * NO host synchronisation anywhere in a step. The reference blocks 3-6 times per iteration (.item() for
//    the pair count, rasterizer/sorting.cu:146; three loss .item()s, trainer.cpp:223-224, :308). Here every
//    frame runs through cugs_b200_render_forward (pair count read on the device, launches sized on a
//    capacity), the loss scalars and the per-frame {P, overflow} words stay on the device and are copied
//    to pinned memory at the end of the step; the caller looks at them whenever it wants.
//  * The whole step is captured ONCE into a CUDA graph and replayed: ~60 launches and memsets per view
//    become one cudaGraphLaunch. The scalars that change every iteration (learning rates, Adam bias
//    corrections, noise scale, step number) live in a small device struct (StepDyn, common.cuh) that the
//    host refreshes with one 48-byte H2D copy in front of each replay. The graph is re-captured only when
//    the view set, the active SH degree or a buffer changes.
//  * The step is transactional with respect to capacity: if a frame's pair count exceeds the capacity the
//    buffers were sized for, dyn.ok = 0 and Adam / noise do nothing; the caller re-creates the trainer
//    with a larger capacity and repeats the step.
//  * Two frames in flight (main + auxiliary stream, fork / join by events, also inside the graph): the
//    memory-bound front end of view v+1 overlaps the issue-bound backward blend of view v.
// View-parallel training calls the two phases separately (views | exchange by the caller | update).
#include "common.cuh"

#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

using namespace cugs;

int cugs_adam_launch(cugs_handle_t* h, void* stream, float* const params[5], const float* const grads[5],
                     float* const m[5], float* const v[5], const int64_t counts[5], const float lr[5], float beta1,
                     float beta2, float eps, float bc1, float bc2, float grad_scale, float lambda_opacity,
                     float lambda_scale, const StepDyn* dyn);
int cugs_noise_launch(cugs_handle_t* h, void* stream, int64_t n, float* positions, const float* scales,
                      const float* opacities, float noise_lr, float gate_k, float gate_t, uint64_t seed,
                      uint32_t step, float* normals_out, const StepDyn* dyn);

namespace {

constexpr int kMaxFrames = 2;
constexpr int kDynRing = 64;  // pinned staging slots for StepDyn: the host may run this many steps ahead

struct FrameBuf {
    void* ws; size_t ws_bytes;
    void* pairs; size_t pairs_bytes;
    float *means_2d, *depths, *cov, *rgb, *opa, *color, *final_T, *dL;
    int32_t *radii, *gidx, *ranges, *n_contrib;
    void* loss_ws; size_t loss_bytes;
    float* scalars3;
    int64_t* status2;
};

// step result block (device) and its pinned mirror: {loss, l1, ssim} as floats, then counters
struct StepResult {
    float scalars[4];        // loss, l1, ssim (mean over the rank's views), unused
    int64_t ok;              // 1 unless a frame overflowed its pair capacity
    int64_t max_pairs;       // largest P of the step's frames
    int64_t views;           // frames folded in
    int64_t adam_steps;      // FusedAdam::step_count_ (fused_adam.cu:141), advanced ON THE DEVICE only by steps that ran
};

}  // namespace

struct cugs_trainer {
    cugs_handle_t* h;
    int64_t n, p_cap;
    int C, W, H, frames;
    cugs_train_config_t cfg;
    cugs_train_tensors_t t;
    FrameBuf f[kMaxFrames];
    StepDyn* dyn_dev;
    StepDyn* dyn_pinned;       // [kDynRing]
    StepResult* res_dev;
    StepResult* res_pinned;
    uint64_t dyn_seq;
    cudaEvent_t ring_ev[kDynRing];
    bool ring_ev_used[kDynRing];
    cudaStream_t aux;  // second frame in flight
    cudaStream_t cap;  // capture origin: the caller's stream may be the legacy default stream, which cannot capture
    cudaStream_t copy; // host -> device copies of the target images, under the rendering of the same view
    cudaEvent_t ev_copy_fork, ev_copied[2], ev_copy_join;
    cudaEvent_t ev_fork, ev_bwd[2], ev_join;
    // the step's views
    std::vector<cugs_view_t> views;
    std::vector<const float*> targets;
    std::vector<const float*> dLs;  // optional per view: a given dL/dcolor replaces the loss (forward+backward only)
    std::vector<const float*> targets_host;  // optional per view: pinned host image copied into targets[v] inside the step
    int total_views;
    // graph cache: one executable per phase mask (1, 2, 3), valid for (views generation, degree)
    cudaGraphExec_t exec[16];
    int exec_degree[16];
    uint64_t exec_gen[16];
    uint64_t views_gen;
    bool warmed[16];
    unsigned long long graph_kernels[16];  // kernel nodes of each captured graph (for the handle's launch counter)
};

namespace {

__global__ void k_fold_frame(const float* __restrict__ scalars3, const int64_t* __restrict__ status2,
                             StepResult* __restrict__ res) {
    // one thread: frames of a step fold in stream order (the backward chain serialises them)
    res->scalars[0] += scalars3[0];
    res->scalars[1] += scalars3[1];
    res->scalars[2] += scalars3[2];
    if (status2[1] != 0) res->ok = 0;
    if (status2[0] > res->max_pairs) res->max_pairs = status2[0];
    res->views += 1;
}

__global__ void k_begin_step(StepResult* __restrict__ res) {
    res->scalars[0] = res->scalars[1] = res->scalars[2] = res->scalars[3] = 0.0f;
    res->ok = 1;
    res->max_pairs = 0;
    res->views = 0;
}

__global__ void k_end_views(StepResult* __restrict__ res, StepDyn* __restrict__ dyn) {
    const float inv = res->views > 0 ? 1.0f / (float)res->views : 0.0f;
    res->scalars[0] *= inv;
    res->scalars[1] *= inv;
    res->scalars[2] *= inv;
    dyn->ok = (int)res->ok;
}

// FusedAdam::step's host prologue (fused_adam.cu:141-149) on the device: the step counter advances and the bias
// corrections are computed (in double, as the reference) only if the step is going to run, so a skipped
// (overflowed) step leaves no trace in the optimizer state.
__global__ void k_prepare_update(StepResult* __restrict__ res, StepDyn* __restrict__ dyn, double beta1, double beta2) {
    if (!dyn->ok) return;
    res->adam_steps += 1;
    const double k = (double)res->adam_steps;
    dyn->bc1 = (float)(1.0 / (1.0 - pow(beta1, k)));
    dyn->bc2 = (float)(1.0 / (1.0 - pow(beta2, k)));
}

size_t a256(size_t x) { return align_up(x ? x : 16, 256); }

size_t frame_bytes(int64_t n, int W, int H, int64_t p_cap, FrameBuf* f, char* base) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? base + off : nullptr;
        off += a256(bytes);
        return p;
    };
    const size_t nn = (size_t)n, px = (size_t)W * H;
    const int tiles = ((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
    FrameBuf tmp;
    FrameBuf& o = f ? *f : tmp;
    o.ws_bytes = cugs_b200_render_workspace_bytes(n, 0);
    o.ws = take(o.ws_bytes);
    o.pairs_bytes = cugs_b200_render_pair_scratch_bytes(p_cap);
    o.pairs = take(o.pairs_bytes);
    o.means_2d = (float*)take(nn * 8);
    o.depths = (float*)take(nn * 4);
    o.cov = (float*)take(nn * 12);
    o.rgb = (float*)take(nn * 12);
    o.opa = (float*)take(nn * 4);
    o.radii = (int32_t*)take(nn * 4);
    o.gidx = (int32_t*)take((size_t)(p_cap > 0 ? p_cap : 1) * 4);
    o.ranges = (int32_t*)take((size_t)tiles * 8);
    o.color = (float*)take(px * 12);
    o.final_T = (float*)take(px * 4);
    o.n_contrib = (int32_t*)take(px * 4);
    o.dL = (float*)take(px * 12);
    o.loss_bytes = cugs_b200_loss_workspace_bytes(W, H);
    o.loss_ws = take(o.loss_bytes);
    o.scalars3 = (float*)take(16);
    o.status2 = (int64_t*)take(16);
    return off;
}

// float schedules exactly as the reference computes them (lr_schedule.hpp:49-57, mcmc_densification.cpp:38-47)
float log_linear(int step, float v0, float v1, int max_steps) {
    if (step >= max_steps) return v1;
    if (step <= 0) return v0;
    const float t = static_cast<float>(step) / static_cast<float>(max_steps);
    const float log_ratio = std::log(v1 / v0);
    return v0 * std::exp(t * log_ratio);
}

int degree_for_step(int step, int max_degree) { return std::min(step / 1000, max_degree); }  // lr_schedule.hpp:70-72

#define CUGS_TRY_RT(h, expr) CUGS_CUDA_TRY(h, expr)

// part 0 = all views; part 1 = everything up to and including the classification pass of the LAST view's backward
// (touch mask, dL/dmeans_2d and statistics final); part 2 = the rest of the last view's backward. The caller
// starts the MAX all-reduce of the mask between parts 1 and 2.
int enqueue_views(cugs_trainer* t, cudaStream_t s, int degree, int part) {
    cugs_handle_t* h = t->h;
    const int V = (int)t->views.size();
    if (part == 2) {
        const bool two2 = t->frames == 2 && V > 1;
        const int v = V - 1, slot = two2 ? (v & 1) : 0;
        FrameBuf& f = t->f[slot];
        cugs_view_t view = t->views[v];
        view.active_sh_degree = degree;
        view.num_coeffs = t->C;
        for (int c = 0; c < 3; ++c) view.bg[c] = t->cfg.background[c];
        view.scale_modifier = 1.0f;
        const float* dL = t->dLs[v] ? t->dLs[v] : f.dL;
        const bool stats = t->cfg.accumulate_stats && t->t.grad_accum;
        const int flags = (v > 0 ? CUGS_BWD_ACCUMULATE : 0) | CUGS_BWD_SPARSE_ROWS | CUGS_BWD_RESUME_AFTER_MASK;
        if (int e = cugs_b200_render_backward(
                h, s, t->n, &view, t->t.params[0], t->t.params[4], t->t.params[3], t->t.params[2], t->t.params[1],
                f.means_2d, f.cov, f.radii, f.rgb, f.opa, f.gidx, f.ranges, f.final_T, f.n_contrib, dL,
                t->t.grads[0], t->t.grads[4], t->t.grads[3], t->t.grads[2], t->t.grads[1], t->t.dL_dmeans_2d,
                stats ? t->t.grad_accum : nullptr, stats ? t->t.grad_count : nullptr,
                stats ? t->t.max_radii : nullptr, t->t.touch_mask, flags, f.ws, f.ws_bytes))
            return e;
        k_end_views<<<1, 1, 0, s>>>(t->res_dev, t->dyn_dev);
        CUGS_LAUNCH_CHECK(h, "k_end_views");
        return CUGS_OK;
    }
    k_begin_step<<<1, 1, 0, s>>>(t->res_dev);
    CUGS_LAUNCH_CHECK(h, "k_begin_step");
    const bool two = t->frames == 2 && V > 1;
    if (two) CUGS_TRY_RT(h, cudaEventRecord(t->ev_fork, s));
    // end-to-end mode: the target image of every view comes from (pinned) host memory INSIDE the step. The copies
    // run on their own stream in view order, each under the rendering of its own view; the loss waits for it.
    bool any_copy = false;
    for (int v = 0; v < V; ++v) any_copy |= t->targets_host[v] != nullptr && t->dLs[v] == nullptr;
    if (any_copy) {
        CUGS_TRY_RT(h, cudaEventRecord(t->ev_copy_fork, s));
        CUGS_TRY_RT(h, cudaStreamWaitEvent(t->copy, t->ev_copy_fork, 0));
    }
    bool have_prev = false;
    int prev_slot = 0;
    for (int v = 0; v < V; ++v) {
        const int slot = two ? (v & 1) : 0;
        cudaStream_t sv = (two && slot == 1) ? t->aux : s;
        if (two && slot == 1 && v == 1) CUGS_TRY_RT(h, cudaStreamWaitEvent(sv, t->ev_fork, 0));
        FrameBuf& f = t->f[slot];
        cugs_view_t view = t->views[v];
        view.active_sh_degree = degree;
        view.num_coeffs = t->C;
        for (int c = 0; c < 3; ++c) view.bg[c] = t->cfg.background[c];
        view.scale_modifier = 1.0f;
        if (int e = cugs_b200_render_forward(h, sv, t->n, t->p_cap, &view, t->t.params[0], t->t.params[4],
                                             t->t.params[3], t->t.params[2], t->t.params[1], f.means_2d, f.depths,
                                             f.cov, f.radii, f.rgb, f.opa, f.gidx, f.ranges, f.color, f.final_T,
                                             f.n_contrib, f.ws, f.ws_bytes, f.pairs, f.pairs_bytes, f.status2))
            return e;
        const float* dL = f.dL;
        if (t->targets_host[v] != nullptr && t->dLs[v] == nullptr) {
            // (every such view has its own device buffer -- checked in set_views -- and the copy stream was forked
            //  from this step's start, i.e. after the previous step's consumers)
            CUGS_TRY_RT(h, cudaMemcpyAsync(const_cast<float*>(t->targets[v]), t->targets_host[v],
                                           (size_t)t->W * t->H * 3 * sizeof(float), cudaMemcpyHostToDevice, t->copy));
            CUGS_TRY_RT(h, cudaEventRecord(t->ev_copied[v & 1], t->copy));
            CUGS_TRY_RT(h, cudaStreamWaitEvent(sv, t->ev_copied[v & 1], 0));
        }
        if (t->dLs[v] != nullptr) {  // forward+backward only: the caller supplies dL/dcolor
            dL = t->dLs[v];
            CUGS_TRY_RT(h, cudaMemsetAsync(f.scalars3, 0, 16, sv));
        } else if (int e = cugs_b200_loss_l1_ssim(h, sv, t->W, t->H, t->cfg.lambda_ssim, 11, f.color, t->targets[v],
                                                  f.dL, f.scalars3, f.loss_ws, f.loss_bytes, nullptr)) {
            return e;
        }
        // the gradient arena is shared: view v adds to what view v-1 wrote
        if (have_prev && two) CUGS_TRY_RT(h, cudaStreamWaitEvent(sv, t->ev_bwd[prev_slot], 0));
        const bool stats = t->cfg.accumulate_stats && t->t.grad_accum;
        const int flags = (v > 0 ? CUGS_BWD_ACCUMULATE : 0) | (t->t.touch_mask ? CUGS_BWD_SPARSE_ROWS : 0) |
                          ((part == 1 && v == V - 1) ? CUGS_BWD_STOP_AFTER_MASK : 0);
        if (int e = cugs_b200_render_backward(
                h, sv, t->n, &view, t->t.params[0], t->t.params[4], t->t.params[3], t->t.params[2], t->t.params[1],
                f.means_2d, f.cov, f.radii, f.rgb, f.opa, f.gidx, f.ranges, f.final_T, f.n_contrib, dL,
                t->t.grads[0], t->t.grads[4], t->t.grads[3], t->t.grads[2], t->t.grads[1], t->t.dL_dmeans_2d,
                stats ? t->t.grad_accum : nullptr, stats ? t->t.grad_count : nullptr,
                stats ? t->t.max_radii : nullptr, t->t.touch_mask, flags, f.ws, f.ws_bytes))
            return e;
        k_fold_frame<<<1, 1, 0, sv>>>(f.scalars3, f.status2, t->res_dev);
        CUGS_LAUNCH_CHECK(h, "k_fold_frame");
        if (two) CUGS_TRY_RT(h, cudaEventRecord(t->ev_bwd[slot], sv));
        have_prev = true;
        prev_slot = slot;
    }
    if (two) {  // join: everything the auxiliary stream did is ordered before what follows on s
        CUGS_TRY_RT(h, cudaEventRecord(t->ev_join, t->aux));
        CUGS_TRY_RT(h, cudaStreamWaitEvent(s, t->ev_join, 0));
    }
    if (any_copy) {  // join the copy stream (needed for graph capture; all its work was waited on already)
        CUGS_TRY_RT(h, cudaEventRecord(t->ev_copy_join, t->copy));
        CUGS_TRY_RT(h, cudaStreamWaitEvent(s, t->ev_copy_join, 0));
    }
    if (part == 1) return CUGS_OK;  // k_end_views closes part 2
    k_end_views<<<1, 1, 0, s>>>(t->res_dev, t->dyn_dev);
    CUGS_LAUNCH_CHECK(h, "k_end_views");
    return CUGS_OK;
}

int enqueue_update(cugs_trainer* t, cudaStream_t s) {
    cugs_handle_t* h = t->h;
    const int64_t n = t->n;
    const int64_t counts[5] = {3 * n, 3 * (int64_t)t->C * n, n, 3 * n, 4 * n};
    const float lr[5] = {0, 0, 0, 0, 0};  // read from dyn
    const float grad_scale = 1.0f / (float)(t->total_views > 0 ? t->total_views : 1);
    const float lo = t->cfg.mcmc ? t->cfg.lambda_opacity : 0.0f, ls = t->cfg.mcmc ? t->cfg.lambda_scale : 0.0f;
    k_prepare_update<<<1, 1, 0, s>>>(t->res_dev, t->dyn_dev, (double)t->cfg.beta1, (double)t->cfg.beta2);
    CUGS_LAUNCH_CHECK(h, "k_prepare_update");
    if (int e = cugs_adam_launch(h, s, t->t.params, t->t.grads, t->t.adam_m, t->t.adam_v, counts, lr, t->cfg.beta1,
                                 t->cfg.beta2, t->cfg.eps, 1.0f, 1.0f, grad_scale, lo, ls, t->dyn_dev))
        return e;
    if (t->cfg.mcmc) {
        if (int e = cugs_noise_launch(h, s, n, t->t.params[0], t->t.params[3], t->t.params[2], 0.0f,
                                      t->cfg.noise_gate_k, t->cfg.noise_gate_t, t->cfg.noise_seed, 0, nullptr,
                                      t->dyn_dev))
            return e;
    }
    return CUGS_OK;
}

int enqueue_phases(cugs_trainer* t, cudaStream_t s, int phases, int degree) {
    if (phases & 1)
        if (int e = enqueue_views(t, s, degree, 0)) return e;
    if (phases & 4)
        if (int e = enqueue_views(t, s, degree, 1)) return e;
    if (phases & 8)
        if (int e = enqueue_views(t, s, degree, 2)) return e;
    if (phases & 2)
        if (int e = enqueue_update(t, s)) return e;
    // results to pinned memory; the caller reads them after any later synchronisation of `s`
    CUGS_CUDA_TRY(t->h, cudaMemcpyAsync(t->res_pinned, t->res_dev, sizeof(StepResult), cudaMemcpyDeviceToHost, s));
    return CUGS_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" size_t cugs_b200_trainer_workspace_bytes(int64_t n, int num_coeffs, int width, int height,
                                                    int64_t p_capacity, int frames_in_flight) {
    (void)num_coeffs;
    if (n < 0 || width <= 0 || height <= 0 || p_capacity < 0) return 0;
    const int fr = frames_in_flight >= 2 ? 2 : 1;
    return (size_t)fr * frame_bytes(n, width, height, p_capacity, nullptr, nullptr) + a256(sizeof(StepDyn)) +
           a256(sizeof(StepResult));
}

extern "C" int cugs_b200_trainer_create(cugs_handle_t* h, int64_t n, int num_coeffs, int width, int height,
                                        int64_t p_capacity, const cugs_train_config_t* cfg,
                                        const cugs_train_tensors_t* tensors, void* workspace, size_t workspace_bytes,
                                        cugs_trainer_t** out) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, out != nullptr, "out is null");
    *out = nullptr;
    CUGS_REQUIRE(h, n > 0 && num_coeffs >= 1 && width > 0 && height > 0 && p_capacity > 0, "bad sizes");
    CUGS_REQUIRE(h, cfg && tensors && workspace, "null pointer");
    CUGS_REQUIRE(h, cfg->max_sh_degree >= 0 && cfg->max_sh_degree <= 3 &&
                        (cfg->max_sh_degree + 1) * (cfg->max_sh_degree + 1) <= num_coeffs, "bad SH degree");
    for (int k = 0; k < 5; ++k)
        CUGS_REQUIRE(h, tensors->params[k] && tensors->adam_m[k] && tensors->adam_v[k] && tensors->grads[k],
                     "null parameter / moment / gradient pointer");
    CUGS_REQUIRE(h, tensors->dL_dmeans_2d != nullptr, "dL_dmeans_2d is null");
    const bool any_stats = tensors->grad_accum || tensors->grad_count || tensors->max_radii;
    CUGS_REQUIRE(h, !any_stats || (tensors->grad_accum && tensors->grad_count && tensors->max_radii),
                 "stats pointers must be all set or all null");
    const size_t need = cugs_b200_trainer_workspace_bytes(n, num_coeffs, width, height, p_capacity,
                                                          cfg->frames_in_flight);
    if (workspace_bytes < need)
        return set_error(h, CUGS_ERR_WORKSPACE, "trainer workspace too small: %zu < %zu", workspace_bytes, need);
    cugs_trainer* t = new cugs_trainer();
    t->h = h; t->n = n; t->p_cap = p_capacity; t->C = num_coeffs; t->W = width; t->H = height;
    t->frames = cfg->frames_in_flight >= 2 ? 2 : 1;
    t->cfg = *cfg;
    t->t = *tensors;
    char* base = static_cast<char*>(workspace);
    size_t off = 0;
    for (int k = 0; k < t->frames; ++k) off += frame_bytes(n, width, height, p_capacity, &t->f[k], base + off);
    t->dyn_dev = reinterpret_cast<StepDyn*>(base + off); off += a256(sizeof(StepDyn));
    t->res_dev = reinterpret_cast<StepResult*>(base + off); off += a256(sizeof(StepResult));
    t->dyn_pinned = nullptr; t->res_pinned = nullptr; t->aux = nullptr; t->cap = nullptr; t->copy = nullptr;
    t->ev_copy_fork = t->ev_copied[0] = t->ev_copied[1] = t->ev_copy_join = nullptr;
    t->dyn_seq = 0; t->total_views = 0; t->views_gen = 0;
    for (int k = 0; k < 16; ++k) {
        t->exec[k] = nullptr; t->exec_degree[k] = -1; t->exec_gen[k] = 0; t->warmed[k] = false; t->graph_kernels[k] = 0;
    }
    for (int k = 0; k < kDynRing; ++k) { t->ring_ev[k] = nullptr; t->ring_ev_used[k] = false; }
    t->ev_fork = t->ev_bwd[0] = t->ev_bwd[1] = t->ev_join = nullptr;
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&t->dyn_pinned), sizeof(StepDyn) * kDynRing,
                                  cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&t->res_pinned), sizeof(StepResult),
                                            cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->aux, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->cap, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->copy, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_copy_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_copied[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_copied[1], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_copy_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_bwd[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_bwd[1], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming);
    for (int k = 0; k < kDynRing && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&t->ring_ev[k], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        set_error(h, (int)e, "trainer_create: %s", cudaGetErrorString(e));
        cugs_b200_trainer_destroy(t);
        return (int)e;
    }
    std::memset(t->res_pinned, 0, sizeof(StepResult));
    e = cudaMemset(t->res_dev, 0, sizeof(StepResult));
    if (e != cudaSuccess) {
        set_error(h, (int)e, "trainer_create: %s", cudaGetErrorString(e));
        cugs_b200_trainer_destroy(t);
        return (int)e;
    }
    *out = t;
    return CUGS_OK;
}

extern "C" void cugs_b200_trainer_destroy(cugs_trainer_t* t) {
    if (!t) return;
    for (int k = 0; k < 16; ++k)
        if (t->exec[k]) cudaGraphExecDestroy(t->exec[k]);
    if (t->dyn_pinned) cudaFreeHost(t->dyn_pinned);
    if (t->res_pinned) cudaFreeHost(t->res_pinned);
    if (t->aux) cudaStreamDestroy(t->aux);
    if (t->cap) cudaStreamDestroy(t->cap);
    if (t->copy) cudaStreamDestroy(t->copy);
    for (cudaEvent_t ev : {t->ev_copy_fork, t->ev_copied[0], t->ev_copied[1], t->ev_copy_join})
        if (ev) cudaEventDestroy(ev);
    cudaEvent_t evs[4] = {t->ev_fork, t->ev_bwd[0], t->ev_bwd[1], t->ev_join};
    for (cudaEvent_t ev : evs)
        if (ev) cudaEventDestroy(ev);
    for (int k = 0; k < kDynRing; ++k)
        if (t->ring_ev[k]) cudaEventDestroy(t->ring_ev[k]);
    delete t;
}

extern "C" int cugs_b200_trainer_set_views(cugs_trainer_t* t, int num_views, const cugs_view_t* views,
                                           const float* const* targets_dev, const float* const* dL_dcolor_dev,
                                           const float* const* targets_host_pinned, int total_views_per_step) {
    if (!t) return CUGS_ERR_INVALID_ARG;
    cugs_handle_t* h = t->h;
    CUGS_REQUIRE(h, num_views >= 0 && (num_views == 0 || (views && (targets_dev || dL_dcolor_dev))), "bad views");
    CUGS_REQUIRE(h, total_views_per_step >= num_views, "total_views_per_step < num_views");
    auto tgt = [&](int v) { return targets_dev ? targets_dev[v] : nullptr; };
    auto dl = [&](int v) { return dL_dcolor_dev ? dL_dcolor_dev[v] : nullptr; };
    auto th = [&](int v) { return targets_host_pinned ? targets_host_pinned[v] : nullptr; };
    bool same = (int)t->views.size() == num_views && t->total_views == total_views_per_step;
    for (int v = 0; same && v < num_views; ++v)
        same = std::memcmp(&t->views[v], &views[v], sizeof(cugs_view_t)) == 0 && t->targets[v] == tgt(v) &&
               t->dLs[v] == dl(v) && t->targets_host[v] == th(v);
    if (same) return CUGS_OK;
    for (int v = 0; v < num_views; ++v) {
        CUGS_REQUIRE(h, views[v].width == t->W && views[v].height == t->H, "view size differs from the trainer's");
        CUGS_REQUIRE(h, tgt(v) != nullptr || dl(v) != nullptr, "a view needs a target image or a given dL/dcolor");
        CUGS_REQUIRE(h, th(v) == nullptr || tgt(v) != nullptr, "a host target needs a device buffer to be copied into");
        for (int u = 0; u < v; ++u)
            CUGS_REQUIRE(h, th(v) == nullptr || tgt(u) != tgt(v), "views with host targets need distinct device buffers");
    }
    t->views.assign(views, views + num_views);
    t->targets.resize(num_views);
    t->dLs.resize(num_views);
    t->targets_host.resize(num_views);
    for (int v = 0; v < num_views; ++v) { t->targets[v] = tgt(v); t->dLs[v] = dl(v); t->targets_host[v] = th(v); }
    t->total_views = total_views_per_step;
    ++t->views_gen;  // invalidates the captured graphs
    return CUGS_OK;
}

// blocking (set-up / inspection only): the step counter lives on the device
extern "C" int cugs_b200_trainer_set_adam_steps(cugs_trainer_t* t, int64_t steps) {
    if (!t || steps < 0) return CUGS_ERR_INVALID_ARG;
    CUGS_CUDA_TRY(t->h, cudaDeviceSynchronize());
    CUGS_CUDA_TRY(t->h, cudaMemcpy(&t->res_dev->adam_steps, &steps, sizeof(int64_t), cudaMemcpyHostToDevice));
    return CUGS_OK;
}

extern "C" int64_t cugs_b200_trainer_adam_steps(const cugs_trainer_t* t) {
    if (!t) return -1;
    int64_t steps = -1;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpy(&steps, &t->res_dev->adam_steps, sizeof(int64_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return steps;
}

// phases: 1 = the rank's views (render -> loss -> backward, gradients summed in place), 2 = the update
// (Adam [+ MCMC regulariser] [+ noise]), 3 = both (single-GPU step). Nothing here blocks.
extern "C" int cugs_b200_trainer_step(cugs_trainer_t* t, void* stream, int step, int phases) {
    if (!t) return CUGS_ERR_INVALID_ARG;
    cugs_handle_t* h = t->h;
    CUGS_REQUIRE(h, phases == 1 || phases == 2 || phases == 3 || phases == 4 || phases == 8,
                 "phases must be 1 (views), 2 (update), 3 (both), 4 (views up to the last mask) or 8 (rest of the views)");
    CUGS_REQUIRE(h, !(phases & 13) || !t->views.empty(), "no views set");
    CUGS_REQUIRE(h, !(phases & 12) || t->t.touch_mask != nullptr, "phases 4 / 8 need the touch mask (sparse rows)");
    cudaStream_t s = (cudaStream_t)stream;
    const int degree = degree_for_step(step, t->cfg.max_sh_degree);  // trainer.cpp:183

    // per-step scalars -> pinned ring slot -> device (in front of the graph / the launches)
    const int slot = (int)(t->dyn_seq % kDynRing);
    if (t->ring_ev_used[slot]) CUGS_CUDA_TRY(h, cudaEventSynchronize(t->ring_ev[slot]));  // only when 64 steps ahead
    StepDyn& d = t->dyn_pinned[slot];
    d.lr[0] = log_linear(step, t->cfg.lr_position_init, t->cfg.lr_position_final, t->cfg.lr_position_max_steps);
    d.lr[1] = t->cfg.lr_sh_coeffs; d.lr[2] = t->cfg.lr_opacities; d.lr[3] = t->cfg.lr_scales; d.lr[4] = t->cfg.lr_rotations;
    d.bc1 = d.bc2 = 1.0f;  // computed on the device from the device-side step counter (k_prepare_update)
    d.noise_lr = t->cfg.mcmc ? log_linear(step, t->cfg.noise_lr_init, t->cfg.noise_lr_final, t->cfg.noise_lr_max_steps) : 0.0f;
    d.step = (unsigned)step;
    d.ok = 1;  // phase 1 overwrites it on the device; an update-only call trusts the caller's exchange
    d.pad[0] = d.pad[1] = 0;
    if (phases == 8) {
        // nothing: the per-step scalars were uploaded by phase 4 of the same step
    } else if (phases == 2) {
        // keep the ok flag the views phase left on the device: copy everything but `ok`
        CUGS_CUDA_TRY(h, cudaMemcpyAsync(t->dyn_dev, &d, offsetof(StepDyn, ok), cudaMemcpyHostToDevice, s));
    } else {
        CUGS_CUDA_TRY(h, cudaMemcpyAsync(t->dyn_dev, &d, sizeof(StepDyn), cudaMemcpyHostToDevice, s));
    }
    CUGS_CUDA_TRY(h, cudaEventRecord(t->ring_ev[slot], s));
    t->ring_ev_used[slot] = true;
    ++t->dyn_seq;

    if (!t->cfg.use_graph) return enqueue_phases(t, s, phases, degree);

    // CUDA graph: the first call of a configuration runs eagerly (sets kernel attributes, warms caches), the
    // second captures, later ones replay
    const bool valid = t->exec[phases] && t->exec_degree[phases] == degree && t->exec_gen[phases] == t->views_gen;
    if (valid) {
        CUGS_CUDA_TRY(h, cudaGraphLaunch(t->exec[phases], s));
        h->launches += t->graph_kernels[phases];  // the kernels inside the graph do launch
        return CUGS_OK;
    }
    if (!t->warmed[phases]) {
        t->warmed[phases] = true;
        return enqueue_phases(t, s, phases, degree);
    }
    if (t->exec[phases]) { cudaGraphExecDestroy(t->exec[phases]); t->exec[phases] = nullptr; }
    // captured on the trainer's own stream (the caller's may be the legacy default stream), replayed on the caller's
    CUGS_CUDA_TRY(h, cudaStreamBeginCapture(t->cap, cudaStreamCaptureModeThreadLocal));
    const unsigned long long l0 = h->launches;
    const int e = enqueue_phases(t, t->cap, phases, degree);
    t->graph_kernels[phases] = h->launches - l0;
    h->launches = l0;  // captured, not launched yet
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(t->cap, &graph);
    if (e != CUGS_OK) {
        if (graph) cudaGraphDestroy(graph);
        return e;
    }
    if (ce != cudaSuccess) return set_error(h, (int)ce, "cudaStreamEndCapture: %s", cudaGetErrorString(ce));
    const cudaError_t ie = cudaGraphInstantiate(&t->exec[phases], graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return set_error(h, (int)ie, "cudaGraphInstantiate: %s", cudaGetErrorString(ie));
    t->exec_degree[phases] = degree;
    t->exec_gen[phases] = t->views_gen;
    CUGS_CUDA_TRY(h, cudaGraphLaunch(t->exec[phases], s));
    h->launches += t->graph_kernels[phases];
    return CUGS_OK;
}

// Blocking read of the most recent step's result: scalars3 = {loss, l1, mean ssim} (mean over this rank's
// views), status3 = {ok, largest pair count of the step, views folded in}.
extern "C" int cugs_b200_trainer_result(cugs_trainer_t* t, void* stream, float scalars3[3], int64_t status3[3]) {
    if (!t) return CUGS_ERR_INVALID_ARG;
    CUGS_CUDA_TRY(t->h, cudaStreamSynchronize((cudaStream_t)stream));
    if (scalars3)
        for (int k = 0; k < 3; ++k) scalars3[k] = t->res_pinned->scalars[k];
    if (status3) {
        status3[0] = t->res_pinned->ok;
        status3[1] = t->res_pinned->max_pairs;
        status3[2] = t->res_pinned->views;
    }
    return CUGS_OK;
}
