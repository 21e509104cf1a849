// api.cu — handle management and the fused render entry points of the C ABI.
//
// cugs_b200_render_plan / _finish / _backward compose the stage kernels exactly as the reference
// host code does (rasterizer/rasterizer.cpp:22-113 render, :115-186 render_backward), minus its
// ~45 ATen allocations and memsets: every scratch buffer lives in one caller-provided workspace.
#include "common.cuh"

#include <cstring>

using namespace cugs;

// internal launchers implemented in the other translation units
int cugs_scan_launch(cugs_handle_t* h, cudaStream_t s, int64_t n, const int32_t* tiles_touched,
                     int32_t* offsets, int64_t* total_dev, int64_t* total_pinned, void* scan_temp,
                     const unsigned* aux_pair, const uint64_t* gather);
int cugs_preprocess_fwd_launch(cugs_handle_t* h, void* stream, int64_t n, const cugs_view_t* v,
                               const float* positions, const float* rotations, const float* scales,
                               const float* opacities, const float* sh_coeffs, float* means_2d, float* depths,
                               float* cov_2d_inv, int32_t* radii, int32_t* tiles_touched, float* rgb,
                               float* opacities_act, float* packed, uint32_t* depth_minmax, uint64_t* gsort);
// (depth_minmax is an optional output of the public stage function only; the fused path sorts all 32
//  depth bits on the N Gaussians -- see DESIGN.md "what was costed and not built")
// tile_binning.cu
size_t cugs_packed_sort_temp_bytes(int64_t n, int passes, int num_tiles);
int cugs_packed_passes(int key_bits);
int cugs_packed_sort(cugs_handle_t* h, cudaStream_t s, int64_t n, int key_bits, uint64_t* a, uint64_t* b,
                     int* out32_last, int num_tiles, int* tile_ranges, void* temp, size_t temp_bytes,
                     const int64_t* n_dev, bool hist_done, const int* payload_src, int* payload_dst);
int cugs_duplicate_sorted_hist(cugs_handle_t* h, cudaStream_t s, int64_t n, int width, int height,
                               const uint64_t* sorted_elts, const float* means_2d, const int32_t* radii,
                               const int32_t* tiles_touched, const int32_t* offsets_sorted, int64_t p, uint64_t* pairs,
                               const int64_t* p_dev, int key_bits, int num_tiles, void* sort_temp,
                               size_t sort_temp_bytes);
int cugs_duplicate_sorted(cugs_handle_t* h, cudaStream_t s, int64_t n, int width, int height,
                          const uint64_t* sorted_elts, const float* means_2d, const int32_t* radii,
                          const int32_t* tiles_touched, const int32_t* offsets_sorted, int64_t p, uint64_t* pairs,
                          const int64_t* p_dev);
int cugs_blend_bwd_accumulate(cugs_handle_t* h, cudaStream_t s, int64_t n, const cugs_view_t* v,
                              const int32_t* tile_ranges, const int32_t* gaussian_idx,
                              const float* means_2d, const float* cov_2d_inv, const float* rgb,
                              const float* opacities_act, const float* packed, const float* dL_dcolor,
                              const float* final_T, const int32_t* n_contrib, float* grad_acc);
int cugs_preprocess_bwd_launch(cugs_handle_t* h, cudaStream_t s, int64_t n, const cugs_view_t* v,
                               const float* positions, const float* rotations, const float* scales,
                               const float* opacities, const float* sh_coeffs, const int32_t* radii,
                               const float* rgb, const float* dL_dmeans_2d, const float* dL_dcov_2d_inv,
                               const float* dL_drgb, const float* dL_dopacity_act, float* dL_dpositions,
                               float* dL_drotations, float* dL_dscales, float* dL_dopacities,
                               float* dL_dsh_coeffs, float* grad_accum, float* grad_count,
                               float* max_radii, const float* grad_acc, float* dL_dmeans_2d_out,
                               bool accumulate, int32_t* touch_mask, bool sparse_rows, int32_t* list,
                               int32_t* list_count, int phase);
extern "C" int cugs_b200_sort_num_passes(int depth_bits, int tile_bits);
extern "C" int cugs_b200_sort_pairs_pingpong(cugs_handle_t* h, void* stream, int64_t p, int depth_bits,
                                             int tile_bits, uint64_t* keys_a, int32_t* vals_a,
                                             uint64_t* keys_b, int32_t* vals_b, void* temp,
                                             size_t temp_bytes, int* result_in_b);

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
extern "C" int cugs_b200_abi_version(void) { return CUGS_B200_ABI_VERSION; }

extern "C" int cugs_b200_create(int device, cugs_handle_t** out) {
    if (!out) return CUGS_ERR_INVALID_ARG;
    *out = nullptr;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return (int)e;
    if (prop.major != 10) return CUGS_ERR_NOT_BLACKWELL;  // built for sm_100a only; no fallback
    cugs_handle* h = new cugs_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->err[0] = 0;
    h->pinned = nullptr;
    e = cudaHostAlloc(reinterpret_cast<void**>(&h->pinned), kPinnedWords * sizeof(int64_t),
                      cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) { delete h; return (int)e; }
    std::memset(h->pinned, 0, kPinnedWords * sizeof(int64_t));
    h->pinned_seq = 0;
    h->timing = false;
    h->launches = 0;
    h->last_sort_passes = 0;
    h->last_sort_key_bits = 0;
    for (int i = 0; i < 12; ++i) { h->ev[i] = nullptr; h->ev_recorded[i] = false; }
    *out = h;
    return CUGS_OK;
}

extern "C" void cugs_b200_destroy(cugs_handle_t* h) {
    if (!h) return;
    if (h->pinned) cudaFreeHost(h->pinned);
    for (int i = 0; i < 12; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    delete h;
}

// ------------------------------------------------------------------------------------------------
// optional stage timing (events on the caller's stream at the stage boundaries)
// ------------------------------------------------------------------------------------------------
extern "C" int cugs_b200_set_stage_timing(cugs_handle_t* h, int enable) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    if (enable) {
        for (int i = 0; i < 12; ++i)
            if (!h->ev[i]) CUGS_CUDA_TRY(h, cudaEventCreate(&h->ev[i]));
    }
    for (int i = 0; i < 12; ++i) h->ev_recorded[i] = false;
    h->timing = enable != 0;
    return CUGS_OK;
}

static inline void mark(cugs_handle_t* h, int i, cudaStream_t s) {
    if (h->timing) {
        cudaEventRecord(h->ev[i], s);
        h->ev_recorded[i] = true;
    }
}

extern "C" int cugs_b200_last_sort_plan(const cugs_handle_t* h, int* passes, int* key_bits) {
    if (!h) return CUGS_ERR_INVALID_ARG;
    if (passes) *passes = h->last_sort_passes;
    if (key_bits) *key_bits = h->last_sort_key_bits;
    return CUGS_OK;
}

extern "C" int cugs_b200_get_stage_ms(cugs_handle_t* h, float* ms8) {
    CUGS_REQUIRE(h, h != nullptr && ms8 != nullptr, "null pointer");
    // {first event, last event} of each stage; events 0-2 plan, 3-7 finish, 8-10 backward
    // stages: preprocess, scan (gather-scan in depth order), duplicate, sort (depth sort of the N
    // Gaussians [1->11] + tile sort of the P pairs incl. histogram/ranges [4->5]), tile_ranges (now
    // part of the sort: 0), blend_fwd, blend_bwd, preprocess_bwd
    static const int span[CUGS_NUM_STAGES][2] = {{0, 1}, {11, 2}, {3, 4}, {4, 5}, {5, 6}, {6, 7}, {8, 9}, {9, 10}};
    for (int k = 0; k < CUGS_NUM_STAGES; ++k) {
        ms8[k] = -1.0f;
        const int a = span[k][0], b = span[k][1];
        if (!h->timing || !h->ev_recorded[a] || !h->ev_recorded[b]) continue;
        CUGS_CUDA_TRY(h, cudaEventSynchronize(h->ev[b]));
        float ms = 0.0f;
        CUGS_CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]));
        ms8[k] = ms;
    }
    if (h->timing && h->ev_recorded[1] && h->ev_recorded[11] && ms8[3] >= 0.0f) {  // add the depth sort to "sort"
        float ms = 0.0f;
        CUGS_CUDA_TRY(h, cudaEventSynchronize(h->ev[11]));
        CUGS_CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev[1], h->ev[11]));
        ms8[3] += ms;
    }
    return CUGS_OK;
}

extern "C" const char* cugs_b200_last_error(const cugs_handle_t* h) { return h ? h->err : "null handle"; }
extern "C" int cugs_b200_sm_count(const cugs_handle_t* h) { return h ? h->sm_count : 0; }
extern "C" uint64_t cugs_b200_launch_count(const cugs_handle_t* h) { return h ? h->launches : 0; }

// what the FP32 / MUFU rooflines of the blend kernels are computed from ("read both from
// cudaGetDeviceProperties, do not assume"): SM count, maximum SM clock in kHz, L2 size in bytes
extern "C" int cugs_b200_device_info(const cugs_handle_t* h, int* sm_count, int* clock_khz, int* l2_bytes) {
    if (!h) return CUGS_ERR_INVALID_ARG;
    int v = 0;
    if (sm_count) *sm_count = h->sm_count;
    if (clock_khz) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, h->device) != cudaSuccess) return CUGS_ERR_INVALID_ARG;
        *clock_khz = v;
    }
    if (l2_bytes) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, h->device) != cudaSuccess) return CUGS_ERR_INVALID_ARG;
        *l2_bytes = v;
    }
    return CUGS_OK;
}

// ------------------------------------------------------------------------------------------------
// workspace layout of one frame
// ------------------------------------------------------------------------------------------------
namespace {

struct FrameWorkspace {
    float* packed;          // [N,12]  blend records
    int32_t* tiles_touched; // [N]
    int32_t* tiles_sorted;  // [N]     tiles_touched in depth order (written by the depth sort's last pass)
    int32_t* offsets;       // [N]     pair offsets in DEPTH order
    void* scan_temp;
    int64_t* total_dev;
    uint64_t* gsort_a;      // [N]     depth_key << 32 | index; sorted result ends up here (4 passes)
    uint64_t* gsort_b;      // [N]
    void* gsort_temp;
    size_t gsort_temp_bytes;
    float* grad_acc;        // [N,12]
    int32_t* bwd_list;      // [N]     compact list of the Gaussians with a non-zero 2-D gradient (sparse backward)
    uint64_t* pairs_a;      // [Pcap]  tile << 32 | index
    uint64_t* pairs_b;      // [Pcap]
    void* psort_temp;
    size_t psort_temp_bytes;
    size_t head_bytes, pair_bytes, total_bytes;
};

constexpr int kMaxTilesForWorkspace = 48 * 1024;  // tile histogram lives in shared memory (<= 192 KB)

// the P-sized regions (pair ping-pong buffers + the pair sort's look-back state)
void carve_pairs(FrameWorkspace& w, void* base, int64_t pcap) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off += align_up(bytes ? bytes : 16, 256);
        return p;
    };
    const size_t pp = (size_t)(pcap > 0 ? pcap : 0);
    w.pairs_a = static_cast<uint64_t*>(take(pp * 8));
    w.pairs_b = static_cast<uint64_t*>(take(pp * 8));
    w.psort_temp_bytes = cugs_packed_sort_temp_bytes(pcap, 4, kMaxTilesForWorkspace);
    w.psort_temp = take(w.psort_temp_bytes);
    w.pair_bytes = off;
}

FrameWorkspace carve(void* base, int64_t n, int64_t pcap) {
    FrameWorkspace w{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off += align_up(bytes ? bytes : 16, 256);
        return p;
    };
    const size_t nn = (size_t)(n > 0 ? n : 0);
    // N-sized regions first: their addresses do not depend on the pair capacity, so plan, finish
    // and render_backward of one frame agree on them whatever P turns out to be.
    w.packed = static_cast<float*>(take(nn * 48));
    w.tiles_touched = static_cast<int32_t*>(take(nn * 4));
    w.tiles_sorted = static_cast<int32_t*>(take(nn * 4));
    w.offsets = static_cast<int32_t*>(take(nn * 4));
    w.scan_temp = take(cugs_b200_scan_temp_bytes(n));
    w.total_dev = static_cast<int64_t*>(take(64));
    w.gsort_a = static_cast<uint64_t*>(take(nn * 8));
    w.gsort_b = static_cast<uint64_t*>(take(nn * 8));
    w.gsort_temp_bytes = cugs_packed_sort_temp_bytes(n, 4, 0);
    w.gsort_temp = take(w.gsort_temp_bytes);
    w.grad_acc = static_cast<float*>(take(nn * 48));
    w.bwd_list = static_cast<int32_t*>(take(nn * 4));
    w.head_bytes = off;
    // P-sized scratch, live only inside render_finish: the tail of the same block, or a separate one
    carve_pairs(w, base ? static_cast<char*>(base) + off : nullptr, pcap);
    w.total_bytes = off + w.pair_bytes;
    return w;
}

int ceil_log2(int x) {
    int b = 0;
    while ((1 << b) < x) ++b;
    return b;
}

}  // namespace

extern "C" size_t cugs_b200_render_workspace_bytes(int64_t n, int64_t p_capacity) {
    return carve(nullptr, n, p_capacity).total_bytes;
}

extern "C" size_t cugs_b200_render_pair_scratch_bytes(int64_t p_capacity) {
    FrameWorkspace w{};
    carve_pairs(w, nullptr, p_capacity);
    return w.pair_bytes;
}

// ------------------------------------------------------------------------------------------------
// render: front (preprocess + depth sort + scan -> P on the device) and back (dup + tile sort +
// ranges + blend, every launch sized on a pair CAPACITY and reading P on the device). The public
// entry points are plan (= front + ONE blocking read of P, as the reference), finish (= back with
// capacity = the P the caller read), and render_forward (= front + back, no host round trip).
// ------------------------------------------------------------------------------------------------
static int forward_front(cugs_handle_t* h, cudaStream_t s, int64_t n, const cugs_view_t* v, const float* positions,
                         const float* rotations, const float* scales, const float* opacities,
                         const float* sh_coeffs, float* means_2d, float* depths, float* cov_2d_inv,
                         int32_t* radii, float* rgb, float* opacities_act, const FrameWorkspace& w,
                         int64_t* total_pinned) {
    mark(h, 0, s);
    if (int e = cugs_preprocess_fwd_launch(h, s, n, v, positions, rotations, scales, opacities, sh_coeffs, means_2d,
                                           depths, cov_2d_inv, radii, w.tiles_touched, rgb, opacities_act, w.packed,
                                           nullptr, w.gsort_a))
        return e;
    mark(h, 1, s);
    // depth sort of the N Gaussians (the depth-bit passes of the reference's 64-bit sort, hoisted
    // in front of duplicateWithKeys), then the scan of tiles_touched in depth order
    // (the last pass also delivers tiles_touched in depth order: the scan and duplicateWithKeys read it contiguously)
    if (int e = cugs_packed_sort(h, s, n, 32, w.gsort_a, w.gsort_b, nullptr, 0, nullptr, w.gsort_temp,
                                 w.gsort_temp_bytes, nullptr, false, w.tiles_touched, w.tiles_sorted))
        return e;
    mark(h, 11, s);
    if (int e = cugs_scan_launch(h, s, n, w.tiles_sorted, w.offsets, w.total_dev, total_pinned, w.scan_temp,
                                 nullptr, nullptr))
        return e;
    mark(h, 2, s);
    return CUGS_OK;
}

// p_cap = capacity of gaussian_idx and of the workspace's pair buffers; the pair count itself is read
// on the device from w.total_dev (written by the scan) and clamped to p_cap
static int forward_back(cugs_handle_t* h, cudaStream_t s, int64_t n, int64_t p_cap, const cugs_view_t* v,
                        const float* means_2d, const float* cov_2d_inv, const int32_t* radii, const float* rgb,
                        const float* opacities_act, int32_t* gaussian_idx, int32_t* tile_ranges, float* color,
                        float* final_T, int32_t* n_contrib, const FrameWorkspace& w) {
    const int ntx = (v->width + kTile - 1) / kTile, nty = (v->height + kTile - 1) / kTile;
    const int num_tiles = ntx * nty;
    mark(h, 3, s);
    if (num_tiles > kMaxTilesForWorkspace)
        return set_error(h, CUGS_ERR_UNSUPPORTED, "%d tiles > %d is not supported by the fused path", num_tiles,
                         kMaxTilesForWorkspace);
    if (n > 0 && p_cap > 0) {
        const int tile_bits = ceil_log2(num_tiles);
#ifndef CUGS_NO_FUSED_PAIR_HIST
        // duplicateWithKeys also takes the pair sort's histograms (no separate read of the P pairs)
        if (int e = cugs_duplicate_sorted_hist(h, s, n, v->width, v->height, w.gsort_a, means_2d, radii,
                                               w.tiles_sorted, w.offsets, p_cap, w.pairs_a, w.total_dev, tile_bits,
                                               num_tiles, w.psort_temp, w.psort_temp_bytes))
            return e;
        const bool hist_done = true;
#else
        if (int e = cugs_duplicate_sorted(h, s, n, v->width, v->height, w.gsort_a, means_2d, radii, w.tiles_sorted,
                                          w.offsets, p_cap, w.pairs_a, w.total_dev))
            return e;
        const bool hist_done = false;
#endif
        mark(h, 4, s);
        h->last_sort_passes = 4 + cugs_packed_passes(tile_bits);
        h->last_sort_key_bits = 32 + tile_bits;
        // tile histogram -> tile ranges, and the stable sort of the pairs by tile id; the last pass
        // writes the Gaussian indices straight into the caller's buffer
        if (int e = cugs_packed_sort(h, s, p_cap, tile_bits, w.pairs_a, w.pairs_b, gaussian_idx, num_tiles,
                                     tile_ranges, w.psort_temp, w.psort_temp_bytes, w.total_dev, hist_done, nullptr,
                                     nullptr))
            return e;
        mark(h, 5, s);
    } else {
        mark(h, 4, s);
        mark(h, 5, s);
        CUGS_CUDA_TRY(h, cudaMemsetAsync(tile_ranges, 0, (size_t)num_tiles * 2 * sizeof(int), s));
    }
    mark(h, 6, s);
    if (int e = cugs_b200_blend_fwd(h, s, v, tile_ranges, gaussian_idx, means_2d, cov_2d_inv, rgb, opacities_act,
                                    n > 0 ? w.packed : nullptr, color, final_T, n_contrib))
        return e;
    mark(h, 7, s);
    return CUGS_OK;
}

extern "C" int cugs_b200_render_plan(cugs_handle_t* h, void* stream, int64_t n, const cugs_view_t* v,
                                     const float* positions, const float* rotations,
                                     const float* scales, const float* opacities,
                                     const float* sh_coeffs, float* means_2d, float* depths,
                                     float* cov_2d_inv, int32_t* radii, float* rgb,
                                     float* opacities_act, void* workspace, size_t workspace_bytes,
                                     int64_t* p_host) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0, "n must be >= 0");
    CUGS_REQUIRE(h, p_host != nullptr, "p_host is null");
    *p_host = 0;
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, workspace != nullptr, "workspace is null");
    if (workspace_bytes < cugs_b200_render_workspace_bytes(n, 0))
        return set_error(h, CUGS_ERR_WORKSPACE, "workspace too small for N=%lld: %zu < %zu", (long long)n,
                         workspace_bytes, cugs_b200_render_workspace_bytes(n, 0));
    cudaStream_t s = (cudaStream_t)stream;
    // the N-sized regions come first in the layout, so carving with pcap = 0 addresses them
    const FrameWorkspace w = carve(workspace, n, 0);
    int64_t* slot = cugs_pinned_slot(h);  // this frame's own pinned word
    if (int e = forward_front(h, s, n, v, positions, rotations, scales, opacities, sh_coeffs, means_2d, depths,
                              cov_2d_inv, radii, rgb, opacities_act, w, slot))
        return e;
    CUGS_CUDA_TRY(h, cudaStreamSynchronize(s));  // the one blocking read (reference: sorting.cu:146)
    *p_host = *slot;
    return CUGS_OK;
}

extern "C" int cugs_b200_render_finish(cugs_handle_t* h, void* stream, int64_t n, int64_t p,
                                       const cugs_view_t* v, const float* means_2d, const float* depths,
                                       const float* cov_2d_inv, const int32_t* radii, const float* rgb,
                                       const float* opacities_act, int32_t* gaussian_idx,
                                       int32_t* tile_ranges, float* color, float* final_T,
                                       int32_t* n_contrib, void* workspace, size_t workspace_bytes,
                                       void* pair_scratch, size_t pair_scratch_bytes) {
    (void)depths;
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, v != nullptr && v->width > 0 && v->height > 0, "bad view");
    CUGS_REQUIRE(h, n >= 0 && p >= 0, "n and p must be >= 0");
    CUGS_REQUIRE(h, tile_ranges && color && final_T && n_contrib, "null output");
    CUGS_REQUIRE(h, p == 0 || gaussian_idx != nullptr, "gaussian_idx is null");
    cudaStream_t s = (cudaStream_t)stream;
    FrameWorkspace w{};
    if (n > 0) {
        CUGS_REQUIRE(h, workspace != nullptr, "workspace is null");
        const int64_t in_ws = pair_scratch ? 0 : p;  // pair buffers: the tail of the workspace, or a separate block
        if (workspace_bytes < cugs_b200_render_workspace_bytes(n, in_ws))
            return set_error(h, CUGS_ERR_WORKSPACE, "workspace too small for N=%lld P=%lld: %zu < %zu",
                             (long long)n, (long long)in_ws, workspace_bytes,
                             cugs_b200_render_workspace_bytes(n, in_ws));
        w = carve(workspace, n, in_ws);
        if (pair_scratch) {
            if (pair_scratch_bytes < cugs_b200_render_pair_scratch_bytes(p))
                return set_error(h, CUGS_ERR_WORKSPACE, "pair scratch too small for P=%lld: %zu < %zu", (long long)p,
                                 pair_scratch_bytes, cugs_b200_render_pair_scratch_bytes(p));
            carve_pairs(w, pair_scratch, p);
        }
    }
    // p is the CAPACITY the caller sized gaussian_idx for (= the P that render_plan returned); the
    // kernels read the frame's own pair count from the workspace, so a stale or foreign p can only
    // truncate the frame, never make a kernel run past a buffer
    return forward_back(h, s, n, p, v, means_2d, cov_2d_inv, radii, rgb, opacities_act, gaussian_idx, tile_ranges,
                        color, final_T, n_contrib, w);
}

namespace cugs {
// {P, P > capacity} -> the caller's status words (device memory; copied to the host whenever the caller
// wants to look, e.g. once per step)
__global__ void k_publish_pairs(const int64_t* __restrict__ total_dev, int64_t p_cap, int64_t* __restrict__ status) {
    const int64_t p = *total_dev;
    status[0] = p;
    status[1] = p > p_cap ? 1 : 0;
}
}  // namespace cugs

extern "C" int cugs_b200_render_forward(cugs_handle_t* h, void* stream, int64_t n, int64_t p_capacity,
                                        const cugs_view_t* v, const float* positions, const float* rotations,
                                        const float* scales, const float* opacities, const float* sh_coeffs,
                                        float* means_2d, float* depths, float* cov_2d_inv, int32_t* radii,
                                        float* rgb, float* opacities_act, int32_t* gaussian_idx,
                                        int32_t* tile_ranges, float* color, float* final_T, int32_t* n_contrib,
                                        void* workspace, size_t workspace_bytes, void* pair_scratch,
                                        size_t pair_scratch_bytes, int64_t* status_dev) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, v != nullptr && v->width > 0 && v->height > 0, "bad view");
    CUGS_REQUIRE(h, n >= 0 && p_capacity >= 0, "n and p_capacity must be >= 0");
    CUGS_REQUIRE(h, tile_ranges && color && final_T && n_contrib, "null output");
    CUGS_REQUIRE(h, p_capacity == 0 || gaussian_idx != nullptr, "gaussian_idx is null");
    cudaStream_t s = (cudaStream_t)stream;
    FrameWorkspace w{};
    if (n > 0) {
        CUGS_REQUIRE(h, workspace != nullptr, "workspace is null");
        const int64_t in_ws = pair_scratch ? 0 : p_capacity;
        if (workspace_bytes < cugs_b200_render_workspace_bytes(n, in_ws))
            return set_error(h, CUGS_ERR_WORKSPACE, "workspace too small for N=%lld capacity=%lld: %zu < %zu",
                             (long long)n, (long long)in_ws, workspace_bytes,
                             cugs_b200_render_workspace_bytes(n, in_ws));
        w = carve(workspace, n, in_ws);
        if (pair_scratch) {
            if (pair_scratch_bytes < cugs_b200_render_pair_scratch_bytes(p_capacity))
                return set_error(h, CUGS_ERR_WORKSPACE, "pair scratch too small for capacity=%lld: %zu < %zu",
                                 (long long)p_capacity, pair_scratch_bytes,
                                 cugs_b200_render_pair_scratch_bytes(p_capacity));
            carve_pairs(w, pair_scratch, p_capacity);
        }
        if (int e = forward_front(h, s, n, v, positions, rotations, scales, opacities, sh_coeffs, means_2d, depths,
                                  cov_2d_inv, radii, rgb, opacities_act, w, nullptr))
            return e;
        if (status_dev) {
            k_publish_pairs<<<1, 1, 0, s>>>(w.total_dev, p_capacity, status_dev);
            CUGS_LAUNCH_CHECK(h, "k_publish_pairs");
        }
    } else if (status_dev) {
        CUGS_CUDA_TRY(h, cudaMemsetAsync(status_dev, 0, 2 * sizeof(int64_t), s));
    }
    return forward_back(h, s, n, p_capacity, v, means_2d, cov_2d_inv, radii, rgb, opacities_act, gaussian_idx,
                        tile_ranges, color, final_T, n_contrib, w);
}

extern "C" int cugs_b200_render_backward(
    cugs_handle_t* h, void* stream, int64_t n, const cugs_view_t* v, const float* positions,
    const float* rotations, const float* scales, const float* opacities, const float* sh_coeffs,
    const float* means_2d, const float* cov_2d_inv, const int32_t* radii, const float* rgb,
    const float* opacities_act, const int32_t* gaussian_idx, const int32_t* tile_ranges,
    const float* final_T, const int32_t* n_contrib, const float* dL_dcolor, float* dL_dpositions,
    float* dL_drotations, float* dL_dscales, float* dL_dopacities, float* dL_dsh_coeffs,
    float* dL_dmeans_2d, float* grad_accum, float* grad_count, float* max_radii, int32_t* touch_mask,
    int flags, void* workspace, size_t workspace_bytes) {
    CUGS_REQUIRE(h, h != nullptr, "handle is null");
    CUGS_REQUIRE(h, n >= 0, "n must be >= 0");
    if (n == 0) return CUGS_OK;
    CUGS_REQUIRE(h, v != nullptr && v->width > 0 && v->height > 0, "bad view");
    CUGS_REQUIRE(h, workspace != nullptr, "workspace is null");
    CUGS_REQUIRE(h, dL_dcolor && final_T && n_contrib && tile_ranges, "null input");
    CUGS_REQUIRE(h, dL_dpositions && dL_drotations && dL_dscales && dL_dopacities && dL_dsh_coeffs &&
                        dL_dmeans_2d, "null output");
    const bool any_stats = grad_accum || grad_count || max_radii;
    CUGS_REQUIRE(h, !any_stats || (grad_accum && grad_count && max_radii),
                 "stats pointers must be all set or all null");
    if (workspace_bytes < cugs_b200_render_workspace_bytes(n, 0))
        return set_error(h, CUGS_ERR_WORKSPACE, "workspace too small for N=%lld: %zu < %zu", (long long)n,
                         workspace_bytes, cugs_b200_render_workspace_bytes(n, 0));
    cudaStream_t s = (cudaStream_t)stream;
    const FrameWorkspace w = carve(workspace, n, 0);
    const bool sparse = (flags & CUGS_BWD_SPARSE_ROWS) != 0 && touch_mask != nullptr;
    const bool stop = (flags & CUGS_BWD_STOP_AFTER_MASK) != 0, resume = (flags & CUGS_BWD_RESUME_AFTER_MASK) != 0;
    CUGS_REQUIRE(h, !(stop || resume) || sparse, "CUGS_BWD_STOP/RESUME_AFTER_MASK need CUGS_BWD_SPARSE_ROWS");
    CUGS_REQUIRE(h, !(stop && resume), "STOP_AFTER_MASK and RESUME_AFTER_MASK are exclusive");
    // stage 1 (rasterizer.cpp:146-158): pixel gradients -> packed per-Gaussian 2-D gradients
    mark(h, 8, s);
    if (!resume)
        if (int e = cugs_blend_bwd_accumulate(h, s, n, v, tile_ranges, gaussian_idx, means_2d, cov_2d_inv, rgb,
                                              opacities_act, w.packed, dL_dcolor, final_T, n_contrib, w.grad_acc))
            return e;
    mark(h, 9, s);
    // stage 2 (rasterizer.cpp:163-176): 2-D gradients -> parameter gradients (+ SH backward, + stats)
    if (int e = cugs_preprocess_bwd_launch(h, s, n, v, positions, rotations, scales, opacities, sh_coeffs,
                                           radii, rgb, nullptr, nullptr, nullptr, nullptr, dL_dpositions,
                                           dL_drotations, dL_dscales, dL_dopacities, dL_dsh_coeffs, grad_accum,
                                           grad_count, max_radii, w.grad_acc, dL_dmeans_2d,
                                           (flags & CUGS_BWD_ACCUMULATE) != 0, touch_mask, sparse,
                                           sparse ? w.bwd_list : nullptr,
                                           sparse ? reinterpret_cast<int32_t*>(w.total_dev + 1) : nullptr,
                                           stop ? 1 : (resume ? 2 : 0)))
        return e;
    mark(h, 10, s);
    return CUGS_OK;
}
