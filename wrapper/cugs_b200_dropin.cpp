// cugs_b200_dropin.cpp — the host side of the drop-in, in the reference's own language (C++ on
// libtorch): DEFINITIONS of the reference's rasterizer / loss / optimizer entry points that only
// validate, allocate the output tensors and call the C ABI of include/cugs_b200.h on raw device
// pointers. It is compiled against the reference's OWN, unmodified headers (the declarations a
// maintainer already has) and replaces, in the reference's build, rasterizer/rasterizer.cpp,
// rasterizer/{projection,sorting,forward,backward,projection_backward}.cu, core/{sh,sh_backward}.cu,
// training/loss.cpp, optimizer/fused_adam.cu and -- the schedule-driven callers that resize the model --
// optimizer/densification.cpp and optimizer/mcmc_densification.cpp. libtorch tensors exist only in
// this file; no CUDA code here.
//
// Build (what oracle/Makefile.ref's `dropin` target does):
//   g++ -std=c++20 -I<reference>/src -I<eigen> -I<torch includes> -Iinclude -c wrapper/cugs_b200_dropin.cpp
//   ... link with libcugs_b200.so
#include <ATen/cuda/CUDAContext.h>
#include <torch/torch.h>

#include <array>
#include <cmath>
#include <mutex>
#include <unordered_map>

#include "core/gaussian.hpp"
#include "core/sh.hpp"
#include "core/sh_backward.hpp"
#include "core/types.hpp"
#include "optimizer/densification.hpp"
#include "optimizer/fused_adam.hpp"
#include "optimizer/mcmc_densification.hpp"
#include "rasterizer/backward.hpp"
#include "rasterizer/forward.hpp"
#include "rasterizer/projection.hpp"
#include "rasterizer/projection_backward.hpp"
#include "rasterizer/rasterizer.hpp"
#include "rasterizer/sorting.hpp"
#include "training/loss.hpp"

#include "cugs_b200.h"

namespace cugs {
namespace {

// ---- handle, stream, status -----------------------------------------------------------------------
cugs_handle_t* handle_for(const torch::Tensor& t) {
    static std::mutex mu;
    static std::array<cugs_handle_t*, 64> handles{};
    const int dev = t.get_device();
    TORCH_CHECK(dev >= 0 && dev < 64, "bad CUDA device index ", dev);
    std::lock_guard<std::mutex> lock(mu);
    if (!handles[dev]) {
        TORCH_CHECK(cugs_b200_abi_version() == CUGS_B200_ABI_VERSION, "libcugs_b200.so has ABI version ",
                    cugs_b200_abi_version(), " but this wrapper was compiled against ", CUGS_B200_ABI_VERSION,
                    ": rebuild the wrapper");
        const int st = cugs_b200_create(dev, &handles[dev]);
        TORCH_CHECK(st == 0, "cugs_b200_create(device=", dev, ") failed with status ", st,
                    " (the library is built for sm_100a only; there is no fallback)");
    }
    return handles[dev];
}

void* current_stream(const torch::Tensor& t) {
    return static_cast<void*>(at::cuda::getCurrentCUDAStream(t.get_device()).stream());
}

#define CUGS_CALL(h, expr)                                                                       \
    do {                                                                                         \
        const int st_ = (expr);                                                                  \
        TORCH_CHECK(st_ == 0, #expr, " failed (status ", st_, "): ", cugs_b200_last_error(h));   \
    } while (0)

torch::Tensor f32c(const torch::Tensor& t) { return t.contiguous().to(torch::kFloat32); }
float* fp(const torch::Tensor& t) { return t.numel() ? t.data_ptr<float>() : nullptr; }
int32_t* ip(const torch::Tensor& t) { return t.numel() ? t.data_ptr<int32_t>() : nullptr; }

cugs_view_t make_view(const CameraInfo& cam, const float bg[3], int degree, int num_coeffs, float scale_mod) {
    cugs_view_t v{};
    v.width = cam.width;
    v.height = cam.height;
    v.fx = cam.intrinsics.fx;
    v.fy = cam.intrinsics.fy;
    v.cx = cam.intrinsics.cx;
    v.cy = cam.intrinsics.cy;
    const Eigen::Matrix4f w2c = cam.world_to_camera();  // core/types.hpp:103-108
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) v.view[r * 4 + c] = w2c(r, c);  // row-major, as projection.cu:227-233
    const Eigen::Vector3f centre = cam.camera_center();  // core/types.hpp:98-100
    v.cam_center[0] = centre.x();
    v.cam_center[1] = centre.y();
    v.cam_center[2] = centre.z();
    for (int i = 0; i < 3; ++i) v.bg[i] = bg ? bg[i] : 0.0f;
    v.active_sh_degree = degree;
    v.num_coeffs = num_coeffs;
    v.scale_modifier = scale_mod;
    return v;
}

int tiles_of(int w, int h) { return ((w + kTileSize - 1) / kTileSize) * ((h + kTileSize - 1) / kTileSize); }

// ---- per-frame workspace cache: RenderOutput has no slot for the packed blend records, so the
// workspace of the most recent frames is remembered by the address of their means_2d tensor ----
struct Frame {
    torch::Tensor workspace;  // the N-sized frame arena (packed blend records, 2-D gradient accumulator)
    torch::Tensor means_2d;   // keeps the key address alive
    uint64_t stamp;           // last use, for LRU eviction
};
std::mutex g_frames_mu;
std::unordered_map<const void*, Frame> g_frames;
uint64_t g_frame_clock = 0;
constexpr size_t kMaxFrames = 8;  // frames a caller may hold between render() and render_backward()

// Bounded LRU: only the least recently used frame is dropped when the cache is full (a caller holding more
// than kMaxFrames RenderOutputs falls back to the stage path for the oldest ones, never for the newest).
void remember_frame(const torch::Tensor& means_2d, const torch::Tensor& ws) {
    std::lock_guard<std::mutex> lock(g_frames_mu);
    if (g_frames.size() >= kMaxFrames && g_frames.find(means_2d.data_ptr()) == g_frames.end()) {
        auto oldest = g_frames.begin();
        for (auto it = g_frames.begin(); it != g_frames.end(); ++it)
            if (it->second.stamp < oldest->second.stamp) oldest = it;
        g_frames.erase(oldest);
    }
    g_frames[means_2d.data_ptr()] = Frame{ws, means_2d, ++g_frame_clock};
}
torch::Tensor recall_frame(const torch::Tensor& means_2d) {
    std::lock_guard<std::mutex> lock(g_frames_mu);
    auto it = g_frames.find(means_2d.data_ptr());
    if (it == g_frames.end()) return torch::Tensor();
    it->second.stamp = ++g_frame_clock;
    return it->second.workspace;
}

}  // namespace

// =====================================================================================================
// rasterizer/rasterizer.hpp
// =====================================================================================================
RenderOutput render(const GaussianModel& model, const CameraInfo& camera, const RenderSettings& settings) {
    TORCH_CHECK(model.is_valid(), "GaussianModel is not valid");
    TORCH_CHECK(model.positions.is_cuda(), "GaussianModel must be on CUDA device");
    const int64_t n = model.num_gaussians();
    const int W = camera.width, H = camera.height;
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(model.positions.device());
    const auto i32 = f32.dtype(torch::kInt32);
    RenderOutput out;
    if (n == 0) {  // background only (rasterizer.cpp:36-55)
        out.color = torch::empty({H, W, 3}, f32);
        for (int c = 0; c < 3; ++c) out.color.select(2, c).fill_(settings.background[c]);
        out.final_T = torch::ones({H, W}, f32);
        out.n_contrib = torch::zeros({H, W}, i32);
        out.means_2d = torch::empty({0, 2}, f32);
        out.depths = torch::empty({0}, f32);
        out.cov_2d_inv = torch::empty({0, 3}, f32);
        out.radii = torch::empty({0}, i32);
        out.rgb = torch::empty({0, 3}, f32);
        out.opacities_act = torch::empty({0}, f32);
        out.gaussian_indices = torch::empty({0}, i32);
        out.tile_ranges = torch::empty({0, 2}, i32);
        return out;
    }
    cugs_handle_t* h = handle_for(model.positions);
    void* stream = current_stream(model.positions);
    const int degree = std::min(settings.active_sh_degree, model.max_sh_degree());  // rasterizer.cpp:60
    const auto pos = f32c(model.positions), rot = f32c(model.rotations), scl = f32c(model.scales),
               opa = f32c(model.opacities), sh = f32c(model.sh_coeffs);
    const cugs_view_t v = make_view(camera, settings.background, degree, (int)sh.size(2), settings.scale_modifier);

    // every element of every output is written by the kernels -> torch::empty, no memsets
    out.means_2d = torch::empty({n, 2}, f32);
    out.depths = torch::empty({n}, f32);
    out.cov_2d_inv = torch::empty({n, 3}, f32);
    out.radii = torch::empty({n}, i32);
    out.rgb = torch::empty({n, 3}, f32);
    out.opacities_act = torch::empty({n}, f32);
    const auto u8 = f32.dtype(torch::kUInt8);
    // the N-sized frame arena lives until render_backward; the P-sized pair scratch is a separate block that
    // goes back to the caching allocator right after render_finish (stream-ordered reuse: no copy, no cudaMalloc)
    auto ws = torch::empty({(int64_t)cugs_b200_render_workspace_bytes(n, 0)}, u8);
    int64_t P = 0;
    CUGS_CALL(h, cugs_b200_render_plan(h, stream, n, &v, fp(pos), fp(rot), fp(scl), fp(opa), fp(sh), fp(out.means_2d),
                                       fp(out.depths), fp(out.cov_2d_inv), ip(out.radii), fp(out.rgb),
                                       fp(out.opacities_act), ws.data_ptr(), (size_t)ws.numel(), &P));
    auto scratch = torch::empty({(int64_t)cugs_b200_render_pair_scratch_bytes(P)}, u8);
    out.gaussian_indices = torch::empty({P}, i32);
    out.tile_ranges = torch::empty({tiles_of(W, H), 2}, i32);
    out.color = torch::empty({H, W, 3}, f32);
    out.final_T = torch::empty({H, W}, f32);
    out.n_contrib = torch::empty({H, W}, i32);
    CUGS_CALL(h, cugs_b200_render_finish(h, stream, n, P, &v, fp(out.means_2d), fp(out.depths), fp(out.cov_2d_inv),
                                         ip(out.radii), fp(out.rgb), fp(out.opacities_act), ip(out.gaussian_indices),
                                         ip(out.tile_ranges), fp(out.color), fp(out.final_T), ip(out.n_contrib),
                                         ws.data_ptr(), (size_t)ws.numel(), scratch.data_ptr(),
                                         (size_t)scratch.numel()));
    remember_frame(out.means_2d, ws);
    return out;
}

BackwardOutput render_backward(const torch::Tensor& dL_dcolor, const RenderOutput& r, const GaussianModel& model,
                               const CameraInfo& camera, const RenderSettings& settings) {
    TORCH_CHECK(dL_dcolor.is_cuda(), "dL_dcolor must be on CUDA device");
    TORCH_CHECK(dL_dcolor.dim() == 3 && dL_dcolor.size(2) == 3, "dL_dcolor must be [H, W, 3]");
    TORCH_CHECK(model.is_valid(), "GaussianModel is not valid");
    const int64_t n = model.num_gaussians();
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dL_dcolor.device());
    BackwardOutput g;
    if (n == 0) {  // rasterizer.cpp:130-139
        g.dL_dpositions = torch::zeros({0, 3}, f32);
        g.dL_drotations = torch::zeros({0, 4}, f32);
        g.dL_dscales = torch::zeros({0, 3}, f32);
        g.dL_dopacities = torch::zeros({0, 1}, f32);
        g.dL_dsh_coeffs = torch::zeros_like(model.sh_coeffs);
        g.dL_dmeans_2d = torch::zeros({0, 2}, f32);
        return g;
    }
    cugs_handle_t* h = handle_for(dL_dcolor);
    void* stream = current_stream(dL_dcolor);
    const int degree = std::min(settings.active_sh_degree, model.max_sh_degree());
    const auto pos = f32c(model.positions), rot = f32c(model.rotations), scl = f32c(model.scales),
               opa = f32c(model.opacities), sh = f32c(model.sh_coeffs);
    const cugs_view_t v = make_view(camera, settings.background, degree, (int)sh.size(2), settings.scale_modifier);
    const auto dL = f32c(dL_dcolor);
    g.dL_dpositions = torch::empty({n, 3}, f32);
    g.dL_drotations = torch::empty({n, 4}, f32);
    g.dL_dscales = torch::empty({n, 3}, f32);
    g.dL_dopacities = torch::empty({n, 1}, f32);
    g.dL_dsh_coeffs = torch::empty_like(sh);
    g.dL_dmeans_2d = torch::empty({n, 2}, f32);
    torch::Tensor ws = recall_frame(r.means_2d);
    if (ws.defined()) {  // the frame came from render() above: packed records are in its workspace
        CUGS_CALL(h, cugs_b200_render_backward(
                         h, stream, n, &v, fp(pos), fp(rot), fp(scl), fp(opa), fp(sh), fp(r.means_2d), fp(r.cov_2d_inv),
                         ip(r.radii), fp(r.rgb), fp(r.opacities_act), ip(r.gaussian_indices), ip(r.tile_ranges),
                         fp(r.final_T), ip(r.n_contrib), fp(dL), fp(g.dL_dpositions), fp(g.dL_drotations),
                         fp(g.dL_dscales), fp(g.dL_dopacities), fp(g.dL_dsh_coeffs), fp(g.dL_dmeans_2d), nullptr, nullptr,
                         nullptr, /*touch_mask=*/nullptr, /*flags=*/0, ws.data_ptr(), (size_t)ws.numel()));
        return g;
    }
    // a RenderOutput assembled by the caller: go through the two stage functions (rasterizer.cpp:146-176)
    auto rb = rasterize_backward(dL, r.means_2d, r.cov_2d_inv, r.rgb, r.opacities_act, r.tile_ranges, r.gaussian_indices,
                                 r.final_T, r.n_contrib, camera.width, camera.height, settings.background, (int)n);
    auto pb = project_backward(rb.dL_dmeans_2d, rb.dL_dcov_2d_inv, rb.dL_drgb, rb.dL_dopacity_act, pos, rot, scl, opa, sh,
                               r.radii, camera, degree, settings.scale_modifier);
    g.dL_dpositions = pb.dL_dpositions;
    g.dL_drotations = pb.dL_drotations;
    g.dL_dscales = pb.dL_dscales;
    g.dL_dopacities = pb.dL_dopacities;
    g.dL_dsh_coeffs = pb.dL_dsh_coeffs;
    g.dL_dmeans_2d = rb.dL_dmeans_2d;
    return g;
}

// =====================================================================================================
// stage functions
// =====================================================================================================
ProjectionOutput project_gaussians(const torch::Tensor& positions, const torch::Tensor& rotations,
                                   const torch::Tensor& scales, const torch::Tensor& opacities,
                                   const torch::Tensor& sh_coeffs, const CameraInfo& camera, int active_sh_degree,
                                   float scale_modifier) {
    TORCH_CHECK(positions.is_cuda(), "positions must be on CUDA");
    TORCH_CHECK(positions.dim() == 2 && positions.size(1) == 3, "positions must be [N,3]");
    TORCH_CHECK(sh_coeffs.dim() == 3 && sh_coeffs.size(1) == 3, "sh_coeffs must be [N, 3, C]");
    const int64_t n = positions.size(0);
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(positions.device());
    const auto i32 = f32.dtype(torch::kInt32);
    ProjectionOutput o;
    o.means_2d = torch::empty({n, 2}, f32);
    o.depths = torch::empty({n}, f32);
    o.cov_2d_inv = torch::empty({n, 3}, f32);
    o.radii = torch::empty({n}, i32);
    o.tiles_touched = torch::empty({n}, i32);
    o.rgb = torch::empty({n, 3}, f32);
    o.opacities_act = torch::empty({n}, f32);
    if (n == 0) return o;
    cugs_handle_t* h = handle_for(positions);
    const auto pos = f32c(positions), rot = f32c(rotations), scl = f32c(scales), opa = f32c(opacities),
               sh = f32c(sh_coeffs);
    const cugs_view_t v = make_view(camera, nullptr, active_sh_degree, (int)sh.size(2), scale_modifier);
    CUGS_CALL(h, cugs_b200_preprocess_fwd(h, current_stream(positions), n, &v, fp(pos), fp(rot), fp(scl), fp(opa), fp(sh),
                                          fp(o.means_2d), fp(o.depths), fp(o.cov_2d_inv), ip(o.radii),
                                          ip(o.tiles_touched), fp(o.rgb), fp(o.opacities_act), nullptr, nullptr));
    return o;
}

SortingOutput sort_gaussians(const torch::Tensor& means_2d, const torch::Tensor& depths, const torch::Tensor& radii,
                             const torch::Tensor& tiles_touched, int img_w, int img_h) {
    TORCH_CHECK(means_2d.is_cuda(), "means_2d must be on CUDA");
    const int64_t n = means_2d.size(0);
    const int nt = tiles_of(img_w, img_h);
    const auto i32 = torch::TensorOptions().dtype(torch::kInt32).device(means_2d.device());
    const auto i64 = i32.dtype(torch::kInt64);
    const auto u8 = i32.dtype(torch::kUInt8);
    SortingOutput o;
    o.total_pairs = 0;
    o.gaussian_keys_sorted = torch::empty({0}, i64);
    o.gaussian_values_sorted = torch::empty({0}, i32);
    o.tile_ranges = torch::zeros({nt, 2}, i32);
    if (n == 0) return o;
    cugs_handle_t* h = handle_for(means_2d);
    void* stream = current_stream(means_2d);
    const auto tiles = tiles_touched.contiguous().to(torch::kInt32);
    auto offsets = torch::empty({n}, i32);
    auto scan_tmp = torch::empty({(int64_t)cugs_b200_scan_temp_bytes(n)}, u8);
    int64_t P = 0;
    CUGS_CALL(h, cugs_b200_scan(h, stream, n, ip(tiles), ip(offsets), nullptr, &P, scan_tmp.data_ptr(),
                                (size_t)scan_tmp.numel()));  // the one blocking read (sorting.cu:146)
    if (P == 0) return o;
    auto keys = torch::empty({P}, i64), keys_sorted = torch::empty({P}, i64);
    auto vals = torch::empty({P}, i32), vals_sorted = torch::empty({P}, i32);
    const auto m2d = f32c(means_2d), dep = f32c(depths);
    const auto rad = radii.contiguous().to(torch::kInt32);
    CUGS_CALL(h, cugs_b200_duplicate_with_keys(h, stream, n, img_w, img_h, fp(m2d), fp(dep), ip(rad), ip(tiles),
                                               ip(offsets), P, reinterpret_cast<uint64_t*>(keys.data_ptr<int64_t>()),
                                               ip(vals)));
    int tile_bits = 0;
    while ((1 << tile_bits) < nt) ++tile_bits;
    auto sort_tmp = torch::empty({(int64_t)cugs_b200_sort_temp_bytes(P)}, u8);
    CUGS_CALL(h, cugs_b200_sort_pairs(h, stream, P, 32, tile_bits, reinterpret_cast<uint64_t*>(keys.data_ptr<int64_t>()),
                                      ip(vals), reinterpret_cast<uint64_t*>(keys_sorted.data_ptr<int64_t>()),
                                      ip(vals_sorted), sort_tmp.data_ptr(), (size_t)sort_tmp.numel()));
    o.tile_ranges = torch::empty({nt, 2}, i32);
    CUGS_CALL(h, cugs_b200_tile_ranges(h, stream, P, reinterpret_cast<const uint64_t*>(keys_sorted.data_ptr<int64_t>()),
                                       nt, ip(o.tile_ranges)));
    o.gaussian_keys_sorted = keys_sorted;
    o.gaussian_values_sorted = vals_sorted;
    o.total_pairs = (int)P;
    return o;
}

ForwardOutput rasterize_forward(const torch::Tensor& means_2d, const torch::Tensor& cov_2d_inv, const torch::Tensor& rgb,
                                const torch::Tensor& opacities, const torch::Tensor& tile_ranges,
                                const torch::Tensor& gaussian_indices, int img_w, int img_h, const float background[3]) {
    TORCH_CHECK(means_2d.is_cuda(), "means_2d must be on CUDA");
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(means_2d.device());
    ForwardOutput o;
    o.color = torch::empty({img_h, img_w, 3}, f32);
    o.final_T = torch::empty({img_h, img_w}, f32);
    o.n_contrib = torch::empty({img_h, img_w}, f32.dtype(torch::kInt32));
    if (img_w == 0 || img_h == 0) return o;
    cugs_handle_t* h = handle_for(means_2d);
    CameraInfo cam;
    cam.width = img_w;
    cam.height = img_h;
    const cugs_view_t v = make_view(cam, background, 0, 1, 1.0f);
    const auto m = f32c(means_2d), c = f32c(cov_2d_inv), col = f32c(rgb), op = f32c(opacities);
    const auto tr = tile_ranges.contiguous(), gi = gaussian_indices.contiguous();
    CUGS_CALL(h, cugs_b200_blend_fwd(h, current_stream(means_2d), &v, ip(tr), ip(gi), fp(m), fp(c), fp(col), fp(op),
                                     nullptr, fp(o.color), fp(o.final_T), ip(o.n_contrib)));
    return o;
}

RasterizeBackwardOutput rasterize_backward(const torch::Tensor& dL_dcolor, const torch::Tensor& means_2d,
                                           const torch::Tensor& cov_2d_inv, const torch::Tensor& rgb,
                                           const torch::Tensor& opacities, const torch::Tensor& tile_ranges,
                                           const torch::Tensor& gaussian_indices, const torch::Tensor& final_T,
                                           const torch::Tensor& n_contrib, int img_w, int img_h,
                                           const float background[3], int n_gaussians) {
    TORCH_CHECK(dL_dcolor.is_cuda(), "dL_dcolor must be on CUDA");
    const int64_t n = n_gaussians;
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dL_dcolor.device());
    RasterizeBackwardOutput o;
    o.dL_drgb = torch::zeros({n, 3}, f32);
    o.dL_dopacity_act = torch::zeros({n}, f32);
    o.dL_dmeans_2d = torch::zeros({n, 2}, f32);
    o.dL_dcov_2d_inv = torch::zeros({n, 3}, f32);
    if (n == 0 || img_w == 0 || img_h == 0) return o;
    cugs_handle_t* h = handle_for(dL_dcolor);
    CameraInfo cam;
    cam.width = img_w;
    cam.height = img_h;
    const cugs_view_t v = make_view(cam, background, 0, 1, 1.0f);
    auto acc = torch::empty({n, 12}, f32);
    const auto dL = f32c(dL_dcolor), m = f32c(means_2d), c = f32c(cov_2d_inv), col = f32c(rgb), op = f32c(opacities),
               fT = f32c(final_T);
    const auto tr = tile_ranges.contiguous(), gi = gaussian_indices.contiguous(), nc = n_contrib.contiguous();
    CUGS_CALL(h, cugs_b200_blend_bwd(h, current_stream(dL_dcolor), n, &v, ip(tr), ip(gi), fp(m), fp(c), fp(col), fp(op),
                                     nullptr, fp(dL), fp(fT), ip(nc), fp(o.dL_drgb), fp(o.dL_dopacity_act),
                                     fp(o.dL_dmeans_2d), fp(o.dL_dcov_2d_inv), fp(acc)));
    return o;
}

ProjectionBackwardOutput project_backward(const torch::Tensor& dL_dmeans_2d, const torch::Tensor& dL_dcov_2d_inv,
                                          const torch::Tensor& dL_drgb, const torch::Tensor& dL_dopacity_act,
                                          const torch::Tensor& positions, const torch::Tensor& rotations,
                                          const torch::Tensor& scales, const torch::Tensor& opacities,
                                          const torch::Tensor& sh_coeffs, const torch::Tensor& radii,
                                          const CameraInfo& camera, int active_sh_degree, float scale_modifier) {
    TORCH_CHECK(positions.is_cuda(), "positions must be on CUDA");
    const int64_t n = positions.size(0);
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(positions.device());
    const auto pos = f32c(positions), rot = f32c(rotations), scl = f32c(scales), opa = f32c(opacities),
               sh = f32c(sh_coeffs);
    ProjectionBackwardOutput o;
    o.dL_dpositions = torch::empty({n, 3}, f32);
    o.dL_drotations = torch::empty({n, 4}, f32);
    o.dL_dscales = torch::empty({n, 3}, f32);
    o.dL_dopacities = torch::empty({n, 1}, f32);
    o.dL_dsh_coeffs = torch::empty_like(sh);
    if (n == 0) return o;
    cugs_handle_t* h = handle_for(positions);
    // the ReLU gate needs the forward colour; this stage function does not receive it -> recompute
    const auto fwd = project_gaussians(pos, rot, scl, opa, sh, camera, active_sh_degree, scale_modifier);
    const cugs_view_t v = make_view(camera, nullptr, active_sh_degree, (int)sh.size(2), scale_modifier);
    const auto dm = f32c(dL_dmeans_2d), dc = f32c(dL_dcov_2d_inv), dr = f32c(dL_drgb), dop = f32c(dL_dopacity_act);
    const auto rad = radii.contiguous().to(torch::kInt32);
    CUGS_CALL(h, cugs_b200_preprocess_bwd(h, current_stream(positions), n, &v, fp(pos), fp(rot), fp(scl), fp(opa), fp(sh),
                                          ip(rad), fp(fwd.rgb), fp(dm), fp(dc), fp(dr), fp(dop), fp(o.dL_dpositions),
                                          fp(o.dL_drotations), fp(o.dL_dscales), fp(o.dL_dopacities), fp(o.dL_dsh_coeffs),
                                          nullptr, nullptr, nullptr));
    return o;
}

// =====================================================================================================
// core/sh.hpp, core/sh_backward.hpp
// =====================================================================================================
torch::Tensor evaluate_sh_cuda(int degree, const torch::Tensor& sh_coeffs, const torch::Tensor& directions) {
    TORCH_CHECK(degree >= 0 && degree <= 3, "SH degree must be 0..3, got ", degree);
    TORCH_CHECK(sh_coeffs.is_cuda(), "sh_coeffs must be on CUDA device");
    TORCH_CHECK(directions.is_cuda(), "directions must be on CUDA device");
    TORCH_CHECK(sh_coeffs.dim() == 3 && sh_coeffs.size(1) == 3, "sh_coeffs must be [N, 3, C]");
    TORCH_CHECK(directions.dim() == 2 && directions.size(1) == 3, "directions must be [N, 3]");
    TORCH_CHECK(sh_coeffs.size(0) == directions.size(0), "Batch size mismatch");
    const int need = (degree + 1) * (degree + 1);
    TORCH_CHECK(sh_coeffs.size(2) >= need, "Need at least ", need, " coefficients for degree ", degree);
    const int64_t n = sh_coeffs.size(0);
    auto out = torch::empty({n, 3}, torch::TensorOptions().dtype(torch::kFloat32).device(sh_coeffs.device()));
    if (n == 0) return out;
    cugs_handle_t* h = handle_for(sh_coeffs);
    const auto sh = f32c(sh_coeffs), d = f32c(directions);
    CUGS_CALL(h, cugs_b200_sh_forward(h, current_stream(sh_coeffs), n, degree, (int)sh.size(2), fp(sh), fp(d), fp(out)));
    return out;
}

torch::Tensor evaluate_sh_backward_cuda(int degree, const torch::Tensor& sh_coeffs, const torch::Tensor& directions,
                                        const torch::Tensor& dL_dcolor) {
    TORCH_CHECK(degree >= 0 && degree <= 3, "SH degree must be 0..3, got ", degree);
    TORCH_CHECK(sh_coeffs.is_cuda() && directions.is_cuda() && dL_dcolor.is_cuda(), "inputs must be on CUDA device");
    TORCH_CHECK(sh_coeffs.dim() == 3 && sh_coeffs.size(1) == 3, "sh_coeffs must be [N, 3, C]");
    TORCH_CHECK(directions.dim() == 2 && directions.size(1) == 3, "directions must be [N, 3]");
    TORCH_CHECK(dL_dcolor.dim() == 2 && dL_dcolor.size(1) == 3, "dL_dcolor must be [N, 3]");
    const int64_t n = sh_coeffs.size(0);
    const auto sh = f32c(sh_coeffs), d = f32c(directions), g = f32c(dL_dcolor);
    auto out = torch::empty_like(sh);
    if (n == 0) return out;
    cugs_handle_t* h = handle_for(sh_coeffs);
    CUGS_CALL(h, cugs_b200_sh_backward(h, current_stream(sh_coeffs), n, degree, (int)sh.size(2), fp(sh), fp(d), fp(g),
                                       fp(out)));
    return out;
}

// =====================================================================================================
// training/loss.hpp — value from the fused kernel, gradient through a custom autograd node so that
// `loss.backward()` (training/trainer.cpp:214-217) keeps working
// =====================================================================================================
namespace {

void validate_pair(const torch::Tensor& rendered, const torch::Tensor& target) {
    for (const auto* t : {&rendered, &target}) {
        TORCH_CHECK(t->dim() == 3, "image must be 3-dimensional [H, W, 3], got ", t->dim(), " dims");
        TORCH_CHECK(t->size(2) == 3, "image must have 3 channels, got ", t->size(2));
        TORCH_CHECK(t->dtype() == torch::kFloat32, "image must be float32");
        TORCH_CHECK(t->is_cuda(), "image must be on a CUDA device");
    }
    TORCH_CHECK(rendered.sizes() == target.sizes(), "rendered and target must have the same shape");
}

// scalars3 = {loss, l1, mean ssim}; optional gradient and SSIM map
torch::Tensor fused_loss(const torch::Tensor& rendered, const torch::Tensor& target, float lambda, torch::Tensor* grad,
                         torch::Tensor* ssim_map, int window_size = 11) {
    validate_pair(rendered, target);
    cugs_handle_t* h = handle_for(rendered);
    const int H = (int)rendered.size(0), W = (int)rendered.size(1);
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(rendered.device());
    auto ws = torch::empty({(int64_t)cugs_b200_loss_workspace_bytes(W, H)}, f32.dtype(torch::kUInt8));
    auto scalars = torch::empty({3}, f32);
    const auto x = rendered.contiguous(), y = target.contiguous();
    if (grad) *grad = torch::empty({H, W, 3}, f32);
    if (ssim_map) *ssim_map = torch::empty({H, W}, f32);
    CUGS_CALL(h, cugs_b200_loss_l1_ssim(h, current_stream(rendered), W, H, lambda, window_size, fp(x), fp(y),
                                        grad ? fp(*grad) : nullptr,
                                        fp(scalars), ws.data_ptr(), (size_t)ws.numel(),
                                        ssim_map ? fp(*ssim_map) : nullptr));
    return scalars;
}

struct FusedLossFn : public torch::autograd::Function<FusedLossFn> {
    static torch::Tensor forward(torch::autograd::AutogradContext* ctx, const torch::Tensor& rendered,
                                 const torch::Tensor& target, double lambda, int64_t which, int64_t window_size) {
        torch::Tensor grad;
        const auto scalars =
            fused_loss(rendered.detach(), target.detach(), (float)lambda, &grad, nullptr, (int)window_size);
        ctx->save_for_backward({grad});
        return scalars[which].clone();
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                   torch::autograd::variable_list grad_out) {
        const auto saved = ctx->get_saved_variables();
        return {saved[0] * grad_out[0], torch::Tensor(), torch::Tensor(), torch::Tensor(), torch::Tensor()};
    }
};

}  // namespace

torch::Tensor l1_loss(const torch::Tensor& rendered, const torch::Tensor& target) {
    validate_pair(rendered, target);
    if (rendered.requires_grad()) return FusedLossFn::apply(rendered, target, 0.0, 0, 11);  // lambda 0: loss == l1
    return fused_loss(rendered, target, 0.0f, nullptr, nullptr)[1];
}

torch::Tensor ssim(const torch::Tensor& rendered, const torch::Tensor& target, int window_size) {
    validate_pair(rendered, target);
    TORCH_CHECK(window_size % 2 == 1, "window_size must be odd, got ", window_size);
    TORCH_CHECK(window_size >= 3, "window_size must be >= 3, got ", window_size);
    torch::Tensor map;
    fused_loss(rendered, target, 1.0f, nullptr, &map, window_size);
    return map;  // [H, W], channel mean of the SSIM map (loss.cpp:123)
}

torch::Tensor ssim_loss(const torch::Tensor& rendered, const torch::Tensor& target, int window_size) {
    TORCH_CHECK(window_size % 2 == 1, "window_size must be odd, got ", window_size);
    TORCH_CHECK(window_size >= 3, "window_size must be >= 3, got ", window_size);
    if (rendered.requires_grad())
        return FusedLossFn::apply(rendered, target, 1.0, 0, window_size);  // lambda 1: loss == 1 - ssim
    return 1.0f - fused_loss(rendered, target, 1.0f, nullptr, nullptr, window_size)[2];
}

torch::Tensor combined_loss(const torch::Tensor& rendered, const torch::Tensor& target, float lambda_) {
    validate_pair(rendered, target);
    if (rendered.requires_grad()) return FusedLossFn::apply(rendered, target, (double)lambda_, 0, 11);
    return fused_loss(rendered, target, lambda_, nullptr, nullptr)[0];
}

// =====================================================================================================
// optimizer/fused_adam.hpp — same class, one multi-tensor launch per step
// =====================================================================================================
FusedAdam::FusedAdam(GaussianModel& model, const AdamConfig& config) : model_(model), config_(config), step_count_(0) {
    torch::Tensor* params[kNumGroups] = {&model_.positions, &model_.sh_coeffs, &model_.opacities, &model_.scales,
                                         &model_.rotations};
    for (int i = 0; i < kNumGroups; ++i) {
        TORCH_CHECK(params[i]->is_cuda(), "FusedAdam: param must be on CUDA");
        params[i]->requires_grad_(true);
        m_[i] = torch::zeros_like(*params[i]);
        v_[i] = torch::zeros_like(*params[i]);
        grads_[i] = torch::Tensor();
    }
    learning_rates_ = {config.position_lr_config.lr_init, config.lr_sh_coeffs, config.lr_opacities, config.lr_scales,
                       config.lr_rotations};
}

void FusedAdam::apply_gradients(const BackwardOutput& grads) {
    grads_ = {grads.dL_dpositions, grads.dL_dsh_coeffs, grads.dL_dopacities, grads.dL_dscales, grads.dL_drotations};
}

void FusedAdam::update_lr(int step) { learning_rates_[0] = position_lr(step, config_.position_lr_config); }

void FusedAdam::zero_grad() {
    for (auto& g : grads_) g = torch::Tensor();
    for (torch::Tensor* p : {&model_.positions, &model_.sh_coeffs, &model_.opacities, &model_.scales, &model_.rotations})
        if (p->grad().defined()) p->mutable_grad().zero_();
}

void FusedAdam::step() {
    ++step_count_;
    const double bc1 = 1.0 / (1.0 - std::pow((double)config_.beta1, step_count_));  // fused_adam.cu:145-149
    const double bc2 = 1.0 / (1.0 - std::pow((double)config_.beta2, step_count_));
    torch::Tensor* params[kNumGroups] = {&model_.positions, &model_.sh_coeffs, &model_.opacities, &model_.scales,
                                         &model_.rotations};
    float* p[5] = {};
    const float* g[5] = {};
    float* m[5] = {};
    float* v[5] = {};
    int64_t counts[5] = {};
    float lr[5] = {};
    std::array<torch::Tensor, kNumGroups> keep;
    for (int i = 0; i < kNumGroups; ++i) {
        if (!grads_[i].defined()) continue;  // fused_adam.cu:157
        TORCH_CHECK(params[i]->is_contiguous() && params[i]->dtype() == torch::kFloat32,
                    "FusedAdam: params must be contiguous float32");
        TORCH_CHECK(grads_[i].numel() == params[i]->numel(), "FusedAdam: param/grad size mismatch");
        keep[i] = f32c(grads_[i]);
        p[i] = params[i]->data_ptr<float>();
        g[i] = keep[i].data_ptr<float>();
        m[i] = m_[i].data_ptr<float>();
        v[i] = v_[i].data_ptr<float>();
        counts[i] = params[i]->numel();
        lr[i] = learning_rates_[i];
    }
    cugs_handle_t* h = handle_for(model_.positions);
    CUGS_CALL(h, cugs_b200_adam_step(h, current_stream(model_.positions), p, g, m, v, counts, lr, config_.beta1,
                                     config_.beta2, config_.eps, (float)bc1, (float)bc2, 1.0f));
}

float FusedAdam::get_lr(ParamGroup group) const { return learning_rates_[static_cast<int>(group)]; }

void FusedAdam::launch_kernel(torch::Tensor& param, const torch::Tensor& grad, torch::Tensor& m, torch::Tensor& v,
                              float lr, float bc1, float bc2) {  // single-group form of step()
    float* p[5] = {param.data_ptr<float>(), nullptr, nullptr, nullptr, nullptr};
    const auto gc = f32c(grad);
    const float* g[5] = {gc.data_ptr<float>(), nullptr, nullptr, nullptr, nullptr};
    float* mm[5] = {m.data_ptr<float>(), nullptr, nullptr, nullptr, nullptr};
    float* vv[5] = {v.data_ptr<float>(), nullptr, nullptr, nullptr, nullptr};
    const int64_t counts[5] = {param.numel(), 0, 0, 0, 0};
    const float lrs[5] = {lr, 0, 0, 0, 0};
    cugs_handle_t* h = handle_for(param);
    CUGS_CALL(h, cugs_b200_adam_step(h, current_stream(param), p, g, mm, vv, counts, lrs, config_.beta1, config_.beta2,
                                     config_.eps, bc1, bc2, 1.0f));
}

// =====================================================================================================
// optimizer/densification.hpp -- DensificationController on the C ABI (densify = two kernels around the
// host policy instead of the mask-index / torch::cat chain of optimizer/densification.cpp:94-329)
// =====================================================================================================
DensificationController::DensificationController(const DensificationConfig& config, float scene_extent)
    : config_(config), scene_extent_(scene_extent) {}

bool DensificationController::should_densify(int step) const {
    const auto& c = config_;
    return step % c.densify_every == 0 && step >= c.densify_from && step <= c.densify_until;
}

bool DensificationController::should_reset_opacity(int step) const {
    const auto& c = config_;
    return c.opacity_reset_every > 0 && step % c.opacity_reset_every == 0 && step >= c.densify_from;
}

void DensificationController::reset_accumulators(int64_t n) {
    const auto opts = torch::TensorOptions().dtype(torch::kFloat32).device(
        grad_accum_.defined() ? grad_accum_.device() : torch::Device(torch::kCUDA));
    grad_accum_ = torch::zeros({n}, opts);
    grad_count_ = torch::zeros({n}, opts);
    max_radii_2d_ = torch::zeros({n}, opts);
}

void DensificationController::accumulate_gradients(const torch::Tensor& dL_dmeans_2d, const torch::Tensor& radii) {
    torch::NoGradGuard no_grad;
    const int64_t n = dL_dmeans_2d.size(0);
    if (!grad_accum_.defined() || grad_accum_.size(0) != n) {  // lazy (re-)initialisation
        grad_accum_ = torch::Tensor();
        grad_accum_ = torch::zeros({n}, dL_dmeans_2d.options().dtype(torch::kFloat32));
        grad_count_ = torch::zeros_like(grad_accum_);
        max_radii_2d_ = torch::zeros_like(grad_accum_);
    }
    if (n == 0) return;
    const auto g = f32c(dL_dmeans_2d);
    const auto r = radii.contiguous().to(torch::kInt32);
    cugs_handle_t* h = handle_for(g);
    CUGS_CALL(h, cugs_b200_accumulate_stats(h, current_stream(g), n, fp(g), ip(r), fp(grad_accum_), fp(grad_count_),
                                            fp(max_radii_2d_)));
}

void DensificationController::reset_opacity(GaussianModel& model) {
    torch::NoGradGuard no_grad;
    model.opacities.fill_(std::log(0.01f / 0.99f));  // inverse_sigmoid(0.01)
}

DensificationStats DensificationController::densify(GaussianModel& model, int step) {
    torch::NoGradGuard no_grad;
    DensificationStats stats;
    const int64_t n = model.num_gaussians();
    stats.num_before = stats.num_after = static_cast<int>(n);
    if (n == 0) return stats;
    if (!grad_accum_.defined() || grad_accum_.size(0) != n) {
        grad_accum_ = torch::zeros({n}, model.positions.options());
        grad_count_ = torch::zeros_like(grad_accum_);
        max_radii_2d_ = torch::zeros_like(grad_accum_);
    }
    cugs_handle_t* h = handle_for(model.positions);
    void* stream = current_stream(model.positions);
    const auto u8 = torch::TensorOptions().dtype(torch::kUInt8).device(model.positions.device());

    cugs_densify_config_t cfg{};
    cfg.grad_threshold = config_.grad_threshold;
    cfg.size_threshold = config_.percent_dense * scene_extent_;
    cfg.opacity_threshold = config_.opacity_threshold;
    cfg.apply_size_pruning = (config_.opacity_reset_every > 0 && step > config_.opacity_reset_every) ? 1 : 0;
    cfg.max_screen_size = static_cast<float>(config_.max_screen_size);
    cfg.ws_threshold = 0.1f * scene_extent_;

    auto pos = f32c(model.positions), sh = f32c(model.sh_coeffs), opa = f32c(model.opacities),
         scl = f32c(model.scales), rot = f32c(model.rotations);
    auto flags = torch::empty({n}, u8);
    auto temp = torch::empty({static_cast<int64_t>(cugs_b200_densify_temp_bytes(n))}, u8);
    int64_t counts[3] = {0, 0, 0};
    CUGS_CALL(h, cugs_b200_densify_classify(h, stream, n, fp(scl), fp(opa), fp(grad_accum_), fp(grad_count_),
                                            fp(max_radii_2d_), &cfg, flags.data_ptr<uint8_t>(), counts,
                                            temp.data_ptr(), temp.numel()));
    int64_t kept = counts[0], n_clone = counts[1], n_split = counts[2];

    // budget caps (max_gaussians): the rarely taken host policy, on the flag array. Candidates are ranked
    // by their average gradient and only the best `budget` keep their bit.
    auto cap = [&](int bit, int64_t budget) {
        auto mask = flags.bitwise_and(bit).ne(0);
        flags.bitwise_and_(static_cast<uint8_t>(~bit & 0xFF));
        if (budget <= 0) return;
        auto avg = grad_accum_ / grad_count_.clamp_min(1);
        auto idx = std::get<1>(avg.masked_fill(~mask, -1.0f).topk(budget));
        flags.index_put_({idx}, flags.index({idx}).bitwise_or(bit));
    };
    bool capped = false;
    if (config_.max_gaussians > 0) {
        const int64_t clone_budget = config_.max_gaussians - n;
        if (n_clone > 0 && n_clone > clone_budget) {
            cap(CUGS_DENSIFY_CLONE, clone_budget);
            n_clone = std::max<int64_t>(std::min(n_clone, clone_budget), 0);
            capped = true;
        }
        if (n_split > 0) {
            const int64_t split_budget = (config_.max_gaussians - (n + n_clone)) / 2;
            if (n_split > split_budget) {
                cap(CUGS_DENSIFY_SPLIT, split_budget);
                n_split = std::max<int64_t>(std::min(n_split, split_budget), 0);
                capped = true;
            }
        }
    }
    if (capped)
        kept = (flags.bitwise_and(CUGS_DENSIFY_KEEP).ne(0) & flags.bitwise_and(CUGS_DENSIFY_SPLIT).eq(0)).sum().item<int64_t>();

    const int64_t n_out = kept + n_clone + 2 * n_split;
    stats.num_cloned = static_cast<int>(n_clone);
    stats.num_split = static_cast<int>(n_split);
    stats.num_pruned = static_cast<int>(n + n_clone + 2 * n_split - n_out);
    stats.num_after = static_cast<int>(n_out);
    if (n_clone > 0 || n_split > 0 || n_out != n) {
        const int C = static_cast<int>(sh.size(2));
        const auto f32 = model.positions.options().dtype(torch::kFloat32);
        torch::Tensor dst[5] = {torch::empty({n_out, 3}, f32), torch::empty({n_out, 3, C}, f32),
                                torch::empty({n_out, 1}, f32), torch::empty({n_out, 3}, f32),
                                torch::empty({n_out, 4}, f32)};
        const float* src_p[5] = {fp(pos), fp(sh), fp(opa), fp(scl), fp(rot)};
        float* dst_p[5] = {fp(dst[0]), fp(dst[1]), fp(dst[2]), fp(dst[3]), fp(dst[4])};
        static uint64_t call_counter = 0;  // a different Philox stream for every densification
        const uint64_t seed = 0xD5171F00ull + (static_cast<uint64_t>(step) << 20) + (call_counter++);
        CUGS_CALL(h, cugs_b200_densify_apply(h, stream, n, n_out, C, flags.data_ptr<uint8_t>(), src_p, dst_p, nullptr,
                                             nullptr, nullptr, nullptr, seed, nullptr, temp.data_ptr(), temp.numel()));
        model.positions = dst[0];
        model.sh_coeffs = dst[1];
        model.opacities = dst[2];
        model.scales = dst[3];
        model.rotations = dst[4];
    }
    reset_accumulators(n_out);
    return stats;
}

// =====================================================================================================
// optimizer/mcmc_densification.hpp -- MCMCController on the C ABI
// =====================================================================================================
MCMCController::MCMCController(const MCMCConfig& config, float scene_extent)
    : config_(config), scene_extent_(scene_extent) {}

bool MCMCController::should_relocate(int step) const {
    const auto& c = config_;
    return step % c.relocate_every == 0 && step >= c.relocate_from && step <= c.relocate_until;
}

float MCMCController::noise_lr(int step) const {  // log-linear decay, in float like the reference
    const auto& c = config_;
    if (step <= 0) return c.noise_lr_init;
    if (step >= c.noise_lr_max_steps) return c.noise_lr_final;
    const float t = static_cast<float>(step) / static_cast<float>(c.noise_lr_max_steps);
    return c.noise_lr_init * std::exp(t * std::log(c.noise_lr_final / c.noise_lr_init));
}

namespace {
constexpr uint64_t kMcmcSeed = 0x5EEDull;  // every rank of a view-parallel run must draw the same numbers
void check_plain(const GaussianModel& m) {
    TORCH_CHECK(m.positions.is_cuda() && m.positions.is_contiguous() && m.positions.scalar_type() == torch::kFloat32 &&
                    m.sh_coeffs.is_contiguous() && m.opacities.is_contiguous() && m.scales.is_contiguous() &&
                    m.rotations.is_contiguous(),
                "MCMC ops update the model in place: tensors must be contiguous float32 on CUDA");
}
}  // namespace

MCMCStats MCMCController::relocate(GaussianModel& model, int step) {
    torch::NoGradGuard no_grad;
    MCMCStats stats;
    const int64_t n = model.num_gaussians();
    stats.num_total = static_cast<int>(n);
    if (n == 0) return stats;
    check_plain(model);
    cugs_handle_t* h = handle_for(model.positions);
    const int64_t max_relocate = static_cast<int>(config_.relocate_cap * n);
    auto temp = torch::empty({static_cast<int64_t>(cugs_b200_mcmc_relocate_temp_bytes(n))},
                             torch::TensorOptions().dtype(torch::kUInt8).device(model.positions.device()));
    int64_t counts[2] = {0, 0};
    CUGS_CALL(h, cugs_b200_mcmc_relocate(h, current_stream(model.positions), n, static_cast<int>(model.sh_coeffs.size(2)),
                                         fp(model.positions), fp(model.sh_coeffs), fp(model.opacities), fp(model.scales),
                                         fp(model.rotations), config_.dead_opacity_threshold, max_relocate, scene_extent_,
                                         kMcmcSeed, static_cast<uint32_t>(step), nullptr, nullptr, counts, temp.data_ptr(),
                                         temp.numel()));
    stats.num_dead = static_cast<int>(counts[0]);
    stats.num_relocated = static_cast<int>(counts[1]);
    return stats;
}

void MCMCController::inject_noise(GaussianModel& model, int step) {
    torch::NoGradGuard no_grad;
    const int64_t n = model.num_gaussians();
    if (n == 0) return;
    check_plain(model);
    cugs_handle_t* h = handle_for(model.positions);
    CUGS_CALL(h, cugs_b200_mcmc_inject_noise(h, current_stream(model.positions), n, fp(model.positions), fp(model.scales),
                                             fp(model.opacities), noise_lr(step), config_.noise_gate_k,
                                             config_.noise_gate_t, kMcmcSeed, static_cast<uint32_t>(step), nullptr));
}

float MCMCController::compute_regularization(const GaussianModel& model, torch::Tensor& reg_dL_dopacities,
                                             torch::Tensor& reg_dL_dscales) {
    // loss = lambda_o mean(sigmoid(opacity)) + lambda_s mean(exp(scale)); its gradient in closed form instead of
    // two autograd graphs. (FusedAdam can also add it inside its own launch: cugs_b200_adam_step_mcmc.)
    torch::NoGradGuard no_grad;
    const auto sg = torch::sigmoid(model.opacities);
    const auto es = torch::exp(model.scales);
    reg_dL_dopacities = sg * (1.0f - sg) * (config_.lambda_opacity / static_cast<float>(sg.numel()));
    reg_dL_dscales = es * (config_.lambda_scale / static_cast<float>(es.numel()));
    return (config_.lambda_opacity * sg.mean() + config_.lambda_scale * es.mean()).item<float>();
}

}  // namespace cugs
