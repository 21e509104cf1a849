// oracle/ref_harness.cpp — TEST INFRASTRUCTURE (not product code).
//
// A thin pybind11 module around the UNMODIFIED reference hot path, compiled from the sources
// where they lie under /root/reference (see oracle/Makefile.ref; nothing is copied into this
// repo). It exists for two reasons only:
//   1. bit oracle: radii / tiles_touched / tile keys / sort order / tile ranges of the
//      reference's own CUDA kernels, run on the same B200 as the new kernels;
//   2. `bench.py --impl reference`: the reference's own render()/render_backward() timed
//      through its public API (rasterizer.hpp:57-60, :88-93).
// The product (cuda_gaussian_splatting_b200/) never imports this module.
//
// Camera is passed as 18 floats: W, H, fx, fy, cx, cy, R(row-major 3x3 world->camera), t(3).

#include <torch/extension.h>

#include <vector>

#include "core/gaussian.hpp"
#include "core/sh.hpp"
#include "core/sh_backward.hpp"
#include "core/types.hpp"
#include "optimizer/fused_adam.hpp"
#define private public  // test harness only: read DensificationController's accumulators
#include "optimizer/densification.hpp"
#undef private
#include "optimizer/mcmc_densification.hpp"
#include "rasterizer/backward.hpp"
#include "rasterizer/forward.hpp"
#include "rasterizer/projection.hpp"
#include "rasterizer/projection_backward.hpp"
#include "rasterizer/rasterizer.hpp"
#include "rasterizer/sorting.hpp"
#include "training/loss.hpp"
#include "utils/ply_io.hpp"

namespace {

using T = torch::Tensor;

cugs::CameraInfo make_camera(const std::vector<double>& c) {
    TORCH_CHECK(c.size() == 18, "camera must be 18 numbers");
    cugs::CameraInfo cam;
    cam.width = static_cast<int>(c[0]);
    cam.height = static_cast<int>(c[1]);
    cam.intrinsics.fx = static_cast<float>(c[2]);
    cam.intrinsics.fy = static_cast<float>(c[3]);
    cam.intrinsics.cx = static_cast<float>(c[4]);
    cam.intrinsics.cy = static_cast<float>(c[5]);
    for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k) cam.rotation(r, k) = static_cast<float>(c[6 + r * 3 + k]);
    for (int r = 0; r < 3; ++r) cam.translation(r) = static_cast<float>(c[15 + r]);
    return cam;
}

cugs::GaussianModel make_model(const T& pos, const T& sh, const T& opa, const T& rot, const T& scl) {
    cugs::GaussianModel m;
    m.positions = pos;
    m.sh_coeffs = sh;
    m.opacities = opa;
    m.rotations = rot;
    m.scales = scl;
    return m;
}

cugs::RenderSettings make_settings(const std::vector<double>& bg, int deg, double scale_mod) {
    cugs::RenderSettings s;
    s.background[0] = static_cast<float>(bg[0]);
    s.background[1] = static_cast<float>(bg[1]);
    s.background[2] = static_cast<float>(bg[2]);
    s.active_sh_degree = deg;
    s.scale_modifier = static_cast<float>(scale_mod);
    return s;
}

std::vector<T> pack(const cugs::RenderOutput& o) {
    return {o.color, o.final_T, o.n_contrib, o.means_2d, o.depths, o.cov_2d_inv,
            o.radii, o.rgb, o.opacities_act, o.gaussian_indices, o.tile_ranges};
}

cugs::RenderOutput unpack(const std::vector<T>& v) {
    TORCH_CHECK(v.size() == 11, "render output must be 11 tensors");
    return cugs::RenderOutput{v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10]};
}

std::vector<T> ref_project(const T& pos, const T& rot, const T& scl, const T& opa, const T& sh,
                           const std::vector<double>& cam, int deg, double scale_mod) {
    auto o = cugs::project_gaussians(pos, rot, scl, opa, sh, make_camera(cam), deg,
                                     static_cast<float>(scale_mod));
    return {o.means_2d, o.depths, o.cov_2d_inv, o.radii, o.tiles_touched, o.rgb, o.opacities_act};
}

std::vector<T> ref_sort(const T& means_2d, const T& depths, const T& radii, const T& tiles,
                        int w, int h) {
    auto o = cugs::sort_gaussians(means_2d, depths, radii, tiles, w, h);
    auto p = torch::tensor({static_cast<int64_t>(o.total_pairs)}, torch::kInt64);
    return {o.gaussian_keys_sorted, o.gaussian_values_sorted, o.tile_ranges, p};
}

std::vector<T> ref_rasterize_forward(const T& means_2d, const T& cov, const T& rgb, const T& opa,
                                     const T& ranges, const T& idx, int w, int h,
                                     const std::vector<double>& bg) {
    float b[3] = {(float)bg[0], (float)bg[1], (float)bg[2]};
    auto o = cugs::rasterize_forward(means_2d, cov, rgb, opa, ranges, idx, w, h, b);
    return {o.color, o.final_T, o.n_contrib};
}

std::vector<T> ref_rasterize_backward(const T& dL_dcolor, const T& means_2d, const T& cov,
                                      const T& rgb, const T& opa, const T& ranges, const T& idx,
                                      const T& final_T, const T& n_contrib, int w, int h,
                                      const std::vector<double>& bg, int n) {
    float b[3] = {(float)bg[0], (float)bg[1], (float)bg[2]};
    auto o = cugs::rasterize_backward(dL_dcolor, means_2d, cov, rgb, opa, ranges, idx, final_T,
                                      n_contrib, w, h, b, n);
    return {o.dL_drgb, o.dL_dopacity_act, o.dL_dmeans_2d, o.dL_dcov_2d_inv};
}

std::vector<T> ref_project_backward(const T& d_m2d, const T& d_cov, const T& d_rgb, const T& d_opa,
                                    const T& pos, const T& rot, const T& scl, const T& opa,
                                    const T& sh, const T& radii, const std::vector<double>& cam,
                                    int deg, double scale_mod) {
    auto o = cugs::project_backward(d_m2d, d_cov, d_rgb, d_opa, pos, rot, scl, opa, sh, radii,
                                    make_camera(cam), deg, static_cast<float>(scale_mod));
    return {o.dL_dpositions, o.dL_drotations, o.dL_dscales, o.dL_dopacities, o.dL_dsh_coeffs};
}

std::vector<T> ref_render(const T& pos, const T& sh, const T& opa, const T& rot, const T& scl,
                          const std::vector<double>& cam, const std::vector<double>& bg, int deg,
                          double scale_mod) {
    return pack(cugs::render(make_model(pos, sh, opa, rot, scl), make_camera(cam),
                             make_settings(bg, deg, scale_mod)));
}

std::vector<T> ref_render_backward(const T& dL_dcolor, const std::vector<T>& render_out,
                                   const T& pos, const T& sh, const T& opa, const T& rot,
                                   const T& scl, const std::vector<double>& cam,
                                   const std::vector<double>& bg, int deg, double scale_mod) {
    auto o = cugs::render_backward(dL_dcolor, unpack(render_out), make_model(pos, sh, opa, rot, scl),
                                   make_camera(cam), make_settings(bg, deg, scale_mod));
    return {o.dL_dpositions, o.dL_drotations, o.dL_dscales, o.dL_dopacities, o.dL_dsh_coeffs,
            o.dL_dmeans_2d};
}

// loss value + autograd gradient, exactly as the trainer wires it (trainer.cpp:214-217).
std::vector<T> ref_combined_loss_with_grad(const T& rendered, const T& target, double lambda) {
    auto r = rendered.clone().detach().requires_grad_(true);
    auto loss = cugs::combined_loss(r, target, static_cast<float>(lambda));
    loss.backward();
    auto l1 = cugs::l1_loss(rendered, target);
    auto s = cugs::ssim(rendered, target).mean();
    return {loss.detach(), l1, s, r.grad().clone()};
}

struct RefAdam {
    cugs::GaussianModel model;
    std::unique_ptr<cugs::FusedAdam> opt;
    RefAdam(const T& pos, const T& sh, const T& opa, const T& rot, const T& scl)
        : model(make_model(pos, sh, opa, rot, scl)) {
        opt = std::make_unique<cugs::FusedAdam>(model, cugs::AdamConfig{});
    }
    // grads in BackwardOutput order: positions, rotations, scales, opacities, sh
    void step(const std::vector<T>& g, int step_idx) {
        torch::NoGradGuard ng;
        cugs::BackwardOutput b{g[0], g[1], g[2], g[3], g[4], T()};
        opt->update_lr(step_idx);
        opt->zero_grad();
        opt->apply_gradients(b);
        opt->step();
    }
    std::vector<T> params() const {
        return {model.positions, model.sh_coeffs, model.opacities, model.rotations, model.scales};
    }
};

// The per-step sequence of Trainer::train_step (training/trainer.cpp:201-242) written against the
// reference's public API only: render -> combined_loss + autograd -> render_backward -> FusedAdam.
// Compiled unchanged into cugs_ref (reference kernels) and cugs_dropin (libcugs_b200 behind
// wrapper/cugs_b200_dropin.cpp): the same caller code runs on both libraries.
// Returns {loss per step [K], positions, sh_coeffs, opacities, rotations, scales}.
std::vector<T> ref_train_steps(const T& pos, const T& sh, const T& opa, const T& rot, const T& scl,
                               const std::vector<double>& cam, const T& target, const std::vector<double>& bg,
                               int max_degree, double lambda, int first_step, int n_steps) {
    auto model = make_model(pos.clone(), sh.clone(), opa.clone(), rot.clone(), scl.clone());
    cugs::FusedAdam optimizer(model, cugs::AdamConfig{});
    const auto camera = make_camera(cam);
    std::vector<float> losses;
    for (int step = first_step; step < first_step + n_steps; ++step) {
        optimizer.update_lr(step);                                                    // trainer.cpp:180
        const int degree = cugs::active_sh_degree_for_step(step, max_degree);         // :183
        const auto settings = make_settings(bg, degree, 1.0);
        cugs::RenderOutput out;
        {
            torch::NoGradGuard ng;
            out = cugs::render(model, camera, settings);                              // :211
        }
        auto rendered = out.color.clone().detach().requires_grad_(true);             // :214-217
        auto loss = cugs::combined_loss(rendered, target, static_cast<float>(lambda));
        loss.backward();
        auto dL_dcolor = rendered.grad().clone();
        torch::NoGradGuard ng;
        auto grads = cugs::render_backward(dL_dcolor, out, model, camera, settings);  // :228
        optimizer.zero_grad();                                                        // :240-242
        optimizer.apply_gradients(grads);
        optimizer.step();
        losses.push_back(loss.item<float>());
    }
    return {torch::tensor(losses), model.positions.detach(), model.sh_coeffs.detach(), model.opacities.detach(),
            model.rotations.detach(), model.scales.detach()};
}

// Gaussian PLY checkpoint format (utils/ply_io.cpp:98-196, :258-351), CPU tensors
bool ref_write_gaussian_ply(const std::string& path, const T& pos, const T& sh, const T& opa, const T& rot,
                            const T& scl) {
    return cugs::write_gaussian_ply(path, make_model(pos, sh, opa, rot, scl));
}
std::vector<T> ref_read_gaussian_ply(const std::string& path) {
    auto m = cugs::read_gaussian_ply(path);
    return {m.positions, m.sh_coeffs, m.opacities, m.rotations, m.scales};
}

// MCMCController::compute_regularization (autograd) -> {loss, dL/dopacities, dL/dscales}
std::vector<T> ref_mcmc_regularization(const T& pos, const T& sh, const T& opa, const T& rot, const T& scl,
                                       double lambda_opacity, double lambda_scale) {
    cugs::MCMCConfig cfg;
    cfg.lambda_opacity = static_cast<float>(lambda_opacity);
    cfg.lambda_scale = static_cast<float>(lambda_scale);
    cugs::MCMCController ctrl(cfg, 1.0f);
    T d_opa, d_scl;
    const float loss = ctrl.compute_regularization(make_model(pos, sh, opa, rot, scl), d_opa, d_scl);
    return {torch::tensor({loss}), d_opa, d_scl};
}

// MCMCController::inject_noise, in place on `pos`; returns the noise learning rate of the step
double ref_mcmc_inject_noise(T pos, const T& sh, const T& opa, const T& rot, const T& scl, int step) {
    cugs::MCMCController ctrl(cugs::MCMCConfig{}, 1.0f);
    auto model = make_model(pos, sh, opa, rot, scl);
    ctrl.inject_noise(model, step);
    return ctrl.noise_lr(step);
}

double ref_mcmc_noise_lr(int step) { return cugs::MCMCController(cugs::MCMCConfig{}, 1.0f).noise_lr(step); }

// DensificationController::accumulate_gradients applied `times` times -> {grad_accum, grad_count, max_radii}
std::vector<T> ref_accumulate_gradients(const T& dL_dmeans_2d, const T& radii, int times) {
    cugs::DensificationController ctrl(cugs::DensificationConfig{}, 1.0f);
    for (int i = 0; i < times; ++i) ctrl.accumulate_gradients(dL_dmeans_2d, radii);
    return {ctrl.grad_accum_, ctrl.grad_count_, ctrl.max_radii_2d_};
}

// DensificationController::densify (optimizer/densification.cpp:94-329) on given accumulators.
// cfg = {grad_threshold, opacity_threshold, percent_dense, max_screen_size, max_gaussians, opacity_reset_every}
// -> {positions, sh_coeffs, opacities, rotations, scales, stats[cloned, split, pruned, before, after]}
std::vector<T> ref_densify(const T& pos, const T& sh, const T& opa, const T& rot, const T& scl, const T& grad_accum,
                           const T& grad_count, const T& max_radii, double scene_extent, int step,
                           const std::vector<double>& cfg) {
    cugs::DensificationConfig c;
    c.grad_threshold = static_cast<float>(cfg[0]);
    c.opacity_threshold = static_cast<float>(cfg[1]);
    c.percent_dense = static_cast<float>(cfg[2]);
    c.max_screen_size = static_cast<int>(cfg[3]);
    c.max_gaussians = static_cast<int>(cfg[4]);
    c.opacity_reset_every = static_cast<int>(cfg[5]);
    c.min_vram_headroom_mb = 0.0f;
    cugs::DensificationController ctrl(c, static_cast<float>(scene_extent));
    ctrl.grad_accum_ = grad_accum.clone();
    ctrl.grad_count_ = grad_count.clone();
    ctrl.max_radii_2d_ = max_radii.clone();
    auto model = make_model(pos.clone(), sh.clone(), opa.clone(), rot.clone(), scl.clone());
    const auto st = ctrl.densify(model, step);
    return {model.positions, model.sh_coeffs, model.opacities, model.rotations, model.scales,
            torch::tensor({st.num_cloned, st.num_split, st.num_pruned, st.num_before, st.num_after})};
}

// MCMCController::relocate (optimizer/mcmc_densification.cpp:56-138)
// -> {positions, sh_coeffs, opacities, rotations, scales, stats[relocated, dead, total]}
std::vector<T> ref_mcmc_relocate(const T& pos, const T& sh, const T& opa, const T& rot, const T& scl,
                                 double scene_extent, double dead_threshold, double cap) {
    cugs::MCMCConfig c;
    c.dead_opacity_threshold = static_cast<float>(dead_threshold);
    c.relocate_cap = static_cast<float>(cap);
    c.min_vram_headroom_mb = 0.0f;
    cugs::MCMCController ctrl(c, static_cast<float>(scene_extent));
    auto model = make_model(pos.clone(), sh.clone(), opa.clone(), rot.clone(), scl.clone());
    const auto st = ctrl.relocate(model, 1000);
    return {model.positions, model.sh_coeffs, model.opacities, model.rotations, model.scales,
            torch::tensor({st.num_relocated, st.num_dead, st.num_total})};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "unmodified reference hot path (test oracle + reference bench arm)";
    m.def("project_gaussians", &ref_project);
    m.def("sort_gaussians", &ref_sort);
    m.def("rasterize_forward", &ref_rasterize_forward);
    m.def("rasterize_backward", &ref_rasterize_backward);
    m.def("project_backward", &ref_project_backward);
    m.def("render", &ref_render);
    m.def("render_backward", &ref_render_backward);
    m.def("combined_loss_with_grad", &ref_combined_loss_with_grad,
          py::call_guard<py::gil_scoped_release>());  // autograd must not run under the GIL
    m.def("write_gaussian_ply", &ref_write_gaussian_ply);
    m.def("read_gaussian_ply", &ref_read_gaussian_ply);
    m.def("train_steps", &ref_train_steps, py::call_guard<py::gil_scoped_release>());
    m.def("mcmc_regularization", &ref_mcmc_regularization, py::call_guard<py::gil_scoped_release>());
    m.def("mcmc_inject_noise", &ref_mcmc_inject_noise);
    m.def("mcmc_noise_lr", &ref_mcmc_noise_lr);
    m.def("accumulate_gradients", &ref_accumulate_gradients);
    m.def("densify", &ref_densify);
    m.def("mcmc_relocate", &ref_mcmc_relocate);
    m.def("evaluate_sh_cuda", &cugs::evaluate_sh_cuda);
    m.def("evaluate_sh_cpu", &cugs::evaluate_sh_cpu);
    m.def("evaluate_sh_backward_cuda", &cugs::evaluate_sh_backward_cuda);
    m.def("l1_loss", &cugs::l1_loss);
    m.def("ssim", &cugs::ssim, py::arg("rendered"), py::arg("target"), py::arg("window_size") = 11);
    m.def("combined_loss", &cugs::combined_loss, py::arg("rendered"), py::arg("target"),
          py::arg("lambda_") = 0.2f);
    py::class_<RefAdam>(m, "FusedAdam")
        .def(py::init<const T&, const T&, const T&, const T&, const T&>())
        .def("step", &RefAdam::step)
        .def("params", &RefAdam::params);
}
