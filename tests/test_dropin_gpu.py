"""Drop-in check on the GPU: the reference's own C++ host API (GaussianModel / CameraInfo /
RenderSettings / FusedAdam, declared by the reference's unmodified headers) DEFINED by
wrapper/cugs_b200_dropin.cpp on top of libcugs_b200.so, driven through the same pybind harness
(oracle/ref_harness.cpp) as the compiled reference. `cugs_dropin` and `cugs_ref` expose identical
functions, so every check is "same call, two libraries"."""
import sys
from pathlib import Path

import numpy as np
import pytest

import cuda_gaussian_splatting_b200 as cugs
from conftest import ROOT, to_torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("no CUDA device")
    return t


@pytest.fixture(scope="module")
def dropin(torch):
    sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
    try:
        import cugs_dropin
        return cugs_dropin
    except Exception as e:  # pragma: no cover
        pytest.skip(f"oracle/_ref/cugs_dropin*.so not built (make -f oracle/Makefile.ref dropin): {e}")


def np_(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("n,w,h,seed,adv", [(5000, 320, 240, 11, False), (20000, 640, 360, 13, True),
                                            (100_000, 1280, 720, 1235, False)])
def test_render_and_backward_same_call_two_libraries(ref, dropin, torch, n, w, h, seed, adv):
    scene = cugs.synth(n, w, h, seed=seed, adversarial=adv)
    m = to_torch(scene)
    cam, bg = scene.camera.as_ref_list(), [0.1, 0.2, 0.3]
    args = (m.positions, m.sh_coeffs, m.opacities, m.rotations, m.scales, cam, bg, 3, 1.0)
    r, d = ref.render(*args), dropin.render(*args)
    for k, name in enumerate(["color", "final_T", "n_contrib", "means_2d", "depths", "cov_2d_inv", "radii", "rgb",
                              "opacities_act", "gaussian_indices", "tile_ranges"]):
        assert r[k].shape == d[k].shape and r[k].dtype == d[k].dtype, name
    for k in (2, 6, 9, 10):  # n_contrib, radii, sort order, tile ranges: bit-exact
        assert torch.equal(r[k], d[k])
    for k in (3, 4, 5, 8):   # means_2d, depths, cov_2d_inv, opacities_act: bit-exact floats
        assert torch.equal(r[k].view(torch.int32), d[k].view(torch.int32))
    assert float((r[0] - d[0]).abs().max()) <= 1e-4 and float((r[1] - d[1]).abs().max()) <= 1e-6
    g = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, size=(h, w, 3)).astype(np.float32)).cuda()
    rb = ref.render_backward(g, r, *args)
    db = dropin.render_backward(g, d, *args)        # cached frame: fused path
    db2 = dropin.render_backward(g, r, *args)       # foreign RenderOutput: stage-function path
    for a, b_, c in zip(rb, db, db2):
        na = float(a.double().norm())
        assert a.shape == b_.shape == c.shape
        assert float((a.double() - b_.double()).norm()) <= 1e-3 * na + 1e-12
        assert float((a.double() - c.double()).norm()) <= 1e-3 * na + 1e-12


def test_stage_functions_same_call_two_libraries(ref, dropin, torch):
    scene = cugs.synth(5000, 320, 240, seed=11)
    m = to_torch(scene)
    cam = scene.camera.as_ref_list()
    pr = ref.project_gaussians(m.positions, m.rotations, m.scales, m.opacities, m.sh_coeffs, cam, 3, 1.0)
    pd = dropin.project_gaussians(m.positions, m.rotations, m.scales, m.opacities, m.sh_coeffs, cam, 3, 1.0)
    assert torch.equal(pr[3], pd[3]) and torch.equal(pr[4], pd[4])  # radii, tiles_touched
    sr = ref.sort_gaussians(pr[0], pr[1], pr[3], pr[4], 320, 240)
    sd = dropin.sort_gaussians(pr[0], pr[1], pr[3], pr[4], 320, 240)
    assert all(torch.equal(a, b) for a, b in zip(sr, sd))           # keys, values, ranges, P
    fr = ref.rasterize_forward(pr[0], pr[2], pr[5], pr[6], sr[2], sr[1], 320, 240, [0.0, 0.0, 0.0])
    fd = dropin.rasterize_forward(pr[0], pr[2], pr[5], pr[6], sr[2], sr[1], 320, 240, [0.0, 0.0, 0.0])
    assert torch.equal(fr[2], fd[2]) and torch.equal(fr[0], fd[0]) and torch.equal(fr[1], fd[1])  # bit-identical blend


def test_loss_autograd_and_adam_same_call_two_libraries(ref, dropin, torch):
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.uniform(size=(90, 130, 3)).astype(np.float32)).cuda()
    y = torch.from_numpy(rng.uniform(size=(90, 130, 3)).astype(np.float32)).cuda()
    lr_, l1r, sr, gr = ref.combined_loss_with_grad(x, y, 0.2)   # combined_loss + loss.backward() + l1 + ssim().mean()
    ld, l1d, sd, gd = dropin.combined_loss_with_grad(x, y, 0.2)
    assert abs(float(lr_) - float(ld)) <= 1e-5 and abs(float(l1r) - float(l1d)) <= 1e-6 and abs(float(sr) - float(sd)) <= 1e-5
    assert float((gr - gd).abs().max()) <= 1e-3 * float(gr.abs().max())
    assert float((ref.ssim(x, y) - dropin.ssim(x, y)).abs().max()) <= 1e-4   # [H,W] map (metrics.cpp:41-46)
    for window in (7, 5, 15):                                                  # any odd window (loss.hpp:33-44)
        assert float((ref.ssim(x, y, window) - dropin.ssim(x, y, window)).abs().max()) <= 1e-4
    with pytest.raises(RuntimeError):
        dropin.ssim(x, y, 8)
    with pytest.raises(RuntimeError):  # c10::Error, as tests/test_loss.cpp:143-170 expects
        dropin.combined_loss(x[:, :, :2], y[:, :, :2])
    scene = cugs.synth(1003, 64, 48, seed=21)
    a, b = to_torch(scene), to_torch(scene)
    oa = ref.FusedAdam(a.positions, a.sh_coeffs, a.opacities, a.rotations, a.scales)
    ob = dropin.FusedAdam(b.positions, b.sh_coeffs, b.opacities, b.rotations, b.scales)
    for step in range(5):
        g = [torch.from_numpy(rng.normal(size=tuple(t.shape)).astype(np.float32)).cuda()
             for t in (a.positions, a.rotations, a.scales, a.opacities, a.sh_coeffs)]
        oa.step(g, step)
        ob.step(g, step)
    for p, q in zip(oa.params(), ob.params()):
        assert torch.equal(p.view(torch.int32), q.view(torch.int32)), "FusedAdam must be bit-identical"


def test_trainer_sequence_same_caller_code_two_libraries(ref, dropin, torch):
    """The per-step sequence of Trainer::train_step (trainer.cpp:201-242), written once in the harness
    against the reference's public C++ API, runs 25 optimisation steps on the reference kernels and on
    libcugs_b200: same losses, and both fit the target (tests/test_training.cpp:159-261)."""
    rng = np.random.default_rng(42)
    f = np.float32
    cam = cugs.CameraInfo(96, 64, 120.0, 120.0, 48.0, 32.0)
    n = 64
    pos = rng.normal(size=(n, 3)) * 0.6
    pos[:, 2] = np.abs(pos[:, 2]) + 3.0
    rot = rng.normal(size=(n, 4)).astype(f)
    gt = cugs.Scene(pos.astype(f), (rng.normal(size=(n, 3, 16)) * 0.4).astype(f), np.full((n, 1), 1.0, f), rot,
                    np.full((n, 3), -1.6, f), cam)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    target = cugs.render(to_torch(gt), cam, cugs.RenderSettings((0, 0, 0), 3, 1.0)).color.clone()
    start = (t(gt.positions), t(np.zeros((n, 3, 16), f)), t(gt.opacities), t(gt.rotations), t(gt.scales))
    args = (*start, cam.as_ref_list(), target, [0.0, 0.0, 0.0], 3, 0.2, 3000, 25)   # step 3000: SH degree 3
    r = ref.train_steps(*args)
    d = dropin.train_steps(*args)
    lr_, ld = r[0].numpy(), d[0].numpy()
    assert ld[-1] < 0.9 * ld[0] and lr_[-1] < 0.9 * lr_[0], "both must fit the target"
    # Adam with eps = 1e-15 moves parameters by ~lr * sign(g) in its first steps, so tiny gradient
    # differences can flip individual updates; the trajectories must still agree closely
    assert np.abs(ld - lr_).max() <= 2e-3 * lr_[0], (lr_, ld)
    assert abs(ld[0] - lr_[0]) <= 1e-5
    for a, b in zip(r[1:], d[1:]):
        assert float((a - b).abs().mean()) <= 2e-3


# ------------------------------------------------------------------------------------------------
# the schedule-driven callers: DensificationController / MCMCController defined by the wrapper
# ------------------------------------------------------------------------------------------------
def density_inputs(torch, n=12000, seed=3):
    scene = cugs.synth(n, 320, 240, seed=seed)
    m = to_torch(scene)
    g = torch.Generator(device="cuda").manual_seed(seed)
    count = torch.randint(0, 6, (n,), device="cuda", generator=g).float()
    accum = torch.rand((n,), device="cuda", generator=g) * 8e-4 * count.clamp_min(1) * (count > 0)
    radii = torch.randint(0, 41, (n,), device="cuda", generator=g).float()
    m.opacities[torch.rand((n,), device="cuda", generator=g) < 0.05] = -7.0
    extent = float(torch.exp(m.scales).max(dim=1).values.median()) / 0.01
    return m, accum, count, radii, extent


@pytest.mark.parametrize("step,max_gaussians", [(600, 0), (3100, 0), (700, 12100)])
def test_densify_same_call_two_libraries(ref, dropin, torch, step, max_gaussians):
    m, accum, count, radii, extent = density_inputs(torch)
    args = (m.positions, m.sh_coeffs, m.opacities, m.rotations, m.scales, accum, count, radii, extent, step,
            [0.0002, 0.005, 0.01, 20, max_gaussians, 3000])
    r, d = ref.densify(*args), dropin.densify(*args)
    assert r[5].tolist() == d[5].tolist()                     # cloned, split, pruned, before, after
    cloned, split, pruned, before, after = r[5].tolist()
    assert cloned > 0 and pruned > 0 and (split > 0 or max_gaussians > 0)
    head = after - 2 * split
    for k in range(5):
        assert r[k].shape == d[k].shape
        assert torch.equal(r[k][:head].view(torch.int32), d[k][:head].view(torch.int32))   # kept + cloned rows
    for k in (1, 2, 3, 4):                                    # children: sh, opacity, rotation, scale are copies
        assert torch.equal(r[k][head:].view(torch.int32), d[k][head:].view(torch.int32))
    if split:
        # child position = parent + N(0,1) * exp(new scale): normalised residuals of both libraries are N(0,1)
        s_new = d[4][head:]
        for lib in (r, d):
            parents = lib[0][head:] - 0  # children
            z = (parents[:split] - parents[split:]) / torch.exp(s_new[:split]) / 2 ** 0.5   # difference of two draws
            assert abs(float(z.mean())) < 5 / (3 * split) ** 0.5 and abs(float(z.std()) - 1) < 0.06


def test_mcmc_controller_same_call_two_libraries(ref, dropin, torch):
    scene = cugs.synth(40000, 320, 240, seed=9)
    m = to_torch(scene)
    m.opacities[::9] = -8.0
    args = (m.positions, m.sh_coeffs, m.opacities, m.rotations, m.scales)
    r, d = ref.mcmc_relocate(*args, 4.0, 0.005, 0.05), dropin.mcmc_relocate(*args, 4.0, 0.005, 0.05)
    assert r[5].tolist() == d[5].tolist() and r[5].tolist()[0] == 2000
    moved_r, moved_d = (r[2] != m.opacities).squeeze(1), (d[2] != m.opacities).squeeze(1)
    assert torch.equal(moved_r, moved_d)                       # the same (first `cap`) dead Gaussians move
    for k in range(5):
        assert torch.equal(r[k][~moved_r].view(torch.int32), d[k][~moved_r].view(torch.int32))
    assert torch.equal(r[2].view(torch.int32), d[2].view(torch.int32))   # logit(0.01) everywhere it moved
    # regulariser: loss and both gradients (autograd there, closed form here)
    rr, dd = ref.mcmc_regularization(*args, 0.01, 0.02), dropin.mcmc_regularization(*args, 0.01, 0.02)
    assert abs(float(rr[0]) - float(dd[0])) <= 1e-6 * max(1.0, abs(float(rr[0])))
    for a, b_ in zip(rr[1:], dd[1:]):
        assert a.shape == b_.shape and float((a - b_).abs().max()) <= 1e-6 * float(a.abs().max()) + 1e-12
    # noise: same learning rate; displacement = lr * exp(scale) * gate * N(0,1) in both
    for step in (0, 100, 15000, 30000):
        assert ref.mcmc_noise_lr(step) == dropin.mcmc_noise_lr(step)
    alive = to_torch(scene)
    alive.opacities.fill_(-3.0)                                # gate ~ 1 everywhere
    p_r, p_d = alive.positions.clone(), alive.positions.clone()
    ref.mcmc_inject_noise(p_r, alive.sh_coeffs, alive.opacities, alive.rotations, alive.scales, 29000)
    dropin.mcmc_inject_noise(p_d, alive.sh_coeffs, alive.opacities, alive.rotations, alive.scales, 29000)
    lr = ref.mcmc_noise_lr(29000)
    for p in (p_r, p_d):
        z = (p - alive.positions) / (lr * torch.exp(alive.scales))
        assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1) < 0.03
    # accumulate_gradients through the controller of either library
    g2 = torch.randn((40000, 2), device="cuda")
    rad = torch.randint(0, 5, (40000,), device="cuda", dtype=torch.int32)
    ra, da = ref.accumulate_gradients(g2, rad, 3), dropin.accumulate_gradients(g2, rad, 3)
    assert torch.allclose(ra[0], da[0], rtol=1e-6, atol=0)     # sum of norms (torch.norm vs sqrt(x*x + y*y))
    assert torch.equal(ra[1], da[1]) and torch.equal(ra[2], da[2])   # visibility count, max radius
