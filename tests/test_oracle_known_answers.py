"""Pins the CPU oracle (oracle/cugs_oracle.c) against the known-answer tests of the reference's
own test-suite (ported one by one, citations relative to /root/reference/tests) and against the
golden vectors produced by the unmodified reference (tests/golden/). CPU only."""
from pathlib import Path

import numpy as np
import pytest

import cuda_gaussian_splatting_b200 as cugs
from cuda_gaussian_splatting_b200 import CameraInfo, Scene

GOLDEN = Path(__file__).resolve().parent / "golden"
C0 = 0.28209479177387814


def cam_640():  # test_projection.cpp:24-35
    return CameraInfo(640, 480, 500.0, 500.0, 320.0, 240.0)


def cam_160(w=160, h=120):  # test_rasterizer.cpp:22-33
    return CameraInfo(w, h, 200.0, 200.0, w / 2.0, h / 2.0)


def single(x, y, z, cam, log_s=(-2.0, -2.0, -2.0), opa=0.0, sh_dc=1.0):  # test_projection.cpp:38-57
    f = np.float32
    return Scene(np.array([[x, y, z]], f), np.full((1, 3, 1), sh_dc, f), np.array([[opa]], f),
                 np.array([[1, 0, 0, 0]], f), np.array([list(log_s)], f), cam)


def project(oracle, s, deg=0, scale_mod=1.0):
    return oracle.preprocess_fwd((s.positions, s.rotations, s.scales, s.opacities, s.sh_coeffs), s.camera, deg, scale_mod)


# ---- projection (test_projection.cpp) ---------------------------------------------------------
def test_single_gaussian_in_front(oracle):  # :64-103
    o = project(oracle, single(0, 0, 5, cam_640()))
    assert o["radii"][0] > 0
    assert abs(o["means_2d"][0, 0] - 320.0) <= 1.0 and abs(o["means_2d"][0, 1] - 240.0) <= 1.0
    assert abs(o["depths"][0] - 5.0) <= 0.01
    assert abs(o["opacities_act"][0] - 0.5) <= 0.01
    assert o["tiles_touched"][0] > 0
    assert (o["rgb"] >= 0).all()
    assert abs(o["rgb"][0, 0] - (C0 * 1.0 + 0.5)) < 1e-6


def test_behind_camera_culled(oracle):  # :109-125
    o = project(oracle, single(0, 0, -5, cam_640()))
    assert o["radii"][0] == 0 and o["tiles_touched"][0] == 0


def test_off_center_projection(oracle):  # :131-149
    o = project(oracle, single(1, 0, 5, cam_640()))
    assert abs(o["means_2d"][0, 0] - 420.0) <= 1.0 and abs(o["means_2d"][0, 1] - 240.0) <= 1.0


def test_random_no_nan(oracle):  # :155-185
    s = cugs.synth(1000, 640, 480, seed=7, num_coeffs=16)
    o = project(oracle, s, deg=3)
    for k in ("means_2d", "depths", "cov_2d_inv", "rgb", "opacities_act"):
        assert np.isfinite(o[k]).all(), k


def test_anisotropy_and_scale_modifier_grow_radius(oracle):  # :191-217, :245-266
    r_iso = project(oracle, single(0, 0, 5, cam_640()))["radii"][0]
    r_aniso = project(oracle, single(0, 0, 5, cam_640(), log_s=(0.0, -2.0, -2.0)))["radii"][0]
    r_mod = project(oracle, single(0, 0, 5, cam_640()), scale_mod=2.0)["radii"][0]
    assert r_aniso > r_iso and r_mod > r_iso


def test_tile_count_quirk_and_filler_keys(oracle):
    """a Gaussian off-screen on BOTH axes keeps a positive tile count (the
    product of two negative extents) but emits no key; its slots stay key 0 / value 0."""
    cam = CameraInfo(1920, 1080, 1440.0, 1440.0, 960.0, 540.0)
    s = single((5000 - 960) * 5 / 1440.0, (3000 - 540) * 5 / 1440.0, 5.0, cam)
    o = project(oracle, s)
    assert o["radii"][0] > 0 and o["tiles_touched"][0] > 0
    off, P = oracle.scan(o["tiles_touched"])
    keys, vals = oracle.fill_keys(o["means_2d"], o["depths"], o["radii"], off, 1920, 1080, P)
    assert P == o["tiles_touched"][0] and (keys == 0).all() and (vals == 0).all()


# ---- SH (test_sh.cpp) -------------------------------------------------------------------------
def test_sh_degree0_constant(oracle):  # :16-35
    dc = (0.7 - 0.5) / C0
    sh = np.full((1, 3, 1), dc, np.float32)
    out = oracle.sh_forward(0, sh, np.array([[0, 0, 1]], np.float32))
    assert np.allclose(out, 0.7, atol=1e-5)


def test_sh_degree1_antisymmetry(oracle):  # :57-82
    sh = np.zeros((1, 3, 4), np.float32)
    sh[0, 0, 1] = 1.0
    yp = oracle.sh_forward(1, sh, np.array([[0, 1, 0]], np.float32))[0, 0] - 0.5
    yn = oracle.sh_forward(1, sh, np.array([[0, -1, 0]], np.float32))[0, 0] - 0.5
    x = oracle.sh_forward(1, sh, np.array([[1, 0, 0]], np.float32))[0, 0] - 0.5
    assert abs(yp + yn) < 1e-5 and abs(x) < 1e-5 and abs(yp + 0.4886025119029199) < 1e-6


def test_sh_higher_degrees_with_zero_coeffs(oracle):  # :110-125
    sh = np.zeros((1, 3, 16), np.float32)
    sh[:, :, 0] = 1.2
    d = np.array([[0.577, 0.577, 0.577]], np.float32)
    assert np.allclose(oracle.sh_forward(0, sh, d), oracle.sh_forward(3, sh, d), atol=1e-5)


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_matches_reference_cpu_golden(oracle, deg):
    """Golden vectors from the UNMODIFIED reference evaluate_sh_cpu (tests/golden/make_golden_cpu.py);
    tolerance = the reference's own CUDA-vs-CPU bar (test_sh.cpp:161-203: allclose 1e-4)."""
    g = np.load(GOLDEN / "sh_cpu_golden.npz")
    out = oracle.sh_forward(deg, g[f"sh_{deg}"], g[f"dir_{deg}"])
    assert np.allclose(out, g[f"rgb_{deg}"], rtol=1e-4, atol=1e-5)
    assert np.abs(out - g[f"rgb_{deg}"]).max() < 2e-6


def test_sh_backward_is_linear_and_gated(oracle):
    rng = np.random.default_rng(0)
    sh = rng.normal(size=(32, 3, 16)).astype(np.float32)
    d = rng.normal(size=(32, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    g = rng.normal(size=(32, 3)).astype(np.float32)
    out = oracle.sh_backward(2, sh, d, g)
    assert (out[:, :, 9:] == 0).all()  # explicit zeros for inactive coefficients (sh_backward.cu:103-110)
    raw = oracle.sh_forward(2, sh, d)
    gate = raw > 0
    assert np.allclose(out[:, :, 0], g * gate * C0, atol=1e-6)


# ---- full forward (test_rasterizer.cpp) ---------------------------------------------------------
def test_empty_scene_is_background(oracle):  # :72-106
    cam = cam_160()
    z = np.zeros
    s = Scene(z((0, 3), np.float32), z((0, 3, 1), np.float32), z((0, 1), np.float32), z((0, 4), np.float32),
              z((0, 3), np.float32), cam)
    o = oracle.render_forward(s, deg=0, bg=(0.3, 0.5, 0.7))
    assert np.allclose(o["color"][60, 80], [0.3, 0.5, 0.7], atol=0.01)
    assert (o["final_T"] == 1).all() and (o["n_contrib"] == 0).all()


def test_single_gaussian_center_and_corner(oracle):  # :112-150, :202-230, :277-302
    cam = cam_160()
    s = single(0, 0, 5, cam, opa=5.0, sh_dc=1.0)
    o = oracle.render_forward(s, deg=0, bg=(1.0, 0.0, 1.0))
    c = o["color"]
    assert c[60, 80, 1] > 0.1 and c[60, 80, 1] > c[0, 0, 1]
    assert np.allclose(c[0, 0], [1.0, 0.0, 1.0], atol=0.05)
    assert o["final_T"][60, 80] < 0.5 and o["n_contrib"][60, 80] >= 1


def test_depth_ordering_front_dominates(oracle):  # :156-196
    cam = cam_160()
    f = np.float32
    s = Scene(np.array([[0, 0, 3], [0, 0, 6]], f), np.array([[[2.0]] * 3, [[-2.0]] * 3], f).reshape(2, 3, 1),
              np.full((2, 1), 5.0, f), np.array([[1, 0, 0, 0]] * 2, f), np.full((2, 3), -1.5, f), cam)
    o = oracle.render_forward(s, deg=0)
    assert o["color"][60, 80, 0] > 0.5
    assert list(o["gaussian_indices"][:1]) == [0] or o["keys_sorted"][0] <= o["keys_sorted"][-1]


def test_random_scene_invariants(oracle):  # :236-271 + bit-level structure
    s = cugs.synth(500, 320, 240, seed=3)
    o = oracle.render_forward(s, deg=3)
    assert np.isfinite(o["color"]).all()
    assert ((o["final_T"] >= 0) & (o["final_T"] <= 1)).all()
    ks = o["keys_sorted"]
    assert (ks[1:] >= ks[:-1]).all()
    assert o["P"] == int(o["tiles_touched"].sum())
    r = o["tile_ranges"]
    assert int((r[:, 1] - r[:, 0]).sum()) == o["P"]


def test_culled_gaussian_has_zero_gradients(oracle):  # test_backward.cpp:181-201
    cam = cam_160(64, 48)
    s = single(0, 0, -5, cam, opa=5.0)
    fwd = oracle.render_forward(s, deg=0)
    g = np.random.default_rng(1).uniform(size=(48, 64, 3)).astype(np.float32)
    b = oracle.render_backward(s, fwd, g, deg=0)
    for k in ("dL_dpositions", "dL_drotations", "dL_dscales", "dL_dopacities", "dL_dsh_coeffs"):
        assert np.abs(b[k]).sum() == 0.0


def make_test_gaussians(n, cam, seed=42):  # test_backward.cpp:73-92 (numpy rng instead of torch's)
    rng = np.random.default_rng(seed)
    f = np.float32
    pos = rng.normal(size=(n, 3)) * 0.3
    pos[:, 2] = np.abs(pos[:, 2]) + 3.5
    rot = rng.normal(size=(n, 4))
    rot /= np.linalg.norm(rot, axis=1, keepdims=True)
    scl = -1.5 + rng.normal(size=(n, 3)) * 0.2
    return Scene(pos.astype(f), (rng.normal(size=(n, 3, 1)) * 0.5).astype(f), np.full((n, 1), 2.0, f), rot.astype(f),
                 scl.astype(f), cam)


@pytest.mark.parametrize("param,eps,rel,absl", [  # test_backward.cpp:355,:372,:389,:406,:423
    ("positions", 2e-3, 0.15, 1e-3), ("scales", 1e-3, 0.10, 1e-4), ("rotations", 1e-3, 0.10, 1e-4),
    ("opacities", 1e-3, 0.05, 1e-4), ("sh_coeffs", 1e-3, 0.05, 1e-4)])
def test_finite_differences(oracle, param, eps, rel, absl):
    """Central finite differences through render + combined_loss, the reference's own gradient
    check (test_backward.cpp:266-336: mixed tolerance, >= 80 % of elements must pass). The fixture
    uses numpy's RNG (torch's CUDA stream of manual_seed(42) cannot be reproduced on the CPU); on it
    the alpha < 1/255 cut-off, which finite differences see and the analytic gradient does not, biases
    the scale derivatives by 4-7 %, so the scale tolerance is 10 % instead of the reference's 5 %."""
    cam = CameraInfo(64, 48, 100.0, 100.0, 32.0, 24.0)
    s = make_test_gaussians(3, cam)
    target = np.random.default_rng(5).uniform(size=(48, 64, 3)).astype(np.float32)

    def loss_of(scene):
        return float(oracle.loss(oracle.render_forward(scene, deg=0)["color"], target, 0.2, want_grad=False)[0][0])

    fwd = oracle.render_forward(s, deg=0)
    _, dL_dcolor = oracle.loss(fwd["color"], target, 0.2)
    grads = oracle.render_backward(s, fwd, dL_dcolor, deg=0)
    analytic = grads["dL_d" + param].reshape(-1)
    arr = getattr(s, param)
    flat = arr.reshape(-1)
    passed = 0
    for i in range(flat.size):
        orig = flat[i]
        flat[i] = orig + eps
        lp = loss_of(s)
        flat[i] = orig - eps
        lm = loss_of(s)
        flat[i] = orig
        num = (lp - lm) / (2 * eps)
        err = abs(analytic[i] - num)
        if err / max(abs(analytic[i]), abs(num), 1e-6) <= rel or err <= absl:
            passed += 1
    assert passed / flat.size >= 0.80, f"{param}: {passed}/{flat.size}"


# ---- loss (test_loss.cpp) ---------------------------------------------------------------------
def test_loss_known_answers(oracle):
    rng = np.random.default_rng(2)
    x = rng.uniform(size=(32, 40, 3)).astype(np.float32)
    y = rng.uniform(size=(32, 40, 3)).astype(np.float32)
    sc, _ = oracle.loss(x, x, 0.2)
    assert abs(sc[1]) < 1e-7 and abs(sc[2] - 1.0) < 1e-4 and abs(sc[0]) < 1e-4  # :39-41, :64-72, :113-121
    a, b = np.full((16, 16, 3), 0.8, np.float32), np.full((16, 16, 3), 0.3, np.float32)
    assert abs(oracle.loss(a, b, 0.0)[0][1] - 0.5) < 1e-5  # :42-48
    blk, wht = np.zeros((32, 32, 3), np.float32), np.ones((32, 32, 3), np.float32)
    assert oracle.loss(blk, wht, 1.0)[0][2] < 0.1  # :74-83
    s1, s2 = oracle.loss(x, y, 1.0)[0][2], oracle.loss(y, x, 1.0)[0][2]
    assert abs(s1 - s2) < 1e-5 and -1.0 <= s1 <= 1.0  # :85-107
    l_small = oracle.loss(x, x + np.float32(0.01), 0.2)[0][0]
    l_big = oracle.loss(x, x + np.float32(0.2), 0.2)[0][0]
    assert l_big > l_small > 0  # :123-137


def test_loss_gradient_matches_finite_differences(oracle):
    rng = np.random.default_rng(4)
    x = rng.uniform(0.1, 0.9, size=(20, 24, 3)).astype(np.float32)
    y = rng.uniform(0.1, 0.9, size=(20, 24, 3)).astype(np.float32)
    _, g = oracle.loss(x, y, 0.2)
    for (r, c, ch) in [(0, 0, 0), (10, 12, 1), (19, 23, 2), (5, 0, 1), (3, 7, 2)]:
        xp, xm = x.copy(), x.copy()
        e = 1e-3
        xp[r, c, ch] += e
        xm[r, c, ch] -= e
        num = (float(oracle.loss(xp, y, 0.2, False)[0][0]) - float(oracle.loss(xm, y, 0.2, False)[0][0])) / (2 * e)
        assert abs(num - g[r, c, ch]) <= 0.03 * abs(num) + 2e-6, (r, c, ch, num, g[r, c, ch])


def test_loss_matches_torch_restatement(oracle):
    """The same op graph as training/loss.cpp:83-135 written with torch CPU ops + autograd."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(9)
    x = torch.from_numpy(rng.uniform(size=(37, 45, 3)).astype(np.float32)).requires_grad_(True)
    y = torch.from_numpy(rng.uniform(size=(37, 45, 3)).astype(np.float32))
    k1 = torch.tensor([np.exp(-float(i - 5) ** 2 / (2 * 1.5 * 1.5)) for i in range(11)], dtype=torch.float32)
    k1 = k1 / k1.sum()
    k2 = k1[:, None] * k1[None, :]
    k2 = (k2 / k2.sum())[None, None].expand(3, 1, 11, 11).contiguous()
    X, Y = x.permute(2, 0, 1)[None], y.permute(2, 0, 1)[None]
    conv = lambda t: F.conv2d(t, k2, padding=5, groups=3)
    mx, my = conv(X), conv(Y)
    sxx, syy, sxy = conv(X * X) - mx * mx, conv(Y * Y) - my * my, conv(X * Y) - mx * my
    smap = ((2 * mx * my + 1e-4) * (2 * sxy + 9e-4)) / ((mx * mx + my * my + 1e-4) * (sxx + syy + 9e-4))
    l1 = (x - y).abs().mean()
    loss = 0.8 * l1 + 0.2 * (1 - smap.mean())
    loss.backward()
    sc, g = oracle.loss(x.detach().numpy(), y.numpy(), 0.2)
    assert abs(sc[0] - loss.item()) < 1e-5 and abs(sc[1] - l1.item()) < 1e-6 and abs(sc[2] - smap.mean().item()) < 1e-5
    assert np.abs(g - x.grad.numpy()).max() < 1e-7 + 1e-3 * np.abs(x.grad.numpy()).max()


# ---- Adam (test_fused_adam.cpp) and stats (test_densification.cpp) ------------------------------
def test_adam_matches_torch_optim(oracle):  # :95-145
    import torch
    rng = np.random.default_rng(11)
    p0 = rng.normal(size=1000).astype(np.float32)
    p = p0.copy()
    m, v = np.zeros_like(p), np.zeros_like(p)
    tp = torch.from_numpy(p0.copy()).requires_grad_(True)
    opt = torch.optim.Adam([tp], lr=1e-3, betas=(0.9, 0.999), eps=1e-15)
    for step in range(1, 11):
        g = rng.normal(size=1000).astype(np.float32)
        bc1 = 1.0 / (1.0 - float(np.float32(0.9)) ** step)
        bc2 = 1.0 / (1.0 - float(np.float32(0.999)) ** step)
        oracle.adam(p, g, m, v, 1e-3, 0.9, 0.999, 1e-15, bc1, bc2)
        tp.grad = torch.from_numpy(g.copy())
        opt.step()
        if step == 1:
            assert np.allclose(p, tp.detach().numpy(), rtol=1e-5, atol=1e-6)
    assert np.allclose(p, tp.detach().numpy(), rtol=1e-4, atol=1e-5)


def test_adam_zero_grad_keeps_params(oracle):  # :202-225
    p = np.random.default_rng(1).normal(size=64).astype(np.float32)
    q = p.copy()
    m, v = np.zeros_like(p), np.zeros_like(p)
    oracle.adam(p, np.zeros_like(p), m, v, 0.05, 0.9, 0.999, 1e-15, 10.0, 1000.0)
    assert (p == q).all()


def test_accumulate_stats_visibility_rule(oracle):  # test_densification.cpp:134-161
    g = np.array([[3, 4], [1, 0], [0, 2]], np.float32)
    r = np.array([5, 0, 2], np.int32)
    acc, cnt, mx = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
    oracle.accumulate_stats(g, r, acc, cnt, mx)
    oracle.accumulate_stats(g, r, acc, cnt, mx)
    assert np.allclose(acc, [10, 0, 4]) and np.allclose(cnt, [2, 0, 2]) and np.allclose(mx, [5, 0, 2])
