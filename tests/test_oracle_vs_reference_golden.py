"""Pins the CPU oracle against outputs of the UNMODIFIED reference CUDA kernels, recorded on a
B200 by tests/golden/make_golden_gpu.py (tests/golden/ref_gpu_golden.npz). CPU only: this is the
check that the checker itself restates the reference -- integer stages bit-exact, floats within the
libm/libdevice last-ulp differences the oracle header names."""
from pathlib import Path

import numpy as np
import pytest

import cuda_gaussian_splatting_b200 as cugs
from cuda_gaussian_splatting_b200.training import PositionLRConfig, position_lr

GOLDEN = Path(__file__).resolve().parent / "golden" / "ref_gpu_golden.npz"
SCENES = {"plain": dict(n=2000, w=160, h=120, seed=77, adversarial=False, deg=3),
          "adversarial": dict(n=1500, w=160, h=120, seed=78, adversarial=True, deg=2)}
BG = (0.1, 0.2, 0.3)


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def scene_of(c):
    return cugs.synth(c["n"], c["w"], c["h"], seed=c["seed"], adversarial=c["adversarial"])


@pytest.mark.parametrize("name", list(SCENES))
def test_forward_matches_reference_cuda(oracle, golden, name):
    c = SCENES[name]
    f = oracle.render_forward(scene_of(c), deg=c["deg"], bg=BG)
    # index / byte work: bit-exact with the reference (projection.cu, sorting.cu, forward.cu)
    for k in ("radii", "tiles_touched", "gaussian_indices", "tile_ranges", "n_contrib"):
        assert np.array_equal(f[k], golden[f"{name}.{k}"]), k
    assert np.array_equal(f["keys_sorted"].view(np.uint64), golden[f"{name}.keys_sorted"])
    vis = f["radii"] > 0
    for k in ("means_2d", "depths"):  # no transcendental on this path: identical bits
        assert np.array_equal(f[k][vis].view(np.uint32), golden[f"{name}.{k}"][vis].view(np.uint32)), k
    for k, tol in (("cov_2d_inv", 2e-6), ("rgb", 1e-6), ("opacities_act", 5e-7)):
        d = np.abs(f[k][vis] - golden[f"{name}.{k}"][vis])
        rel = d / np.maximum(np.abs(golden[f"{name}.{k}"][vis]), 1.0)
        assert rel.max() <= tol, (k, rel.max())
    assert np.abs(f["color"] - golden[f"{name}.color"]).max() <= 2e-6     # north-star bar is 1e-4
    assert np.abs(f["final_T"] - golden[f"{name}.final_T"]).max() <= 2e-6
    if name == "adversarial":  # the scene must exercise the culls and the A.2 filler-key quirk
        assert (f["radii"] == 0).any() and ((f["radii"] > 0) & (f["tiles_touched"] == 0)).any()


@pytest.mark.parametrize("name", list(SCENES))
def test_backward_matches_reference_cuda(oracle, golden, name):
    c = SCENES[name]
    s = scene_of(c)
    dL = np.random.default_rng(c["seed"] + 1000).uniform(-1, 1, size=(c["h"], c["w"], 3)).astype(np.float32)
    fwd = {k: golden[f"{name}.{k}"] for k in ("tile_ranges", "gaussian_indices", "means_2d", "cov_2d_inv", "rgb",
                                              "opacities_act", "final_T", "n_contrib", "radii")}
    b = oracle.render_backward(s, fwd, dL, deg=c["deg"], bg=BG)
    for k in ("dL_dpositions", "dL_drotations", "dL_dscales", "dL_dopacities", "dL_dsh_coeffs", "dL_dmeans_2d"):
        x, y = b[k].astype(np.float64), golden[f"{name}.{k}"].astype(np.float64)
        rel = np.linalg.norm(x - y) / np.linalg.norm(y)
        assert rel <= 2e-5, (k, rel)                      # north-star bar is 1e-3 (atomic order on the GPU)
        culled = golden[f"{name}.radii"] == 0
        assert not x[culled].any() and not y[culled].any(), k   # culled rows carry exact zeros on both sides


def test_loss_matches_reference_cuda(oracle, golden):
    rng = np.random.default_rng(5)
    x = rng.uniform(size=(40, 56, 3)).astype(np.float32)
    y = rng.uniform(size=(40, 56, 3)).astype(np.float32)
    sc, g = oracle.loss(x, y, 0.2)
    assert np.abs(sc - golden["loss.scalars"]).max() <= 1e-6
    assert np.abs(g - golden["loss.grad"]).max() <= 1e-8 + 1e-4 * np.abs(golden["loss.grad"]).max()


def test_adam_matches_reference_cuda_bitwise(oracle, golden):
    s = cugs.synth(257, 64, 48, seed=79)
    order = ("positions", "rotations", "scales", "opacities", "sh_coeffs")     # fused_adam.hpp group order
    lrs = dict(positions=1.6e-4, sh_coeffs=2.5e-3, opacities=5e-2, rotations=1e-3, scales=5e-3)   # fused_adam.hpp:28-32
    p = {k: np.ascontiguousarray(getattr(s, k)).copy() for k in order}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v = {k: np.zeros_like(x) for k, x in p.items()}
    b1, b2, eps = float(np.float32(0.9)), float(np.float32(0.999)), 1e-15     # (double)config_.beta1, fused_adam.cu:145-146
    for step in range(3):
        flat, o = golden[f"adam.grads{step}"], 0
        bc1, bc2 = 1.0 / (1.0 - b1 ** (step + 1)), 1.0 / (1.0 - b2 ** (step + 1))   # doubles, fused_adam.cu:145-149
        lrs["positions"] = position_lr(step, PositionLRConfig())                     # update_lr, lr_schedule.hpp:49-57
        for k in order:
            g = flat[o:o + p[k].size].reshape(p[k].shape)
            o += p[k].size
            oracle.adam(p[k], g, m[k], v[k], lrs[k], b1, b2, eps, bc1, bc2)
    for k in order:
        assert np.array_equal(p[k].view(np.uint32), golden[f"adam.{k}"].view(np.uint32)), k
