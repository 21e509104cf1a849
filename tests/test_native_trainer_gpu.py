"""The C++ training step (csrc/trainer.cu, cugs_b200_trainer_*): Trainer::train_step of the reference
(training/trainer.cpp:178-316) sequenced by the library, without host synchronisation, replayed as a CUDA
graph. Parity: against SyntheticTrainer (the Python driver, itself pinned against the reference's own
sequence in test_gpu_parity.py::test_training_step_matches_reference_sequence and test_dropin_gpu.py), and
directly against the reference sequence compiled into oracle/_ref."""
import numpy as np
import pytest

import cuda_gaussian_splatting_b200 as cugs
from conftest import to_torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("no CUDA device")
    return t


def _setup(torch, n=4000, W=176, H=120, views=3, seed=31, sigma=4.0):
    scene = cugs.synth(n, W, H, seed=seed, sigma_px=sigma)
    cams = [scene.camera] + cugs.ring_cameras(scene, views - 1, radius_frac=0.05)
    rng = np.random.default_rng(9)
    targets = [torch.from_numpy(rng.uniform(size=(H, W, 3)).astype(np.float32)).cuda() for _ in range(views)]
    return scene, cams, targets


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


PARAMS = ("positions", "sh_coeffs", "opacities", "rotations", "scales")


@pytest.mark.parametrize("use_graph,frames", [(True, 2), (False, 1), (True, 1), (False, 2)])
def test_native_step_matches_python_driver(torch, use_graph, frames):
    """Same steps, two drivers: parameters agree to summation-order noise (the backward blend adds with float
    atomics), visibility counts exactly; crosses the SH-degree switch at step 1000 (graph re-capture)."""
    scene, cams, targets = _setup(torch)
    cfg = cugs.TrainConfig()
    py_model, nat_model = to_torch(scene), to_torch(scene)
    py = cugs.SyntheticTrainer(py_model, cams, targets, cfg)
    nat = cugs.NativeTrainer(nat_model, cams, targets, cfg, frames_in_flight=frames, use_graph=use_graph)
    steps = list(range(996, 1004)) + [3000, 3001, 3002]
    for s in steps:
        py_sc = py.train_step(s)
        nat.train_step(s)
        sc, ok, pmax, nv = nat.result()
        assert ok and nv == len(cams) and 0 < pmax <= nat.pair_capacity
        assert np.allclose(sc, py_sc.cpu().numpy(), rtol=2e-5, atol=1e-6), (s, sc, py_sc)
    for nm in PARAMS:
        assert _rel(getattr(nat_model, nm), getattr(py_model, nm)) <= 2e-5, nm
    assert torch.equal(nat.stats.grad_count, py.stats.grad_count)
    assert torch.equal(nat.stats.max_radii_2d, py.stats.max_radii_2d)
    assert _rel(nat.stats.grad_accum, py.stats.grad_accum) <= 1e-4
    assert nat.adam_steps == len(steps) == py.optimizer.step_count
    nat.close()


def test_native_step_vs_reference_sequence(ref, torch):
    """The library's step against the reference's own sequence (render -> combined_loss + autograd ->
    render_backward -> FusedAdam::step, oracle/ref_harness.cpp: ref_train_steps) on the same model."""
    scene, cams, targets = _setup(torch, n=2000, W=160, H=112, views=1, seed=33)
    if not hasattr(ref, "train_steps"):
        pytest.skip("harness without train_steps")
    mine, theirs = to_torch(scene), to_torch(scene)
    nat = cugs.NativeTrainer(mine, cams, targets, cugs.TrainConfig())
    steps = 12
    first = 3000
    for s in range(first, first + steps):
        nat.train_step(s)
    sc, ok, _, _ = nat.result()
    assert ok
    # the reference sequence does not modify its inputs: it returns {losses, positions, sh, opacities, rotations, scales}
    r = ref.train_steps(theirs.positions, theirs.sh_coeffs, theirs.opacities, theirs.rotations, theirs.scales,
                        cams[0].as_ref_list(), targets[0], [0.0, 0.0, 0.0], 3, 0.2, first, steps)
    for nm, rt in zip(("positions", "sh_coeffs", "opacities", "rotations", "scales"), r[1:]):
        assert _rel(getattr(mine, nm), rt) <= 1e-4, nm
    assert abs(sc[0] - float(r[0][-1])) <= 1e-4
    nat.close()


def test_mcmc_mode_matches_python_driver(torch):
    scene, cams, targets = _setup(torch, views=2)
    # (noise scale chosen for the synthetic scene's units: the default 5e5 is meant for COLMAP-scale scenes)
    cfg = cugs.TrainConfig(mcmc=cugs.MCMCConfig(noise_lr_init=2e-2, noise_lr_final=1e-3))
    a, b = to_torch(scene), to_torch(scene)
    py = cugs.SyntheticTrainer(a, cams, targets, cfg)
    nat = cugs.NativeTrainer(b, cams, targets, cfg)
    for s in range(3000, 3006):
        py.train_step(s)
        nat.train_step(s)
    assert nat.result()[1]
    for nm in PARAMS:
        assert _rel(getattr(b, nm), getattr(a, nm)) <= 2e-5, nm
    nat.close()


def test_overflow_skips_the_update_and_grow_recovers(torch):
    scene, cams, targets = _setup(torch, views=2)
    model, twin = to_torch(scene), to_torch(scene)
    nat = cugs.NativeTrainer(model, cams, targets, cugs.TrainConfig(), pair_capacity=2048)
    before = [getattr(model, nm).clone() for nm in PARAMS]
    nat.train_step(3000)
    sc, ok, pmax, _ = nat.result()
    assert not ok and pmax > 2048
    for nm, old in zip(PARAMS, before):
        assert torch.equal(getattr(model, nm), old), f"{nm} changed although the step overflowed"
    assert all(float(m.abs().max()) == 0.0 for m in nat.m), "Adam moments changed although the step overflowed"
    got = nat.train_step_checked(3000)            # grows, repeats
    full = cugs.NativeTrainer(twin, cams, targets, cugs.TrainConfig())
    full.train_step(3000)
    want = full.result()[0]
    assert np.allclose(got, want, rtol=1e-5)
    for nm in PARAMS:
        assert _rel(getattr(model, nm), getattr(twin, nm)) <= 2e-5, nm
    assert nat.adam_steps == 1
    nat.close(); full.close()


@pytest.mark.parametrize("use_graph", [True, False])
def test_host_targets_copied_inside_the_step(torch, use_graph):
    """End-to-end mode: the target images live in pinned host memory and are copied inside every step (copy
    stream, under the rendering of the same view). Same losses and parameters as with device-resident targets;
    changing the host image between steps changes the next step's loss (the copy really happens per step)."""
    scene, cams, targets = _setup(torch, views=3)
    a, b = to_torch(scene), to_torch(scene)
    dev_tr = cugs.NativeTrainer(a, cams, targets, cugs.TrainConfig(), use_graph=use_graph)
    host = [t.cpu().pin_memory() for t in targets]
    staging = [torch.empty_like(t) for t in targets]
    host_tr = cugs.NativeTrainer(b, cams, staging, cugs.TrainConfig(), use_graph=use_graph)
    host_tr.set_views(cams, staging, None, host)
    for s in range(3000, 3005):
        dev_tr.train_step(s)
        host_tr.train_step(s)
        assert np.allclose(host_tr.result()[0], dev_tr.result()[0], rtol=2e-5, atol=1e-6), s
    for nm in PARAMS:
        assert _rel(getattr(b, nm), getattr(a, nm)) <= 2e-5, nm
    before = host_tr.result()[0][0]
    host[0].zero_()                      # a different image in the same pinned buffer: no re-capture needed
    host_tr.train_step(3005)
    assert abs(host_tr.result()[0][0] - before) > 1e-3
    dev_tr.close(); host_tr.close()
