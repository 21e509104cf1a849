"""Host logic of the model-resizing schedules (CPU only): the schedule predicates ported from the
reference's own tests (tests/test_densification.cpp:48-104, tests/test_mcmc.cpp:46-82) and the
threshold arithmetic handed to the kernels."""
import numpy as np

import cuda_gaussian_splatting_b200 as cugs


def test_should_densify_boundaries():  # test_densification.cpp:48-80
    ctrl = cugs.DensificationController(cugs.DensificationConfig(densify_from=500, densify_until=15000,
                                                                 densify_every=100), 10.0, 0, "cpu")
    for s in (0, 100, 400, 499, 501, 550, 999, 15100, 20000):
        assert not ctrl.should_densify(s), s
    for s in (500, 600, 1000, 14900, 15000):
        assert ctrl.should_densify(s), s


def test_should_reset_opacity():  # test_densification.cpp:82-104
    ctrl = cugs.DensificationController(cugs.DensificationConfig(densify_from=500, opacity_reset_every=3000), 10.0, 0,
                                        "cpu")
    assert not ctrl.should_reset_opacity(0)
    assert all(ctrl.should_reset_opacity(s) for s in (3000, 6000, 9000))
    assert not ctrl.should_reset_opacity(3001) and not ctrl.should_reset_opacity(4000)
    off = cugs.DensificationController(cugs.DensificationConfig(opacity_reset_every=0), 10.0, 0, "cpu")
    assert not off.should_reset_opacity(3000)


def test_should_relocate_boundaries():  # test_mcmc.cpp:46-82
    cfg = cugs.MCMCConfig(relocate_from=500, relocate_until=15000, relocate_every=100)
    for s in (0, 100, 499, 501, 550, 15100, 20000):
        assert not cugs.mcmc_should_relocate(s, cfg), s
    for s in (500, 600, 1000, 15000):
        assert cugs.mcmc_should_relocate(s, cfg), s


def test_native_thresholds_follow_the_reference_arithmetic():
    ctrl = cugs.DensificationController(cugs.DensificationConfig(), 7.3, 0, "cpu")
    early, late = ctrl._native_config(3000), ctrl._native_config(3001)
    assert early.apply_size_pruning == 0 and late.apply_size_pruning == 1      # step > opacity_reset_every (:416-417)
    assert np.float32(early.size_threshold) == np.float32(0.01) * np.float32(7.3)   # percent_dense * extent, float
    assert np.float32(early.ws_threshold) == np.float32(0.1) * np.float32(7.3)
    assert early.max_screen_size == 20.0 and np.float32(early.grad_threshold) == np.float32(0.0002)
    assert ctrl._native_config(10**6).apply_size_pruning == 1
    off = cugs.DensificationController(cugs.DensificationConfig(opacity_reset_every=0), 1.0, 0, "cpu")
    assert off._native_config(10**6).apply_size_pruning == 0


def test_budget_cap_keeps_the_highest_average_gradients():  # densification.cpp:128-137, :196-213
    import torch
    ctrl = cugs.DensificationController(cugs.DensificationConfig(max_gaussians=12), 1.0, 10, "cpu")
    ctrl.grad_accum = torch.tensor([9., 1., 8., 2., 7., 3., 6., 4., 5., 0.])
    ctrl.grad_count = torch.tensor([1., 1., 2., 1., 1., 1., 1., 1., 1., 0.])     # averages: 9 1 4 2 7 3 6 4 5 0
    flags = torch.tensor([3, 3, 3, 1, 3, 5, 5, 1, 3, 1], dtype=torch.uint8)       # KEEP=1, CLONE=2, SPLIT=4
    ctrl._cap(flags, 2, 3)                                                         # three best clone candidates
    assert flags.tolist() == [3, 1, 1, 1, 3, 5, 5, 1, 3, 1]                        # rows 0 (9), 4 (7), 8 (5) stay
    ctrl._cap(flags, 4, 1)                                                         # one split: row 6 (6) beats row 5 (3)
    assert flags.tolist() == [3, 1, 1, 1, 3, 1, 5, 1, 3, 1]
    ctrl._cap(flags, 2, 0)                                                         # no budget: all candidates dropped
    assert flags.tolist() == [1, 1, 1, 1, 1, 1, 5, 1, 1, 1]


def test_lazy_resize_keeps_the_controller_intact():
    """densification.cpp:66-68: accumulate_gradients re-initialises the accumulators when N changed. The
    re-initialisation must not run the controller's own constructor (config / scene_extent would be
    clobbered)."""
    import torch
    cfg = cugs.DensificationConfig(grad_threshold=0.123)
    ctrl = cugs.DensificationController(cfg, 7.5, 5, "cpu", seed=99)
    ctrl.grad_accum += 1.0
    ctrl.accumulate_gradients(torch.zeros((0, 2)), torch.zeros((0,), dtype=torch.int32))  # resize 5 -> 0, nothing to add
    assert ctrl.grad_accum.shape == (0,) and ctrl.grad_count.shape == (0,) and ctrl.max_radii_2d.shape == (0,)
    assert ctrl.config is cfg and ctrl.scene_extent == 7.5 and ctrl.seed == 99
    ctrl.reset_accumulators(3, "cpu")
    assert ctrl.grad_accum.shape == (3,) and float(ctrl.grad_accum.sum()) == 0.0 and ctrl.config is cfg
