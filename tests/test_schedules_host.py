"""The schedules the step driver evaluates on the host (CPU only), ported from the reference's own
tests: position learning rate and SH-degree warm-up (tests/test_training.cpp:29-81) and the MCMC
noise learning rate (tests/test_mcmc.cpp:84-120)."""
import math

import numpy as np

from cuda_gaussian_splatting_b200.training import (MCMCConfig, PositionLRConfig, active_sh_degree_for_step,
                                                   mcmc_noise_lr, position_lr)


def test_position_lr_endpoints():  # test_training.cpp:29-45
    cfg = PositionLRConfig()
    assert abs(position_lr(0, cfg) - 1.6e-4) <= 1e-8
    assert abs(position_lr(cfg.max_steps, cfg) - 1.6e-6) <= 1e-10
    assert abs(position_lr(cfg.max_steps + 1000, cfg) - cfg.lr_final) <= 1e-10


def test_position_lr_monotonic_and_midpoint():  # test_training.cpp:47-63
    cfg = PositionLRConfig()
    prev = position_lr(0, cfg)
    for step in range(100, cfg.max_steps + 1, 100):
        cur = position_lr(step, cfg)
        assert cur < prev, step
        prev = cur
    mid = position_lr(cfg.max_steps // 2, cfg)
    assert abs(mid - math.sqrt(cfg.lr_init * cfg.lr_final)) <= 1e-8      # geometric mean at half way


def test_position_lr_is_computed_in_float32_like_the_reference():  # lr_schedule.hpp:49-57
    cfg = PositionLRConfig()
    for step in (1, 777, 15000, 29999):
        t = np.float32(step) / np.float32(cfg.max_steps)
        want = np.float32(cfg.lr_init) * np.float32(math.exp(np.float32(t * np.float32(
            math.log(np.float32(cfg.lr_final) / np.float32(cfg.lr_init))))))
        assert np.float32(position_lr(step, cfg)) == np.float32(want)


def test_active_sh_degree_schedule():  # test_training.cpp:65-81
    for step, want in ((0, 0), (500, 0), (999, 0), (1000, 1), (1999, 1), (2000, 2), (3000, 3), (30000, 3)):
        assert active_sh_degree_for_step(step, 3) == want
    assert active_sh_degree_for_step(5000, 1) == 1
    assert active_sh_degree_for_step(0, 0) == 0 and active_sh_degree_for_step(5000, 0) == 0


def test_mcmc_noise_lr_endpoints_and_decay():  # test_mcmc.cpp:84-120
    cfg = MCMCConfig(noise_lr_init=5e5, noise_lr_final=1e3, noise_lr_max_steps=30000)
    assert np.float32(mcmc_noise_lr(0, cfg)) == np.float32(5e5)
    assert np.float32(mcmc_noise_lr(30000, cfg)) == np.float32(1e3)
    assert np.float32(mcmc_noise_lr(50000, cfg)) == np.float32(1e3)
    prev = mcmc_noise_lr(0, cfg)
    for step in range(1000, 30001, 1000):
        cur = mcmc_noise_lr(step, cfg)
        assert cur < prev, step
        prev = cur
