"""Argument validation of the host mirror (CPU only): the error behaviour the reference's tests pin with
EXPECT_THROW(..., c10::Error) -- tests/test_loss.cpp:143-170, tests/test_sh.cpp:127-143 -- raised here as
RuntimeError before anything is launched. A CPU tensor is always an error: there is no CPU fallback."""
import pytest
import torch

import cuda_gaussian_splatting_b200 as cugs


def test_loss_input_validation():  # test_loss.cpp:143-170 (the checks that do not need a device)
    cpu = torch.rand(32, 32, 3)
    for fn in (cugs.l1_loss, cugs.ssim_loss, cugs.combined_loss, cugs.ssim):
        with pytest.raises(RuntimeError):
            fn(cpu, cpu)                                        # CPU tensor (should require CUDA)
    with pytest.raises(RuntimeError):
        cugs.l1_loss(cpu, torch.rand(64, 64, 3))                # mismatched shapes
    with pytest.raises(RuntimeError):
        cugs.l1_loss(torch.rand(32, 32, 4), torch.rand(32, 32, 4))   # wrong number of channels
    with pytest.raises(RuntimeError):
        cugs.ssim(cpu, cpu, window_size=10)                     # even window size


def test_sh_input_validation():  # test_sh.cpp:127-143, through the CUDA entry point's host checks
    with pytest.raises(RuntimeError):
        cugs.evaluate_sh_cuda(1, torch.zeros(5, 3, 4), torch.zeros(3, 3))     # batch sizes differ / not CUDA
    with pytest.raises(RuntimeError):
        cugs.evaluate_sh_cuda(1, torch.zeros(1, 3, 1), torch.zeros(1, 3))     # not enough coefficients
    for degree in (4, -1):
        with pytest.raises(RuntimeError):
            cugs.evaluate_sh_cuda(degree, torch.zeros(1, 3, 16), torch.zeros(1, 3))
    with pytest.raises(RuntimeError):
        cugs.evaluate_sh_backward_cuda(5, torch.zeros(1, 3, 16), torch.zeros(1, 3), torch.zeros(1, 3))


def test_fused_adam_refuses_host_parameters():  # fused_adam.cu:84-92
    m = cugs.GaussianModel(torch.randn(4, 3), torch.randn(4, 3, 16), torch.randn(4, 1), torch.randn(4, 4),
                           torch.randn(4, 3))
    with pytest.raises(RuntimeError):
        cugs.FusedAdam(m)


def test_projection_and_sort_refuse_host_tensors():
    cam = cugs.CameraInfo(64, 48, 50.0, 50.0, 32.0, 24.0)
    with pytest.raises(RuntimeError):
        cugs.project_gaussians(torch.randn(4, 3), torch.randn(4, 4), torch.randn(4, 3), torch.randn(4, 1),
                               torch.randn(4, 3, 16), cam, 3)
    with pytest.raises(RuntimeError):
        cugs.sort_gaussians(torch.zeros(4, 2), torch.zeros(4), torch.zeros(4, dtype=torch.int32),
                            torch.zeros(4, dtype=torch.int32), 64, 48)
