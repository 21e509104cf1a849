"""GaussianModel validity rules (core/gaussian.hpp:67-102), ported from the reference's
tests/test_gaussian_model.cpp:30-78. CPU only: the checks are host logic on tensor shapes."""
import torch

import cuda_gaussian_splatting_b200 as cugs


def make_test_model(n, degree=3):  # test_gaussian_model.cpp:16-27
    c = (degree + 1) ** 2
    return cugs.GaussianModel(torch.randn(n, 3), torch.randn(n, 3, c), torch.randn(n, 1), torch.randn(n, 4),
                              torch.randn(n, 3))


def test_valid_after_creation_and_counts():  # :36-47
    m = make_test_model(100)
    assert m.is_valid() and m.num_gaussians() == 100
    for d in range(4):
        assert make_test_model(10, d).max_sh_degree() == d


def test_invalid_shapes_detected():  # :56-67
    m = make_test_model(10)
    saved = m.positions
    m.positions = torch.randn(10, 4)
    assert not m.is_valid()
    m.positions = saved
    assert m.is_valid()
    m.opacities = torch.randn(5, 1)
    assert not m.is_valid()
    m = make_test_model(10)
    m.sh_coeffs = torch.randn(10, 16, 3)       # channel-major [N,3,C] is the contract
    assert not m.is_valid()
    m = make_test_model(10)
    m.rotations = torch.randn(10, 3)
    assert not m.is_valid()


def test_empty_model_is_valid():  # :69-78
    m = cugs.GaussianModel(torch.zeros(0, 3), torch.zeros(0, 3, 16), torch.zeros(0, 1), torch.zeros(0, 4),
                           torch.zeros(0, 3))
    assert m.is_valid() and m.num_gaussians() == 0 and m.max_sh_degree() == 3


def test_render_refuses_host_tensors_without_a_cpu_fallback():
    # rasterizer.cpp:27-28: "GaussianModel must be on CUDA device"; there is no CPU path to fall back to
    m = make_test_model(4)
    cam = cugs.CameraInfo(64, 48, 50.0, 50.0, 32.0, 24.0)
    try:
        cugs.render(m, cam, cugs.RenderSettings())
    except RuntimeError as e:
        assert "CUDA" in str(e)
    else:
        raise AssertionError("render() accepted CPU tensors")
