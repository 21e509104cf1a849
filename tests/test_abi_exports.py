"""The C-ABI library loads on a CPU-only box and exports every symbol the header declares."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "cugs_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cugs_b200_\w+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for need in ["cugs_b200_preprocess_fwd", "cugs_b200_scan", "cugs_b200_duplicate_with_keys",
                 "cugs_b200_sort_pairs", "cugs_b200_tile_ranges", "cugs_b200_blend_fwd", "cugs_b200_blend_bwd",
                 "cugs_b200_preprocess_bwd", "cugs_b200_render_plan", "cugs_b200_render_finish",
                 "cugs_b200_render_backward", "cugs_b200_loss_l1_ssim", "cugs_b200_adam_step",
                 "cugs_b200_accumulate_stats"]:
        assert need in syms


def test_library_exports_every_declared_symbol():
    from cuda_gaussian_splatting_b200 import _lib
    lib = _lib.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/cugs_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    header = (ROOT / "include" / "cugs_b200.h").read_text()
    assert lib.cugs_b200_abi_version() == int(re.search(r"#define CUGS_B200_ABI_VERSION (\d+)", header).group(1))


def test_size_queries_need_no_gpu():
    from cuda_gaussian_splatting_b200 import _lib
    lib = _lib.load_library()
    assert lib.cugs_b200_scan_temp_bytes(3_000_000) > 0
    assert lib.cugs_b200_sort_temp_bytes(15_000_000) > 0
    small = lib.cugs_b200_render_workspace_bytes(1000, 0)
    big = lib.cugs_b200_render_workspace_bytes(1000, 100000)
    assert big > small > 0
    assert lib.cugs_b200_loss_workspace_bytes(1920, 1080) >= 9 * 1920 * 1080 * 4
    # 45 key bits at 1080p -> 6 passes; trimmed depth range (26 bits) -> 5; full 64 bits -> 8
    assert lib.cugs_b200_sort_num_passes(32, 13) == 6
    assert lib.cugs_b200_sort_num_passes(26, 13) == 5
    assert lib.cugs_b200_sort_num_passes(32, 32) == 8


def test_no_cpu_fallback_without_gpu():
    """On a box without a B200 the product must fail loudly, not fall back."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cuda_gaussian_splatting_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.handle(0)
    import cuda_gaussian_splatting_b200 as m
    s = m.synth(16, 64, 48)
    model = m.GaussianModel(*(torch.from_numpy(a) for a in (s.positions, s.sh_coeffs, s.opacities, s.rotations, s.scales)))
    with pytest.raises(RuntimeError):
        m.render(model, s.camera, m.RenderSettings())
