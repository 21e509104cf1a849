"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` covers the CPU oracle (against known answers and golden vectors), the host logic
and that the C-ABI library loads and exports every symbol include/cugs_b200.h declares.
`-m gpu` holds the parity tests proper; they call through the C ABI on a real B200.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def cugs():
    import cuda_gaussian_splatting_b200 as m
    return m


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference kernels (oracle/_ref, built by oracle/Makefile.ref); GPU tests
    that use it are skipped when the module was not built / shipped."""
    ref_dir = ROOT / "oracle" / "_ref"
    sys.path.insert(0, str(ref_dir))
    try:
        import torch  # noqa: F401  (libtorch must be loaded first)
        import cugs_ref
        return cugs_ref
    except Exception as e:  # pragma: no cover
        pytest.skip(f"compiled reference oracle/_ref/cugs_ref*.so not available: {e}")


def to_torch(scene, device="cuda"):
    import torch
    import cuda_gaussian_splatting_b200 as m
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return m.GaussianModel(t(scene.positions), t(scene.sh_coeffs), t(scene.opacities), t(scene.rotations),
                           t(scene.scales))


@pytest.fixture(autouse=True)
def _seed_torch():
    """The reference controllers draw from torch's global generator (randn_like, multinomial): seed it so the
    statistical checks on their draws see the same numbers in every run."""
    import torch
    torch.manual_seed(20251018)
    yield
