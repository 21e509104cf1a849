"""Generates tests/golden/gaussians_ref_c{1,4,16}.ply with the UNMODIFIED reference writer
cugs::write_gaussian_ply (reference src/utils/ply_io.cpp:98-196) through oracle/_ref/cugs_ref*.so,
from the seeded synthetic scenes the test regenerates. Runs in the build container (CPU only).

    python tests/golden/make_golden_ply.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
import cugs_ref  # noqa: E402
import cuda_gaussian_splatting_b200 as cugs  # noqa: E402  (synth only)

for c in (1, 4, 16):
    s = cugs.synth(37, 64, 48, seed=100 + c, num_coeffs=c)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    out = ROOT / "tests" / "golden" / f"gaussians_ref_c{c}.ply"
    assert cugs_ref.write_gaussian_ply(str(out), t(s.positions), t(s.sh_coeffs), t(s.opacities), t(s.rotations), t(s.scales))
    back = cugs_ref.read_gaussian_ply(str(out))
    assert torch.equal(back[1], t(s.sh_coeffs))
    print("wrote", out, out.stat().st_size, "bytes")
