"""Generates tests/golden/sh_cpu_golden.npz by running the UNMODIFIED reference function
cugs::evaluate_sh_cpu (reference src/core/sh.cpp:8-87) through oracle/_ref/cugs_ref*.so
(built by oracle/Makefile.ref). Runs in the build container (no GPU needed).

    python tests/golden/make_golden_cpu.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
import cugs_ref  # noqa: E402

rng = np.random.default_rng(20261018)
out = {}
for deg in range(4):
    n = 64
    sh = rng.normal(0, 0.7, size=(n, 3, 16)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rgb = cugs_ref.evaluate_sh_cpu(deg, torch.from_numpy(sh), torch.from_numpy(d)).numpy()
    out[f"sh_{deg}"] = sh
    out[f"dir_{deg}"] = d
    out[f"rgb_{deg}"] = rgb
np.savez_compressed(ROOT / "tests" / "golden" / "sh_cpu_golden.npz", **out)
print("wrote sh_cpu_golden.npz")
