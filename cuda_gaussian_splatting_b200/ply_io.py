"""Gaussian PLY checkpoints in the reference's on-disk format (utils/ply_io.cpp:98-196 writer,
:258-351 reader): binary little-endian, one float32 per property,
``x y z nx ny nz f_dc_0..2 f_rest_0..(3(C-1)-1) opacity scale_0..2 rot_0..3``, the higher-order SH
interleaved ``[k][channel]`` (f_rest_{3(k-1)+ch} = sh[ch][k]). Files are byte-identical to the
reference's for the same model (tests/golden/gaussians_ref.ply), and the reader accepts any property
order, like the reference's name-indexed reader. Host-side format code: numpy only."""
from __future__ import annotations

from pathlib import Path
from typing import Union

import numpy as np
import torch

from .rasterizer import GaussianModel, _check


def _property_names(num_coeffs: int):
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    names += [f"f_rest_{i}" for i in range(3 * (num_coeffs - 1))]
    names += ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    return names


def write_gaussian_ply(path: Union[str, Path], model: GaussianModel) -> bool:
    """utils/ply_io.hpp:53. Returns False for an invalid model (as the reference)."""
    if not model.is_valid():
        return False
    f = lambda t: t.detach().cpu().contiguous().to(torch.float32).numpy()
    pos, sh, opa, scl, rot = f(model.positions), f(model.sh_coeffs), f(model.opacities), f(model.scales), f(model.rotations)
    n, c = pos.shape[0], sh.shape[2]
    names = _property_names(c)
    rows = np.zeros((n, len(names)), dtype="<f4")
    rows[:, 0:3] = pos                                   # normals stay zero
    rows[:, 6:9] = sh[:, :, 0]                           # DC: f_dc_ch
    rows[:, 9:9 + 3 * (c - 1)] = sh[:, :, 1:].transpose(0, 2, 1).reshape(n, 3 * (c - 1))  # [k][ch] interleaved
    o = 9 + 3 * (c - 1)
    rows[:, o] = opa[:, 0]
    rows[:, o + 1:o + 4] = scl
    rows[:, o + 4:o + 8] = rot
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {n}\n"
    header += "".join(f"property float {nm}\n" for nm in names) + "end_header\n"
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(rows.tobytes())
    return True


def read_gaussian_ply(path: Union[str, Path], device="cpu") -> GaussianModel:
    """utils/ply_io.hpp:64. Raises RuntimeError for a missing file / property / short data."""
    try:
        data = Path(path).read_bytes()
    except OSError as e:
        raise RuntimeError(f"Failed to open PLY file: {path}") from e
    end = data.find(b"end_header\n")
    _check(data[:4] == b"ply\n" and end > 0, f"not a PLY file: {path}")
    names, n, in_vertex = [], 0, False
    for line in data[:end].decode("ascii", "replace").split("\n"):
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            _check(tok[1] == "binary_little_endian", "only binary_little_endian PLY files are supported")
        elif tok[0] == "element":
            in_vertex = tok[1] == "vertex"
            if in_vertex:
                n = int(tok[2])
        elif tok[0] == "property" and in_vertex:
            _check(tok[1] in ("float", "float32"), f"property {tok[-1]} is not float")
            names.append(tok[-1])
    idx = {nm: i for i, nm in enumerate(names)}
    num_rest = 0
    while f"f_rest_{num_rest}" in idx:
        num_rest += 1
    c = 1 + num_rest // 3
    body = data[end + len(b"end_header\n"):]
    _check(len(body) >= n * len(names) * 4, "Failed to read PLY binary data")
    rows = np.frombuffer(body, dtype="<f4", count=n * len(names)).reshape(n, len(names))

    def col(nm):
        if nm not in idx:
            raise RuntimeError(f"Missing PLY property: {nm}")
        return rows[:, idx[nm]]

    pos = np.stack([col("x"), col("y"), col("z")], axis=1)
    sh = np.zeros((n, 3, c), np.float32)
    for ch in range(3):
        sh[:, ch, 0] = col(f"f_dc_{ch}")
        for k in range(1, c):
            sh[:, ch, k] = col(f"f_rest_{(k - 1) * 3 + ch}")
    opa = col("opacity")[:, None]
    scl = np.stack([col(f"scale_{i}") for i in range(3)], axis=1)
    rot = np.stack([col(f"rot_{i}") for i in range(4)], axis=1)
    # (np.array copies: the columns are views of the read-only file buffer, a model must own writable memory)
    t = lambda a: torch.from_numpy(np.array(a, dtype=np.float32, order="C")).to(device)
    return GaussianModel(t(pos), t(sh), t(opa), t(rot), t(scl))
