"""Gaussian PLY checkpoint format against golden files written by the UNMODIFIED reference writer
(tests/golden/make_golden_ply.py). CPU only; reference tests: tests/test_ply_io.cpp,
tests/test_gaussian_model.cpp:98-116 (round trip allclose 1e-5)."""
from pathlib import Path

import numpy as np
import pytest
import torch

import cuda_gaussian_splatting_b200 as cugs

GOLDEN = Path(__file__).resolve().parent / "golden"


def model_of(c):
    s = cugs.synth(37, 64, 48, seed=100 + c, num_coeffs=c)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    return cugs.GaussianModel(t(s.positions), t(s.sh_coeffs), t(s.opacities), t(s.rotations), t(s.scales))


@pytest.mark.parametrize("c", [1, 4, 16])
def test_writer_is_byte_identical_to_the_reference(tmp_path, c):
    out = tmp_path / "mine.ply"
    assert cugs.write_gaussian_ply(out, model_of(c))
    assert out.read_bytes() == (GOLDEN / f"gaussians_ref_c{c}.ply").read_bytes()


@pytest.mark.parametrize("c", [1, 4, 16])
def test_reader_reads_the_reference_file(c):
    m, g = model_of(c), cugs.read_gaussian_ply(GOLDEN / f"gaussians_ref_c{c}.ply")
    for a, b in zip((m.positions, m.sh_coeffs, m.opacities, m.rotations, m.scales),
                    (g.positions, g.sh_coeffs, g.opacities, g.rotations, g.scales)):
        assert a.shape == b.shape and torch.equal(a, b)
    assert g.max_sh_degree() == {1: 0, 4: 1, 16: 3}[c] and g.is_valid()


def test_reader_is_name_indexed_and_strict(tmp_path):
    data = (GOLDEN / "gaussians_ref_c1.ply").read_bytes()
    # swap two header lines AND the matching columns: the reader must follow the names
    hdr_end = data.find(b"end_header\n") + len(b"end_header\n")
    hdr = data[:hdr_end].replace(b"property float opacity\n", b"property float TMP\n") \
                        .replace(b"property float scale_0\n", b"property float opacity\n") \
                        .replace(b"property float TMP\n", b"property float scale_0\n")
    p = tmp_path / "swapped.ply"
    p.write_bytes(hdr + data[hdr_end:])
    m, g = model_of(1), cugs.read_gaussian_ply(p)
    assert torch.equal(g.opacities[:, 0], m.scales[:, 0]) and torch.equal(g.scales[:, 0], m.opacities[:, 0])
    bad = tmp_path / "missing.ply"
    bad.write_bytes(data[:hdr_end].replace(b"property float rot_3\n", b"property float other\n") + data[hdr_end:])
    with pytest.raises(RuntimeError):
        cugs.read_gaussian_ply(bad)
    with pytest.raises(RuntimeError):
        cugs.read_gaussian_ply(tmp_path / "does_not_exist.ply")
    short = tmp_path / "short.ply"
    short.write_bytes(data[:-8])
    with pytest.raises(RuntimeError):
        cugs.read_gaussian_ply(short)


def test_invalid_model_is_not_written(tmp_path):
    m = model_of(4)
    m.opacities = m.opacities[:-1]
    assert cugs.write_gaussian_ply(tmp_path / "x.ply", m) is False


@pytest.mark.parametrize("degree,n", [(3, 50), (0, 20), (2, 30), (3, 0)])
def test_roundtrip(tmp_path, degree, n):  # tests/test_gaussian_model.cpp:98-160 (degrees 3 / 0 / 2, empty model)
    import torch
    c = (degree + 1) ** 2
    g = torch.Generator().manual_seed(degree * 100 + n)
    m = cugs.GaussianModel(torch.randn(n, 3, generator=g), torch.randn(n, 3, c, generator=g),
                           torch.randn(n, 1, generator=g), torch.randn(n, 4, generator=g),
                           torch.randn(n, 3, generator=g))
    path = tmp_path / f"d{degree}_{n}.ply"
    assert cugs.write_gaussian_ply(path, m) and path.exists()
    back = cugs.read_gaussian_ply(path)
    assert back.is_valid() and back.num_gaussians() == n and back.max_sh_degree() == degree
    for a, b in zip((m.positions, m.sh_coeffs, m.opacities, m.rotations, m.scales),
                    (back.positions, back.sh_coeffs, back.opacities, back.rotations, back.scales)):
        assert torch.equal(a, b)          # float32 in, float32 on disk: exact
    back.positions.add_(1.0)              # the loaded model owns writable memory
