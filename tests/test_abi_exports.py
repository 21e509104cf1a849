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


def test_capacity_sized_scratch_serves_every_smaller_count():
    """A buffer sized for a pair CAPACITY must be large enough for every smaller pair count, also across the
    element count from which the packed sort switches to 8192-element tiles (half the look-back state): the size
    queries are monotone in their count argument."""
    from cuda_gaussian_splatting_b200 import _lib
    lib = _lib.load_library()
    limit = 8 << 20
    counts = [1, 4095, 4096, 4097, 1_000_000, limit - 8193, limit - 1, limit, limit + 1, limit + 8193, 2 * limit,
              18_596_764, 124_800_000]
    for fn in (lambda c: lib.cugs_b200_render_pair_scratch_bytes(c),
               lambda c: lib.cugs_b200_sort_packed_temp_bytes(c, 13, 8160),
               lambda c: lib.cugs_b200_sort_packed_temp_bytes(c, 32, 0),
               lambda c: lib.cugs_b200_render_workspace_bytes(c, 0),
               lambda c: lib.cugs_b200_trainer_workspace_bytes(1000, 16, 640, 360, c, 2)):
        sizes = [int(fn(c)) for c in counts]
        assert all(b >= a > 0 for a, b in zip(sizes, sizes[1:])), sizes


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


def header_prototypes():
    """name -> (return type, [parameter declarations]) for every prototype of the header."""
    text = (ROOT / "include" / "cugs_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(cugs_b200_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        protos[name] = (ret, plist)
    return protos


def test_ctypes_signatures_match_the_header_prototypes():
    """Arity and the pointer / integer / float class of every parameter: a ctypes table that drifts from the
    header corrupts arguments silently."""
    import ctypes as C
    from cuda_gaussian_splatting_b200 import _lib
    protos = header_prototypes()
    assert set(protos) == set(declared_symbols())
    pointer_types = (C.c_void_p, C.c_char_p)

    def kind(ct):
        if ct is None:
            return "void"
        if ct in pointer_types or isinstance(ct, type) and issubclass(ct, (C._Pointer, C.Array)):
            return "ptr"
        if ct in (C.c_float, C.c_double):
            return "float"
        return "int"

    def ckind(decl):
        if "*" in decl or "[" in decl:
            return "ptr"
        if re.search(r"\b(float|double)\b", decl):
            return "float"
        return "int"

    for name, (ret, plist) in protos.items():
        res, args = _lib.SIGNATURES[name]
        assert len(args) == len(plist), f"{name}: header has {len(plist)} parameters, ctypes table {len(args)}"
        for i, (decl, ct) in enumerate(zip(plist, args)):
            assert ckind(decl) == kind(ct), f"{name}: parameter {i} `{decl}` is bound as {ct}"
        want = "void" if ret == "void" else ("ptr" if "*" in ret else "int")
        assert kind(res) == want, f"{name}: return type `{ret}` is bound as {res}"
        # 64-bit sizes must not be bound as 32-bit ints
        for decl, ct in zip(plist, args):
            if re.search(r"\b(int64_t|uint64_t|size_t)\b", decl) and "*" not in decl:
                assert C.sizeof(ct) == 8, f"{name}: `{decl}` is bound as a {C.sizeof(ct)}-byte integer"


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """sizeof / offsetof of the two structs that cross the boundary, as gcc lays them out, against the ctypes
    mirrors in _lib.py (also proves the header is plain C)."""
    import ctypes as C
    import subprocess
    from cuda_gaussian_splatting_b200 import _lib
    structs = {"cugs_view_t": _lib.CugsView, "cugs_densify_config_t": _lib.CugsDensifyConfig}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "cugs_b200.h"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def test_error_convention_without_a_device():
    """0 = ok, < 0 = argument error (CUGS_ERR_*), > 0 = cudaError_t; nothing throws, nothing crashes on a
    null handle (SURVEY 8b error conventions)."""
    import ctypes as C
    import torch
    from cuda_gaussian_splatting_b200 import _lib
    lib = _lib.load_library()
    p = C.c_int64(0)
    assert lib.cugs_b200_scan(None, None, 10, None, None, None, C.byref(p), None, 0) == -1
    assert lib.cugs_b200_render_plan(None, None, 10, None, *([None] * 11), None, 0, C.byref(p)) == -1
    assert lib.cugs_b200_accumulate_stats(None, None, 10, None, None, None, None, None) == -1
    assert lib.cugs_b200_densify_classify(None, None, 10, None, None, None, None, None, None, None, C.byref(p),
                                          None, 0) == -1
    assert lib.cugs_b200_mcmc_relocate(None, None, 10, 16, None, None, None, None, None, 0.005, 1, 1.0, 1, 1, None,
                                       None, None, None, 0) == -1
    assert lib.cugs_b200_last_error(None) == b"null handle"
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert lib.cugs_b200_create(0, C.byref(h)) > 0 and not h.value      # a cudaError_t, no handle, no fallback


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing in the package, the wrapper or the C sources may import,
    link or open it (bench.py and __graft_entry__.smoke() may, as checker / CPU baseline only)."""
    offenders = []
    for path in list((ROOT / "cuda_gaussian_splatting_b200").rglob("*")) + list((ROOT / "wrapper").rglob("*")) + \
            list((ROOT / "include").rglob("*")):
        if path.is_file() and path.suffix in {".py", ".cu", ".cuh", ".cpp", ".h", ".hpp"}:
            text = path.read_text(errors="replace")
            if re.search(r"\boracle_py\b|\bcugs_oracle\b|\bcugs_ref\b|from\s+oracle|import\s+oracle|oracle/_ref", text):
                offenders.append(str(path.relative_to(ROOT)))
    assert not offenders, offenders
    build_sh = (ROOT / "build.sh").read_text()
    assert "oracle" not in build_sh
