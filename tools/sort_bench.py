#!/usr/bin/env python
"""Sort micro-benchmark (GPU box): CUB as the reference calls it vs trimmed CUB vs this repository's
hand-written onesweep, on the REAL (tile|depth, index) pairs of a synthetic workload.

    python tools/sort_bench.py [workload ...]        # default: B E   (P = 18.6 M and 124.8 M)

Arms (reference call: rasterizer/sorting.cu:190-211):
  cub64      cub::DeviceRadixSort::SortPairs(u64 key, i32 value), bits [0, 64)      = the reference call
  cub_trim   the same with end_bit = 32 + ceil(log2(tiles))                          = the SURVEY's bar
  stage      cugs_b200_sort_pairs (radix_sort.cu): 64-bit keys, only the bits that can differ
  fused      the render path's sort stage: depth sort of the N Gaussians before duplicateWithKeys +
             tile sort of the P packed pairs + tile ranges from the tile histogram (tile_binning.cu),
             timed with the library's stage events inside cugs.render()
All four produce the same order (checked here bit for bit before timing). Times are CUDA events on the
launching stream, median of 20 after 5 warm-ups; the input pairs are restored outside the timed region.
Bytes for the GB/s column: SURVEY §8(d) (8 + 24 p) B/pair with p = passes of THAT arm's formulation.
The CUB library is a benchmark baseline only (tools/cub_sort_baseline.cu); the product never links it.
"""
import ctypes as C
import json
import math
import statistics
import subprocess
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import cuda_gaussian_splatting_b200 as cugs  # noqa: E402
from cuda_gaussian_splatting_b200 import _lib  # noqa: E402

sys.path.insert(0, str(ROOT))
from bench import WORKLOADS  # noqa: E402


def load_cub():
    so = ROOT / "tools" / "_build" / "libcub_sort_baseline.so"
    if not so.exists():
        so.parent.mkdir(exist_ok=True)
        subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
                               "-Xcompiler", "-fPIC", str(ROOT / "tools" / "cub_sort_baseline.cu"), "-o", str(so)])
    lib = C.CDLL(str(so))
    lib.cub_sort_pairs_temp_bytes.restype = C.c_size_t
    lib.cub_sort_pairs_temp_bytes.argtypes = [C.c_int64, C.c_int, C.c_int]
    lib.cub_sort_pairs.restype = C.c_int
    lib.cub_sort_pairs.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_int, C.c_void_p]
    return lib


def timed(fn, restore=None, warm=5, iters=20):
    ts = []
    for it in range(warm + iters):
        if restore:
            restore()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= warm:
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts)


def run(workload: str) -> dict:
    n, W, H, seed, desc = WORKLOADS[workload]
    dev = torch.device("cuda", 0)
    scene = cugs.synth(n, W, H, seed=seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    model = cugs.GaussianModel(t(scene.positions), t(scene.sh_coeffs), t(scene.opacities), t(scene.rotations),
                               t(scene.scales))
    lib, h = _lib.load_library(), _lib.handle(0)
    s = torch.cuda.current_stream().cuda_stream
    T = cugs.rasterizer.num_tiles(W, H)
    tile_bits = max(0, math.ceil(math.log2(T)))

    # the real unsorted pairs: project -> scan -> duplicateWithKeys (stage functions)
    pr = cugs.project_gaussians(model.positions, model.rotations, model.scales, model.opacities, model.sh_coeffs,
                                scene.camera, 3)
    offsets = torch.empty((n,), dtype=torch.int32, device=dev)
    tmp = torch.empty((lib.cugs_b200_scan_temp_bytes(n),), dtype=torch.uint8, device=dev)
    total = C.c_int64(0)
    _lib.check(h, lib.cugs_b200_scan(h, s, n, pr.tiles_touched.data_ptr(), offsets.data_ptr(), None, C.byref(total),
                                     tmp.data_ptr(), tmp.numel()), "scan")
    P = int(total.value)
    keys0 = torch.empty((P,), dtype=torch.int64, device=dev)
    vals0 = torch.empty((P,), dtype=torch.int32, device=dev)
    _lib.check(h, lib.cugs_b200_duplicate_with_keys(h, s, n, W, H, pr.means_2d.data_ptr(), pr.depths.data_ptr(),
                                                    pr.radii.data_ptr(), pr.tiles_touched.data_ptr(),
                                                    offsets.data_ptr(), P, keys0.data_ptr(), vals0.data_ptr()), "dup")
    keys_in, vals_in = keys0.clone(), vals0.clone()
    keys_out, vals_out = torch.empty_like(keys0), torch.empty_like(vals0)

    def restore():
        keys_in.copy_(keys0)
        vals_in.copy_(vals0)

    cub = load_cub()
    res = {"workload": desc, "N": n, "P": P, "tiles": T, "key_bits": 32 + tile_bits, "arms": {}}

    def arm(name, passes, fn, need_restore, check=True):
        med, mn = timed(fn, restore if need_restore else None)
        nbytes = (8 + 24 * passes) * P
        res["arms"][name] = {"ms_median": round(med, 4), "ms_min": round(mn, 4), "passes": passes,
                             "formulation_bytes": nbytes, "gbs": round(nbytes / (med * 1e-3) / 1e9, 1),
                             "mpairs_per_s": round(P / (med * 1e-3) / 1e6, 1)}
        return med

    # --- CUB, all 64 bits (the reference call) ---
    tb = cub.cub_sort_pairs_temp_bytes(P, 0, 64)
    ctemp = torch.empty((tb,), dtype=torch.uint8, device=dev)

    def cub64():
        st = cub.cub_sort_pairs(ctemp.data_ptr(), tb, keys_in.data_ptr(), keys_out.data_ptr(), vals_in.data_ptr(),
                                vals_out.data_ptr(), P, 0, 64, s)
        assert st == 0
    restore(); cub64(); torch.cuda.synchronize()
    ref_keys, ref_vals = keys_out.clone(), vals_out.clone()
    arm("cub64_as_called", 8, cub64, False)

    # --- CUB trimmed to the significant bits ---
    end_bit = 32 + tile_bits

    def cubt():
        st = cub.cub_sort_pairs(ctemp.data_ptr(), tb, keys_in.data_ptr(), keys_out.data_ptr(), vals_in.data_ptr(),
                                vals_out.data_ptr(), P, 0, end_bit, s)
        assert st == 0
    restore(); cubt(); torch.cuda.synchronize()
    assert torch.equal(keys_out, ref_keys) and torch.equal(vals_out, ref_vals)
    arm(f"cub_end_bit_{end_bit}", math.ceil(end_bit / 8), cubt, False)

    # --- this repository's 64-bit stage sort ---
    stmp = torch.empty((lib.cugs_b200_sort_temp_bytes(P),), dtype=torch.uint8, device=dev)

    def stage():
        _lib.check(h, lib.cugs_b200_sort_pairs(h, s, P, 32, tile_bits, keys_in.data_ptr(), vals_in.data_ptr(),
                                               keys_out.data_ptr(), vals_out.data_ptr(), stmp.data_ptr(), stmp.numel()),
                   "sort_pairs")
    restore(); stage(); torch.cuda.synchronize()
    assert torch.equal(keys_out, ref_keys) and torch.equal(vals_out, ref_vals), "stage sort order differs from CUB"
    arm("repo_stage_sort_64bit_keys", math.ceil(end_bit / 8), stage, True)
    del keys_in, vals_in, keys_out, stmp, ctemp
    torch.cuda.empty_cache()

    # --- the fused path's sort stage (inside render), by the library's stage events ---
    buf = cugs.FrameBuffers(n, W, H, 16, dev)
    settings = cugs.RenderSettings((0.0, 0.0, 0.0), 3, 1.0)
    out = cugs.render(model, scene.camera, settings, buf)
    assert torch.equal(out.gaussian_indices, ref_vals), "fused sort order differs from CUB"
    lib.cugs_b200_set_stage_timing(h, 1)
    ms8 = (C.c_float * 8)()
    fs, binning = [], []
    for it in range(25):
        cugs.render(model, scene.camera, settings, buf)
        lib.cugs_b200_get_stage_ms(h, ms8)
        if it >= 5:
            fs.append(float(ms8[3]))
            binning.append(float(ms8[1]) + float(ms8[2]) + float(ms8[3]) + max(float(ms8[4]), 0.0))
    lib.cugs_b200_set_stage_timing(h, 0)
    med = statistics.median(fs)
    passes = math.ceil(end_bit / 8)
    res["arms"]["repo_fused_depth_sort_plus_tile_sort"] = {
        "ms_median": round(med, 4), "ms_min": round(min(fs), 4), "passes": passes,
        "formulation_bytes": (8 + 24 * passes) * P,
        "gbs_equivalent": round((8 + 24 * passes) * P / (med * 1e-3) / 1e9, 1),
        "mpairs_per_s": round(P / (med * 1e-3) / 1e6, 1),
        "note": "gbs_equivalent divides the REFERENCE formulation's bytes by this design's time (the design moves "
                "8-byte packed elements and sorts the depth bits on N, so it is not a bandwidth)",
        "scan_dup_sort_ranges_ms": round(statistics.median(binning), 4)}
    a = res["arms"]
    res["speedup_vs_cub64"] = {k: round(a["cub64_as_called"]["ms_median"] / v["ms_median"], 3) for k, v in a.items()}
    res["speedup_vs_cub_trimmed"] = {k: round(a[f"cub_end_bit_{end_bit}"]["ms_median"] / v["ms_median"], 3)
                                     for k, v in a.items()}
    return res


if __name__ == "__main__":
    wl = sys.argv[1:] or ["B", "E"]
    for w in wl:
        print(json.dumps(run(w)), flush=True)
        torch.cuda.empty_cache()
