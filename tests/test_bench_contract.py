"""bench.py's contract pieces that can be checked without a GPU: the workload table follows
BASELINE.json's configs, the stage list matches the library's stage-timing slots, the algorithmic byte
model (DESIGN.md 4 / SURVEY 8d) is self-consistent, and the committed bench lines carry every key the
driver reads."""
import json
import re
from pathlib import Path

import bench

ROOT = Path(__file__).resolve().parent.parent


def test_workloads_follow_baseline_configs():
    w = bench.WORKLOADS
    assert w["A"][:3] == (100_000, 1280, 720)          # configs[0]
    assert w["B"][:3] == (3_000_000, 1920, 1080)       # configs[1]: the metric's configuration
    assert w["C"][:3] == (1_000_000, 1920, 1080)       # configs[2]
    assert w["E"][:3] == (20_000_000, 3840, 2160)      # configs[4]
    metric = json.loads((ROOT / "BASELINE.json").read_text())["metric"]
    assert "3M Gaussians 1080p" in metric and "3M Gaussians 1080p" in bench.METRIC


def test_stage_list_matches_the_library():
    header = (ROOT / "include" / "cugs_b200.h").read_text()
    n = int(re.search(r"#define CUGS_NUM_STAGES (\d+)", header).group(1))
    assert len(bench.STAGES) == n == 8
    assert bench.STAGES[0] == "preprocess_fwd" and bench.STAGES[-1] == "preprocess_bwd"


def test_algorithmic_bytes_model():
    """`achieved` uses SURVEY 8(d)'s per-unit figures (the reference formulation of each stage); what this design
    really moves is reported separately (design_bytes) and never used for the roofline fraction."""
    n, p, w, h = 3_000_000, 18_596_764, 1920, 1080
    a = bench.algorithmic_bytes(n, p, w, h, tile_passes=2, views=2, touched=0.18)
    d = bench.design_bytes(n, p, w, h, tile_passes=2, views=2, touched=0.18)
    hbm = {"preprocess_fwd", "scan", "duplicate_with_keys", "sort", "tile_ranges", "preprocess_bwd", "loss", "adam"}
    assert set(a) == hbm == set(d)
    assert a["preprocess_fwd"] == 284 * n and d["preprocess_fwd"] == 340 * n   # + 48-B record + 8-B sort element
    assert a["scan"] == 8 * n and a["duplicate_with_keys"] == 20 * n + 12 * p
    assert a["sort"] == (8 + 24 * 6) * p == bench.reference_sort_bytes(p, w, h)  # 152 B/pair at 1080p
    assert a["preprocess_bwd"] == 336 * n and d["preprocess_bwd"] < a["preprocess_bwd"]  # sparse rows move fewer bytes
    assert a["adam"] == 28 * 59 * n and a["loss"] == 36 * w * h
    assert bench.reference_sort_bytes(p, 3840, 2160) == (8 + 24 * 6) * p    # 32 400 tiles -> 15 bits -> 47 key bits
    assert bench.reference_sort_bytes(p, 64, 64) == (8 + 24 * 5) * p        # 16 tiles -> 36 key bits
    # SURVEY 8(d) work units of the blend kernels
    assert bench.FWD_FLOP == {"rejected": 15, "contributing": 24} and bench.BWD_FLOP == {"rejected": 15, "contributing": 55}
    assert bench.BWD_MUFU["contributing"] == 2 and bench.calibrated_steps(20, 4.0) >= 250   # >= 1 s timed region


import pytest


@pytest.mark.parametrize("rnd", ["r01", "r02"])
def test_committed_bench_lines_carry_the_contract_keys(rnd):
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"}
    line = json.loads((ROOT / "profiles" / rnd / "bench_b200_final.json").read_text().strip().splitlines()[-1])
    assert need <= set(line), need - set(line)
    assert line["unit"] == "views/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["dtype"] == "f32" and line["vs_baseline"] is None and line["warmup"] >= 3
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] == 2 * 1920 * 1080 * 3 * 4 and line["gpu_launches"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    ref = json.loads((ROOT / "profiles" / rnd / "bench_reference_final.json").read_text().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["metric"] == line["metric"] and ref["unit"] == line["unit"]
    if rnd == "r02":   # round 2: work-based roofline of the dominant kernel, every stage in roofline_all, >= 1 s timed
        r = line["roofline"]
        assert r["kernel"] == "k_blend_bwd" and r["bound"] == "fp32-issue" and r["unit"] == "TFLOP/s"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and {"mufu", "work"} <= set(r)
        assert {"loss", "adam", "blend_fwd", "blend_bwd", "sort", "preprocess_fwd"} <= set(line["roofline_all"])
        w = line["work_view0"]
        assert w["E_bwd"] == w["bwd_rejected"] + w["bwd_contributing"] and w["E_fwd"] > 0
        assert line["ms_per_step"] * line["steps"] >= 1000.0
        drop = json.loads((ROOT / "profiles" / rnd / "bench_dropin_final.json").read_text().strip().splitlines()[-1])
        assert drop["impl"] == "dropin" and drop["unit"] == "views/s" and drop["value"] > ref["value"]


def test_exchange_buffers_fall_back_together_when_symmetric_memory_is_missing():
    """bench.exchange_buffers: a failing symmetric allocation / rendezvous turns into the NCCL row exchange on
    every rank (agreed through a MIN all-reduce), and the reason lands in the bench line."""
    import torch
    import torch.distributed as dist

    class FakeCugs:
        made = []

        @staticmethod
        def FrameBuffers(n, W, H, coeffs, dev, symmetric=False):
            if symmetric:
                raise RuntimeError("no fabric handles on this box")
            FakeCugs.made.append((n, W, H, coeffs))
            return "plain-buffers"

        @staticmethod
        def P2PExchange(buf):
            raise AssertionError("not reached")

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29641", rank=0, world_size=1)
    try:
        buf, p2p, note = bench.exchange_buffers(FakeCugs, torch, dist, 10, 64, 64, 16, "cpu", True)
        assert buf == "plain-buffers" and p2p is None and "no fabric handles" in note and "NCCL" in note
        buf, p2p, note = bench.exchange_buffers(FakeCugs, torch, dist, 10, 64, 64, 16, "cpu", False)
        assert buf == "plain-buffers" and p2p is None and note is None
    finally:
        dist.destroy_process_group()
