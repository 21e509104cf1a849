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
    n, p, w, h = 3_000_000, 18_596_764, 1920, 1080
    a = bench.algorithmic_bytes(n, p, w, h, tile_passes=2, views=2, touched=0.18)
    assert set(a) == set(bench.STAGES)
    assert a["preprocess_fwd"] == 340 * n                         # 284 reference bytes + 48 record + 8 sort element
    assert a["sort"] > 36 * p and a["duplicate_with_keys"] == 28 * n + 8 * p
    dense = bench.algorithmic_bytes(n, p, w, h, 2, views=1)["preprocess_bwd"]
    assert dense == 336 * n and a["preprocess_bwd"] < dense        # sparse rows move fewer bytes
    # the reference formulation of the sort: 12-byte pairs, one histogram read + 6 passes at 1080p
    assert bench.reference_sort_bytes(p, w, h) == (8 + 24 * 6) * p
    assert bench.reference_sort_bytes(p, 3840, 2160) == (8 + 24 * 6) * p    # 32 400 tiles -> 15 bits -> 47 key bits
    assert bench.reference_sort_bytes(p, 64, 64) == (8 + 24 * 5) * p        # 16 tiles -> 36 key bits


def test_committed_bench_lines_carry_the_contract_keys():
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"}
    line = json.loads((ROOT / "profiles" / "r01" / "bench_b200_final.json").read_text().strip().splitlines()[-1])
    assert need <= set(line), need - set(line)
    assert line["unit"] == "views/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["dtype"] == "f32" and line["vs_baseline"] is None and line["warmup"] >= 3
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] == 2 * 1920 * 1080 * 3 * 4 and line["gpu_launches"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    ref = json.loads((ROOT / "profiles" / "r01" / "bench_reference_final.json").read_text().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["metric"] == line["metric"] and ref["unit"] == line["unit"]
