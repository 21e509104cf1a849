#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native rasterizer hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload B]

Metric (BASELINE.json): forward+backward throughput of the differentiable rasterizer at 3 M
synthetic Gaussians / SH degree 3 / 1920x1080 (workload "B" = BASELINE.json configs[1]), as
views/s (whole job) with ms/view beside it, plus the sort's GB/s.

One "step" = `--views-per-gpu` views (default 2) rendered forward + backward on every GPU,
parameter gradients summed in place; with N > 1 GPUs the step ends with ONE NCCL all-reduce(sum)
of the gradient arena (59 N floats + 2 N densification statistics), as BASELINE.json config[3].
Per-GPU work is fixed as N grows ("scaling": "weak").

  value     : device-timed (CUDA events, max over ranks); everything resident in HBM.
  e2e       : the same step through the public API with HOST per-view inputs: every view the
              target image is copied host->device from pinned memory, the fused L1+SSIM loss
              produces dL/dcolor, and the three loss scalars are read back device->host.
              Gaussian parameters are model state (as in the reference's Trainer, which uploads
              them once: training/trainer.cpp:83) and stay resident.
  roofline  : the dominant kernel of the step, per-stage CUDA events recorded by the library on
              the launching stream (cugs_b200_set_stage_timing) over a second pass of K steps.
  cpu_baseline : the CPU oracle port (oracle/cugs_oracle.c, OpenMP, all host cores) on ONE full
              view of the same workload (rank 0, N = 1 only).

`--impl reference` times the UNMODIFIED reference (oracle/_ref/cugs_ref*.so = its own CUDA
kernels compiled for sm_100 from /root/reference by oracle/Makefile.ref) through its public
render()/render_backward() API on the same scene; if that module cannot be loaded it falls back
to the CPU oracle port. Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (N, W, H, seed, description)
    "A": (100_000, 1280, 720, 1235, "A: 100k Gaussians SH3 1280x720 fwd+bwd"),
    "B": (3_000_000, 1920, 1080, 1236, "B: 3M Gaussians SH3 1920x1080 fwd+bwd"),
    "C": (1_000_000, 1920, 1080, 1237, "C: 1M Gaussians SH3 1920x1080 fwd+bwd"),
    "E": (20_000_000, 3840, 2160, 1239, "E: 20M Gaussians SH3 3840x2160 fwd+bwd"),
}
METRIC = "fwd+bwd views/s at 3M Gaussians 1080p"
STAGES = ["preprocess_fwd", "scan", "duplicate_with_keys", "sort", "tile_ranges", "blend_fwd", "blend_bwd",
          "preprocess_bwd"]


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls nvidia-smi (profiling recipe's clocks line) from before the warm-up until after the timed
    region; `stop()` summarises the samples whose timestamp falls INSIDE the timed window
    (mark_begin / mark_end), falling back to all samples under load when the window is too short
    to contain one."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.thread = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def summarise(rows):
            sm, mx, pw, reasons = [], [], [], set()
            for _, ln in rows:
                parts = [p.strip() for p in ln.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                    pw.append(float(parts[3]))
                except ValueError:
                    continue
                for nm, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            return sm, mx, pw, reasons

        window = "timed region"
        rows = [r for r in self.lines if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1 + 0.02]
        sm, mx, pw, reasons = summarise(rows)
        if len(sm) < 2:  # region shorter than the sampling period: use every sample taken under load
            window = "warm-up + timed region + stage pass (timed region shorter than the sampling period)"
            sm, mx, pw, reasons = summarise(self.lines)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def measured_peaks() -> tuple:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def sort_passes(P_bits_depth: int, num_tiles: int) -> int:
    import math
    tb = max(0, math.ceil(math.log2(num_tiles))) if num_tiles > 1 else 0
    return max(1, (P_bits_depth + tb + 7) // 8)


def algorithmic_bytes(n: int, p: int, w: int, h: int, tile_passes: int, views: int = 1, touched=None) -> dict:
    """ALGORITHMIC bytes per frame of each stage: each distinct input read once +
    each output written once, for the kernels as designed (depth passes hoisted before the key
    duplication: the sort moves 8-byte packed elements)."""
    tiles = ((w + 15) // 16) * ((h + 15) // 16)
    return {
        "preprocess_fwd": (284 + 48 + 8) * n,        # reference outputs + packed blend record + depth-sort element
        "scan": 16 * n,                              # sorted element (8) + gathered tile count (4) + offset (4)
        "duplicate_with_keys": 28 * n + 8 * p,       # element, offset, radius, tiles, mean (8+4+4+4+8) + packed pair
        "sort": (8 + 4 * 16) * n + (8 + 16 * (tile_passes - 1) + 12) * p + 12 * tiles,
        "tile_ranges": 0,                            # part of the sort (tile histogram -> ranges)
        "blend_fwd": 52 * p + 20 * w * h,            # index (4) + packed record (48) per pair, 20 B per pixel
        "blend_bwd": 52 * p + 20 * w * h + 48 * n,   # + dL/dcolor, final_T, n_contrib per pixel, 48-B gradient record
        # dense: 336 B/Gaussian, views after the first of a step also READ the 236-B gradient row they
        # add to. Sparse rows (touched = fraction of Gaussians with a gradient): 120 B/Gaussian of inputs,
        # mask and dL/dmeans_2d, plus the 236-B row written (first view) or read+written (later views)
        # for the touched fraction only.
        "preprocess_bwd": int(((336 + 236 * (views - 1) / max(views, 1)) if touched is None else
                               (120 + 236 * touched * (2 * views - 1) / max(views, 1))) * n),
    }


def reference_sort_bytes(p: int, w: int, h: int) -> int:
    """The sort as the reference formulates it: 12-byte (u64 key, u32 value) pairs,
    one histogram read + ceil((32 + tile_bits) / 8) onesweep passes."""
    tiles = ((w + 15) // 16) * ((h + 15) // 16)
    passes = (32 + max(0, (tiles - 1).bit_length()) + 7) // 8
    return (8 + 24 * passes) * p


# ------------------------------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def bench_views(scene, count: int):
    """View 0 = the scene's camera 0 (BASELINE config B's camera); further views = small jitters
    around it (ring of 2 % of the median depth), so every view carries the same work."""
    import cuda_gaussian_splatting_b200 as cugs
    cams = [scene.camera]
    if count > 1:
        cams += cugs.ring_cameras(scene, count - 1, radius_frac=0.02)
    return cams


def run_b200(args) -> dict:
    import numpy as np
    import torch
    import torch.distributed as dist

    import cuda_gaussian_splatting_b200 as cugs
    from cuda_gaussian_splatting_b200 import _lib

    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a B200 (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, W, H, seed, desc = WORKLOADS[args.workload]
    V = args.views_per_gpu
    scene = cugs.synth(n, W, H, seed=seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    model = cugs.GaussianModel(t(scene.positions), t(scene.sh_coeffs), t(scene.opacities), t(scene.rotations),
                               t(scene.scales))
    all_cams = bench_views(scene, world * V)
    cams = [all_cams[i] for i in cugs.shard_views(world * V, world, rank)]
    settings = cugs.RenderSettings((0.0, 0.0, 0.0), 3, 1.0)
    buf = cugs.FrameBuffers(n, W, H, 16, dev)
    lib, h = _lib.load_library(), _lib.handle(local)

    if args.mode == "train_step":
        return run_b200_train_step(args, cugs, torch, dist, scene, model, cams, rank, world, local, dev, desc)

    # synthetic targets (host, pinned) and the resident dL/dcolor of each view
    rng = np.random.default_rng(4321 + rank)
    targets_host = [torch.from_numpy(rng.uniform(size=(H, W, 3)).astype(np.float32)).pin_memory() for _ in range(V)]
    target_dev = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    scal_host = torch.empty((3,), dtype=torch.float32).pin_memory()
    dLs, Ps = [], []
    for v in range(V):
        out = cugs.render(model, cams[v], settings, buf)
        Ps.append(int(out.gaussian_indices.numel()))
        target_dev.copy_(targets_host[v], non_blocking=True)
        _, g = cugs.combined_loss_with_grad(out.color, target_dev, 0.2)
        dLs.append(g)
    torch.cuda.synchronize()

    exchange = {"mode": "single"}

    def allreduce():
        # the step's gradient exchange: sparse (MAX-reduce of the touch mask + ONE sum all-reduce of the
        # touched rows) unless --dense-allreduce (ONE sum all-reduce of the whole 61N-float arena)
        if world == 1:
            return
        if args.dense_allreduce:
            cugs.allreduce_step(buf.grad_arena)
            exchange.update(mode="dense")
        else:
            exchange.update(cugs.sparse_allreduce_step(buf, with_stats=False))
    # touch mask + sparse gradient rows (rows a view does not touch are neither read nor written) unless
    # the dense exchange is requested (it sums rows this rank's mask does not know about)
    touch = None if args.dense_allreduce else buf.touch_mask

    # Two frames in flight inside one step: view v runs on stream v % 2 with its own frame buffers, so
    # the preprocess / sort / forward blend of view v+1 overlap the (issue-bound) backward blend of view v.
    # The gradient arena is shared: view v's backward waits for view v-1's (in-place accumulation), and a
    # step starts only after the previous step has completely finished (no overlap across steps, where a
    # real training loop has its optimizer update).
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)] if V > 1 and not args.no_overlap else None
    bufs = [buf, cugs.FrameBuffers(n, W, H, 16, dev, share_grads_with=buf)] if streams else [buf, buf]

    def run_views(per_view):
        if streams is None:
            for v in range(V):
                per_view(v, bufs[0], None)
            return
        cur = torch.cuda.current_stream(dev)
        start = torch.cuda.Event()
        start.record(cur)
        prev_bwd = None
        for v in range(V):
            s = streams[v % 2]
            s.wait_event(start)
            with torch.cuda.stream(s):
                prev_bwd = per_view(v, bufs[v % 2], prev_bwd)
        cur.wait_event(prev_bwd)

    def view_resident(v, b, prev_bwd):
        out = cugs.render(model, cams[v], settings, b)
        if prev_bwd is not None:
            torch.cuda.current_stream(dev).wait_event(prev_bwd)
        cugs.render_backward(dLs[v], out, model, cams[v], settings, b, accumulate=(v > 0), touch_mask=touch,
                             sparse_rows=touch is not None)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        return ev

    def step_resident():
        run_views(view_resident)
        allreduce()

    uploader = cugs.TargetUploader(H, W, dev)
    uploader.prefetch(targets_host[0])

    scal_hosts = [torch.empty((3,), dtype=torch.float32).pin_memory() for _ in range(V)]

    def view_e2e(v, b, prev_bwd):
        # every view's target is copied host->device (pinned, side stream) INSIDE the step; the copy of
        # the next view overlaps the rendering of the current one
        tgt = uploader.get()
        uploader.prefetch(targets_host[(v + 1) % V])                          # H2D of the next view's target
        out = cugs.render(model, cams[v], settings, b)
        sc, g = cugs.combined_loss_with_grad(out.color, tgt, 0.2)
        uploader.release()
        if prev_bwd is not None:
            torch.cuda.current_stream(dev).wait_event(prev_bwd)
        cugs.render_backward(g, out, model, cams[v], settings, b, accumulate=(v > 0), touch_mask=touch,
                             sparse_rows=touch is not None)
        scal_hosts[v].copy_(sc, non_blocking=True)                            # D2H of {loss, l1, ssim}
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        return ev

    def step_e2e():
        run_views(view_e2e)
        allreduce()
        torch.cuda.current_stream().synchronize()                             # the losses are on the host now

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    l0 = int(lib.cugs_b200_launch_count(h))
    sampler.mark_begin()
    total_ms = timed(step_resident, args.steps)
    sampler.mark_end()
    launches = int(lib.cugs_b200_launch_count(h)) - l0

    # second pass: per-stage device timing (events recorded by the library on the launching stream)
    import ctypes as C
    stage_sum = [0.0] * len(STAGES)
    stage_cnt = 0
    lib.cugs_b200_set_stage_timing(h, 1)
    ms8 = (C.c_float * 8)()
    for _ in range(args.steps):
        for v in range(V):  # (one frame in flight here: the stage events are per handle)
            out = cugs.render(model, cams[v], settings, buf)
            cugs.render_backward(dLs[v], out, model, cams[v], settings, buf, accumulate=(v > 0), touch_mask=touch,
                                 sparse_rows=touch is not None)
            lib.cugs_b200_get_stage_ms(h, ms8)
            for k in range(8):
                stage_sum[k] += max(float(ms8[k]), 0.0)
            stage_cnt += 1
    lib.cugs_b200_set_stage_timing(h, 0)
    stages_ms = {nm: stage_sum[k] / max(stage_cnt, 1) for k, nm in enumerate(STAGES)}

    for _ in range(3):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else {}

    views = world * V * args.steps
    value = views / (total_ms * 1e-3)
    e2e_value = views / (e2e_ms * 1e-3)

    res = None
    if rank == 0:
        P = Ps[0]
        c_passes, c_bits = C.c_int(0), C.c_int(0)
        lib.cugs_b200_last_sort_plan(h, C.byref(c_passes), C.byref(c_bits))
        passes, key_bits = int(c_passes.value), int(c_bits.value)
        touched_frac = float(buf.touch_mask.float().mean()) if touch is not None else None
        alg = algorithmic_bytes(n, P, W, H, max(passes - 4, 1), V, touched_frac)
        peak, peak_src = measured_peaks()
        hbm_stages = ["preprocess_fwd", "scan", "duplicate_with_keys", "sort", "preprocess_bwd"]
        rl_all = {}
        for nm in STAGES:
            ms = stages_ms[nm]
            gbs = alg[nm] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            rl_all[nm] = {"ms": round(ms, 4), "algorithmic_bytes": alg[nm], "achieved_gbs": round(gbs, 1),
                          "frac_of_hbm_peak": round(gbs / peak, 4),
                          "bound": "hbm" if nm in hbm_stages else "fp32-issue/L2-atomics (not HBM)"}
        # `roofline`: the largest HBM-bound stage that is ONE kernel launch (the sort stage is a dozen
        # launches); the two blend kernels dominate the step but are instruction-issue bound, their
        # HBM fraction is not a quality measure — see roofline_all / issue_bound and DESIGN.md §4
        single = ["preprocess_fwd", "preprocess_bwd", "duplicate_with_keys", "scan"]
        dom = max(single, key=lambda k: stages_ms[k])
        kname = {"preprocess_fwd": "k_preprocess_fwd", "preprocess_bwd": "k_preprocess_bwd",
                 "duplicate_with_keys": "k_duplicate_sorted", "scan": "k_scan_exclusive"}[dom]
        roofline = {"kernel": kname, "bound": "hbm", "achieved": rl_all[dom]["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": rl_all[dom]["frac_of_hbm_peak"], "traffic": TRAFFIC.get(kname),
                    "peak_source": peak_src, "ms": rl_all[dom]["ms"],
                    "algorithmic_bytes_per_launch": alg[dom]}
        ref_sort_gbs = reference_sort_bytes(P, W, H) / (stages_ms["sort"] * 1e-3) / 1e9 if stages_ms["sort"] > 0 else 0.0
        res = {
            "metric": METRIC, "value": round(value, 3), "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / args.steps, 4),
            "ms_per_view": round(total_ms / args.steps / V, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "views_per_gpu_per_step": V, "frames_in_flight": 2 if streams else 1,
                       "P_pairs_view0": P, "sort_passes": passes,
                       "sort_key_bits": key_bits,
                       "collective": "none" if world == 1 else (
                           "one NCCL all-reduce(sum) of the 61N-float arena per step" if args.dense_allreduce else
                           "per step: int32 MAX all-reduce of the touch mask (8 B/Gaussian) + ONE all-reduce(sum) of the "
                           "touched gradient rows"),
                       "gradient_exchange": exchange, "touched_fraction": touched_frac,
                       "l2": "no flush needed: per-step inputs (708 MB of Gaussian parameters at 3M) exceed the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 3), "unit": "views/s", "ms_per_view": round(e2e_ms / args.steps / V, 4),
                    "h2d_bytes_per_step": V * H * W * 3 * 4, "d2h_bytes_per_step": V * 12,
                    "what": "per view: H2D target (pinned, side stream, overlapped with the previous view) -> render -> "
                            "fused L1+SSIM loss+grad -> render_backward -> D2H loss"},
            "gpu_launches": launches,
            "roofline": roofline,
            "stages_ms": {k: round(v_, 4) for k, v_ in stages_ms.items()},
            "stages_sum_ms": round(sum(stages_ms.values()), 4),
            "roofline_all": rl_all,
            "sort_gbs": round(ref_sort_gbs, 1),
            "sort_gbs_note": "bytes of the reference formulation (12-B pairs, 6 onesweep passes at 1080p = 152 B/pair) "
                             "divided by the time of this library's whole sort stage (depth sort of N + tile sort of P); "
                             "roofline_all.sort uses the bytes this design actually has to move",
            "sort_mpairs_per_s": round(P / (stages_ms["sort"] * 1e-3) / 1e6, 1) if stages_ms["sort"] > 0 else None,
        }
        if world == 1 and not args.no_cpu_baseline:
            res["cpu_baseline"] = cpu_baseline(scene, args.workload)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return res


def run_b200_train_step(args, cugs, torch, dist, scene, model, cams, rank, world, local, dev, desc):
    """BASELINE.json configs[2]: the full training step (render fwd + fused L1/SSIM loss + render bwd
    with fused densification statistics + one all-reduce + ONE fused Adam launch) through
    SyntheticTrainer (the Trainer::train_step skeleton, training/trainer.cpp:178-316)."""
    import numpy as np
    from cuda_gaussian_splatting_b200 import _lib
    n, W, H = scene.n, scene.camera.width, scene.camera.height
    V = args.views_per_gpu
    rng = np.random.default_rng(4321 + rank)
    targets = [torch.from_numpy(rng.uniform(size=(H, W, 3)).astype(np.float32)).to(dev) for _ in range(V)]
    trainer = cugs.SyntheticTrainer(model, cams, targets, cugs.TrainConfig(), total_views_per_step=world * V)
    lib, h = _lib.load_library(), _lib.handle(local)
    step_no = [3000]  # SH degree 3 active (lr_schedule.hpp:70-72)

    def step():
        trainer.train_step(step_no[0])
        step_no[0] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = int(lib.cugs_b200_launch_count(h))
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    sampler.mark_end()
    launches = int(lib.cugs_b200_launch_count(h)) - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else {}
    loss = float(trainer.last_scalars[0])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None
    views = world * V * args.steps
    return {"metric": "training views/s (render fwd+bwd + L1/SSIM loss + densification stats + fused Adam)",
            "value": round(views / (total_ms * 1e-3), 3), "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / args.steps, 4),
            "ms_per_view": round(total_ms / args.steps / V, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc.replace("fwd+bwd", "full training step"), "views_per_gpu_per_step": V,
                       "adam_elements": 59 * n, "l2": "inputs exceed L2"},
            "clocks": clocks, "gpu_launches": launches, "final_loss": loss}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
# captures (profiles/), filled in per round; None = not captured yet.
TRAFFIC = {}
try:
    TRAFFIC = json.loads((ROOT / "profiles" / "traffic.json").read_text())
except Exception:
    TRAFFIC = {}


def cpu_baseline(scene, workload: str, views: int = 1) -> dict:
    """The CPU oracle port (same math, OpenMP over all host cores) on `views` full views."""
    import numpy as np
    from oracle import oracle_py  # checker / baseline only
    H, W = scene.camera.height, scene.camera.width
    g = np.random.default_rng(1).uniform(-1, 1, size=(H, W, 3)).astype(np.float32)
    t0 = time.perf_counter()
    for _ in range(views):
        fwd = oracle_py.render_forward(scene, deg=3)
        oracle_py.render_backward(scene, fwd, g, deg=3)
    dt = time.perf_counter() - t0
    return {"value": round(views / dt, 5), "unit": "views/s", "ms_per_view": round(dt / views * 1e3, 1),
            "cores": oracle_py.num_threads(), "kind": "port",
            "sample": f"{views} full view(s) of workload {workload} fwd+bwd (oracle/cugs_oracle.c, OpenMP)"}


# ------------------------------------------------------------------------------------------------
def run_reference(args) -> dict:
    rank, world, local = dist_env()
    if rank != 0:
        return None
    import numpy as np
    import torch

    import cuda_gaussian_splatting_b200 as cugs  # synth() only: the scene generator, no kernels
    n, W, H, seed, desc = WORKLOADS[args.workload]
    scene = cugs.synth(n, W, H, seed=seed)
    base = {"impl": "reference", "metric": METRIC, "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "views_per_gpu_per_step": args.views_per_gpu}}
    ref = None
    why = ""
    try:
        sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
        import cugs_ref as ref  # the unmodified reference, compiled from /root/reference for sm_100
        assert torch.cuda.is_available()
    except Exception as e:  # noqa: BLE001
        ref, why = None, f"{type(e).__name__}: {e}"
    if ref is None:
        cb = cpu_baseline(scene, args.workload, views=max(1, min(args.steps, 2)))
        cb["sample"] += f"; compiled reference unavailable ({why})"
        base.update({"value": cb["value"], "ms_per_step": cb["ms_per_view"], "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return base

    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pos, sh, opa, rot, scl = t(scene.positions), t(scene.sh_coeffs), t(scene.opacities), t(scene.rotations), t(scene.scales)
    V = args.views_per_gpu
    cams, bg = [c.as_ref_list() for c in bench_views(scene, V)], [0.0, 0.0, 0.0]
    rng = np.random.default_rng(4321)
    targets_host = [torch.from_numpy(rng.uniform(size=(H, W, 3)).astype(np.float32)).pin_memory() for _ in range(V)]
    dLs = []
    for v in range(V):
        out = ref.render(pos, sh, opa, rot, scl, cams[v], bg, 3, 1.0)
        _, _, _, dL = ref.combined_loss_with_grad(out[0], targets_host[v].to(dev), 0.2)
        dLs.append(dL)

    def step_resident():  # the reference has no gradient accumulation: each view's gradients are separate tensors
        for v in range(V):
            o = ref.render(pos, sh, opa, rot, scl, cams[v], bg, 3, 1.0)
            ref.render_backward(dLs[v], o, pos, sh, opa, rot, scl, cams[v], bg, 3, 1.0)

    def step_e2e():
        for v in range(V):
            tg = targets_host[v].to(dev, non_blocking=True)
            o = ref.render(pos, sh, opa, rot, scl, cams[v], bg, 3, 1.0)
            loss, _, _, g = ref.combined_loss_with_grad(o[0], tg, 0.2)   # trainer.cpp:214-225
            ref.render_backward(g, o, pos, sh, opa, rot, scl, cams[v], bg, 3, 1.0)
            loss.item()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler.mark_begin()
    ms = timed(step_resident, args.steps)
    sampler.mark_end()
    clocks = sampler.stop()
    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    value = args.steps * V / (ms * 1e-3)
    base.update({
        "value": round(value, 3), "ms_per_step": round(ms / args.steps, 4), "ms_per_view": round(ms / args.steps / V, 4),
        "clocks": clocks,
        "cpu_baseline": {"value": round(value, 3), "unit": "views/s", "cores": 0, "kind": "reference",
                         "sample": "the unmodified reference's own CUDA render()+render_backward() (oracle/_ref, built "
                                   "for sm_100), full workload, run on the GPU: the reference has no CPU rasterizer"},
        "e2e": {"value": round(args.steps * V / (e2e_ms * 1e-3), 3), "unit": "views/s",
                "ms_per_view": round(e2e_ms / args.steps / V, 4),
                "h2d_bytes_per_step": V * H * W * 3 * 4, "d2h_bytes_per_step": V * 4,
                "what": "H2D target (pinned) -> ref render -> ref combined_loss + autograd -> ref render_backward -> loss.item()"},
        "P_pairs": int(out[9].numel()),
    })
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="B", choices=sorted(WORKLOADS))
    ap.add_argument("--views-per-gpu", type=int, default=2,
                    help="views rendered fwd+bwd per GPU per step (2 = BASELINE config[3]: 16 views/step on 8 GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dense-allreduce", action="store_true",
                    help="N > 1: all-reduce the whole gradient arena instead of only the touched rows")
    ap.add_argument("--no-overlap", action="store_true", help="one frame in flight (no 2-stream view pipelining)")
    ap.add_argument("--mode", default="fwd_bwd", choices=["fwd_bwd", "train_step"],
                    help="fwd_bwd = the headline metric; train_step = BASELINE config[2] (full step incl. loss, Adam, stats)")
    args = ap.parse_args()
    res = run_reference(args) if args.impl == "reference" else run_b200(args)
    if res is not None:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
