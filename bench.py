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
    """ALGORITHMIC bytes per frame of each HBM-bound stage = SURVEY.md §8(d)'s per-unit figure x the units
    of one frame (each distinct input read once + each output written once, for the REFERENCE formulation
    of the stage). `design_bytes` (below) is what this design actually has to move; it is reported beside
    it, never used for `achieved`."""
    tiles = ((w + 15) // 16) * ((h + 15) // 16)
    return {
        "preprocess_fwd": 284 * n,                   # 236 in (12+16+12+4+192) + 48 out
        "scan": 8 * n,                               # tiles_touched in, offsets out
        "duplicate_with_keys": 20 * n + 12 * p,      # mean, radius, tiles, offset per Gaussian + 12-B pair out
        "sort": (8 + 24 * (4 + tile_passes)) * p,    # 1 histogram read of the keys + passes x (read 12 + write 12)
        "tile_ranges": 8 * p + 8 * tiles,
        "preprocess_bwd": 336 * n,                   # 44 params + 4 radii + 36 incoming + 12 rgb in, 236 out
        "loss": 36 * w * h,                          # rendered + target in, dL/dcolor out
        "adam": 28 * 59 * n,                         # p, g, m, v in; p, m, v out
    }


def design_bytes(n: int, p: int, w: int, h: int, tile_passes: int, views: int = 1, touched=None) -> dict:
    """Bytes the kernels of THIS design move per frame: the depth passes are hoisted in
    front of duplicateWithKeys (8-byte packed elements), preprocess also writes the 48-byte blend record and
    the 8-byte depth-sort element, preprocess_bwd skips the untouched gradient rows."""
    tiles = ((w + 15) // 16) * ((h + 15) // 16)
    return {
        "preprocess_fwd": (284 + 48 + 8) * n,
        "scan": 16 * n,
        "duplicate_with_keys": 28 * n + 8 * p,
        "sort": (8 + 4 * 16) * n + (8 + 16 * (tile_passes - 1) + 12) * p + 12 * tiles,
        "tile_ranges": 0,
        "preprocess_bwd": int(((336 + 236 * (views - 1) / max(views, 1)) if touched is None else
                               (120 + 236 * touched * (2 * views - 1) / max(views, 1))) * n),
        "loss": (36 + 72) * w * h,
        "adam": 28 * 59 * n,
    }


# SURVEY.md §8(d) work units of the two blend kernels: FLOP (FMA = 2) and MUFU operations per
# (pixel, Gaussian) evaluation of the REFERENCE traversal (counted exactly by cugs_b200_count_evaluations)
FWD_FLOP = {"rejected": 15, "contributing": 24}
BWD_FLOP = {"rejected": 15, "contributing": 55}
FWD_MUFU = {"rejected": 1, "contributing": 1}
BWD_MUFU = {"rejected": 1, "contributing": 2}


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


MIN_TIMED_MS = 1000.0  # the timed region lasts at least this long whatever --steps says (clock sampling, noise)


def calibrated_steps(requested: int, est_ms_per_step: float) -> int:
    import math
    return max(int(requested), int(math.ceil(MIN_TIMED_MS * 1.05 / max(est_ms_per_step, 1e-3))))


def exchange_buffers(cugs, torch, dist, n, W, H, coeffs, dev, want_p2p):
    """Gradient buffers of one rank + the peer-to-peer exchange over them. The P2P kernels need torch's symmetric
    memory (NVLink peer mappings of every rank's arena); where a box cannot give it, ALL ranks fall back together to
    the NCCL row exchange on ordinary device memory (still this library's gather / scatter kernels around NCCL) and
    the bench line says so in config.gradient_exchange."""
    if not want_p2p:
        return cugs.FrameBuffers(n, W, H, coeffs, dev), None, None
    buf = p2p = why = None
    try:
        buf = cugs.FrameBuffers(n, W, H, coeffs, dev, symmetric=True)
        p2p = cugs.P2PExchange(buf)
    except Exception as e:  # noqa: BLE001 -- any failure of the symmetric allocation / rendezvous
        why = f"{type(e).__name__}: {str(e)[:160]}"
    ok = torch.tensor([0 if why else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 1:
        return buf, p2p, None
    del p2p, buf
    return cugs.FrameBuffers(n, W, H, coeffs, dev), None, "p2p exchange unavailable on this box (" + (
        why or "another rank failed") + "): NCCL row exchange"


def run_b200(args) -> dict:
    import ctypes as C

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
    strong = args.views_total > 0
    if strong:
        assert args.views_total % world == 0, "--views-total must be a multiple of the number of GPUs"
    V = args.views_total // world if strong else args.views_per_gpu
    scene = cugs.synth(n, W, H, seed=seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    model = cugs.GaussianModel(t(scene.positions), t(scene.sh_coeffs), t(scene.opacities), t(scene.rotations),
                               t(scene.scales))
    all_cams = bench_views(scene, world * V)
    cams = [all_cams[i] for i in cugs.shard_views(world * V, world, rank)]
    settings = cugs.RenderSettings((0.0, 0.0, 0.0), 3, 1.0)
    lib, h = _lib.load_library(), _lib.handle(local)

    if args.mode == "train_step":
        return run_b200_train_step(args, cugs, torch, dist, scene, model, cams, rank, world, local, dev, desc)
    # N > 1: the gradient arena lives in symmetric memory so that the exchange can run peer to peer over NVLink
    use_p2p = world > 1 and args.exchange == "p2p" and not args.dense_allreduce
    buf, p2p, p2p_note = exchange_buffers(cugs, torch, dist, n, W, H, 16, dev, use_p2p)

    # synthetic targets (host, pinned) and the resident dL/dcolor of each view; the blocking renders also
    # establish the pair capacity of every buffer used below
    rng = np.random.default_rng(4321 + rank)
    targets_host = [torch.from_numpy(rng.uniform(size=(H, W, 3)).astype(np.float32)).pin_memory() for _ in range(V)]
    target_dev = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    dLs, Ps = [], []
    work = None
    for v in range(V):
        out = cugs.render(model, cams[v], settings, buf)
        Ps.append(int(out.gaussian_indices.numel()))
        if v == 0 and rank == 0:
            work = cugs.count_evaluations(out, cams[0])   # E_fwd / E_bwd of view 0 (measurement only)
        target_dev.copy_(targets_host[v], non_blocking=True)
        _, g = cugs.combined_loss_with_grad(out.color, target_dev, 0.2)
        dLs.append(g)
    torch.cuda.synchronize()
    p_cap = int(max(Ps) * 1.25) + 4096

    with_stats = world > 1   # north_star: "parameter gradients plus densification statistics" in the exchange
    exchange = {"mode": "single"}
    if p2p_note:
        exchange["note"] = p2p_note
    exch_state = {}

    def allreduce(mask_reduced=False):
        # the step's gradient exchange: sparse (MAX-reduce of the touch mask + ONE sum all-reduce of the
        # touched rows and the additive statistics) unless --dense-allreduce (ONE sum all-reduce of the whole
        # 61N-float arena + the MAX-reduce of max_radii)
        if world == 1:
            return
        if args.dense_allreduce:
            cugs.allreduce_step(buf.grad_arena, buf.step_max_radii if with_stats else None)
            exchange.update(mode="dense")
        elif p2p is not None:
            exchange.update(p2p.exchange(with_stats=with_stats))
        else:
            exchange.update(cugs.sparse_allreduce_step(buf, with_stats=with_stats, state=exch_state,
                                                       mask_reduced=mask_reduced))
    # touch mask + sparse gradient rows (rows a view does not touch are neither read nor written) unless
    # the dense exchange is requested (it sums rows this rank's mask does not know about)
    sparse = not args.dense_allreduce
    touch = buf.touch_mask if sparse else None

    # ---- headline path: the C++ step driver's "views" phase (cugs_b200_trainer_step, phases = 1) with a given
    # dL/dcolor per view = forward + backward only. No host round trip per view (device-side pair count),
    # two frames in flight on two streams, the whole step replayed as ONE CUDA graph.
    tcfg = cugs.TrainConfig(densify=with_stats)
    nat = cugs.NativeTrainer(model, cams, [None] * V, tcfg, total_views_per_step=world * V, pair_capacity=p_cap,
                             frames_in_flight=1 if args.no_overlap else 2, use_graph=not args.no_graph,
                             sparse_rows=sparse, grad_buffers=buf, dL_dcolors=dLs)
    step_no = [3000]
    # the MAX all-reduce of the touch mask (and the scan of the union) run on a side stream under the chain rule
    # of the step's last backward
    overlap = cugs.MaskOverlap(dev) if (world > 1 and sparse and p2p is None and not args.no_mask_overlap) else None

    def step_resident():
        if overlap is not None:
            nat.step_views_until_mask(step_no[0])
            overlap.start(buf, exch_state)
            nat.step_views_rest(step_no[0])
            overlap.finish()
            allreduce(mask_reduced=True)
        else:
            nat.step_views(step_no[0])
            allreduce()
        step_no[0] += 1

    # ---- end-to-end path: the same C++ step driver through its public Python face, but with the loss in the step
    # and HOST inputs: every view's target image is copied from pinned host memory inside the step (a copy stream,
    # under the rendering of that view), the fused L1+SSIM loss produces dL/dcolor, and the step's {loss, l1, ssim,
    # ok, largest P} block is copied device -> host at its end; the host then waits for it.
    targets_dev = [torch.empty((H, W, 3), dtype=torch.float32, device=dev) for _ in range(V)]
    nat_e2e = cugs.NativeTrainer(model, cams, targets_dev, tcfg, total_views_per_step=world * V, pair_capacity=p_cap,
                                 frames_in_flight=1 if args.no_overlap else 2, use_graph=not args.no_graph,
                                 sparse_rows=sparse, grad_buffers=buf)
    nat_e2e.set_views(cams, targets_dev, None, targets_host)

    def step_e2e():
        if overlap is not None:
            nat_e2e.step_views_until_mask(step_no[0])
            overlap.start(buf, exch_state)
            nat_e2e.step_views_rest(step_no[0])
            overlap.finish()
            allreduce(mask_reduced=True)
        else:
            nat_e2e.step_views(step_no[0])
            allreduce()
        step_no[0] += 1
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

    def agree_steps(k):  # every rank must run the same number of steps (collectives inside)
        tk = torch.tensor([k], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        return int(tk.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_resident()
    steps = agree_steps(calibrated_steps(args.steps, timed(step_resident, 5) / 5))
    l0 = int(lib.cugs_b200_launch_count(h))
    sampler.mark_begin()
    total_ms = timed(step_resident, steps)
    sampler.mark_end()
    launches = int(lib.cugs_b200_launch_count(h)) - l0
    _, ok, pmax, _ = nat.result()
    assert ok, f"a frame overflowed the pair capacity ({pmax} > {p_cap}): the timed region is invalid"

    # ---- per-stage device timing (events recorded by the library on the launching stream), one frame in flight
    stage_sum = [0.0] * len(STAGES)
    stage_cnt = 0
    lib.cugs_b200_set_stage_timing(h, 1)
    ms8 = (C.c_float * 8)()
    for _ in range(min(steps, 30)):
        for v in range(V):
            out = cugs.render(model, cams[v], settings, buf, sync=False)
            cugs.render_backward(dLs[v], out, model, cams[v], settings, buf, accumulate=(v > 0), touch_mask=touch,
                                 sparse_rows=sparse)
            lib.cugs_b200_get_stage_ms(h, ms8)
            for k in range(8):
                stage_sum[k] += max(float(ms8[k]), 0.0)
            stage_cnt += 1
    lib.cugs_b200_set_stage_timing(h, 0)
    stages_ms = {nm: stage_sum[k] / max(stage_cnt, 1) for k, nm in enumerate(STAGES)}

    # ---- the per-step kernels either side of the rasterizer, timed alone with an L2
    # flush between iterations (their inputs at 1080p fit in the 126 MB L2)
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)

    def timed_alone(fn, iters=10):
        ts = []
        for _ in range(iters + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts[2:])

    extra_ms = {}
    if rank == 0:
        out0 = cugs.render(model, cams[0], settings, buf)
        target_dev.copy_(targets_host[0])
        extra_ms["loss"] = timed_alone(lambda: cugs.combined_loss_with_grad(out0.color, target_dev, 0.2))
        scratch = cugs.GaussianModel(*[x.clone() for x in (model.positions, model.sh_coeffs, model.opacities,
                                                           model.rotations, model.scales)])
        opt = cugs.FusedAdam(scratch)
        opt.apply_gradients(cugs.BackwardOutput(buf.dL_dpositions, buf.dL_drotations, buf.dL_dscales, buf.dL_dopacities,
                                                buf.dL_dsh_coeffs, buf.dL_dmeans_2d))
        extra_ms["adam"] = timed_alone(opt.step)
        del scratch, opt
    exchange_ms = None
    if world > 1:   # the exchange alone, on the gradients of a finished step (every rank, max over ranks)
        step_resident()
        exchange_ms = timed(allreduce, 10) / 10
    del flush

    for _ in range(3):
        step_e2e()
    e2e_steps = agree_steps(calibrated_steps(args.steps, timed(step_e2e, 5) / 5))
    e2e_ms = timed(step_e2e, e2e_steps)
    e2e_scalars, e2e_ok, _, _ = nat_e2e.result()
    assert e2e_ok, "an end-to-end frame overflowed its pair capacity"
    clocks = sampler.stop() if rank == 0 else {}

    views = world * V * steps
    value = views / (total_ms * 1e-3)
    e2e_value = world * V * e2e_steps / (e2e_ms * 1e-3)

    res = None
    if rank == 0:
        P = Ps[0]
        c_passes, c_bits = C.c_int(0), C.c_int(0)
        lib.cugs_b200_last_sort_plan(h, C.byref(c_passes), C.byref(c_bits))
        passes, key_bits = int(c_passes.value), int(c_bits.value)
        touched_frac = float((buf.touch_mask != 0).float().mean()) if touch is not None else None
        tile_passes = max(passes - 4, 1)
        alg = algorithmic_bytes(n, P, W, H, tile_passes, V, touched_frac)
        des = design_bytes(n, P, W, H, tile_passes, V, touched_frac)
        peak, peak_src = measured_peaks()
        sms, khz, l2 = C.c_int(0), C.c_int(0), C.c_int(0)
        lib.cugs_b200_device_info(h, C.byref(sms), C.byref(khz), C.byref(l2))
        clock_ghz = khz.value / 1e6
        fp32_peak = sms.value * 128 * 2 * clock_ghz / 1e3      # TFLOP/s: SMs x 128 lanes x 2 (FMA) x clock
        mufu_peak = sms.value * 16 * clock_ghz / 1e3           # T op/s: SMs x 16 SFU lanes x clock
        all_ms = dict(stages_ms)
        all_ms.update(extra_ms)
        rl_all = {}
        for nm in ["preprocess_fwd", "scan", "duplicate_with_keys", "sort", "tile_ranges", "preprocess_bwd", "loss", "adam"]:
            ms = all_ms.get(nm, 0.0)
            gbs = alg[nm] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            rl_all[nm] = {"ms": round(ms, 4), "bound": "hbm", "algorithmic_bytes": alg[nm], "design_bytes": des[nm],
                          "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4),
                          "design_gbs": round(des[nm] / (ms * 1e-3) / 1e9, 1) if ms > 0 else 0.0}
        rl_all["tile_ranges"]["note"] = "no separate pass: the ranges come from the tile histogram inside the sort stage"
        rl_all["sort"]["note"] = ("algorithmic bytes = the reference formulation (12-B pairs, 6 passes): the design sorts the "
                                  "depth bits on the N Gaussians and only the tile bits on the pairs, so achieved_gbs can "
                                  "exceed the HBM peak; design_gbs is the traffic this design really has")
        rl_all["loss"]["note"] = "k_ssim_moments + k_ssim_gradient + k_loss_finalize, timed alone with an L2 flush in between"
        rl_all["adam"]["note"] = "k_adam_multi over 59 N floats, timed alone with an L2 flush in between"
        for nm, fl, mu, e_r, e_c in (("blend_fwd", FWD_FLOP, FWD_MUFU, "fwd_rejected", "fwd_contributing"),
                                     ("blend_bwd", BWD_FLOP, BWD_MUFU, "bwd_rejected", "bwd_contributing")):
            ms = stages_ms[nm]
            flop = fl["rejected"] * work[e_r] + fl["contributing"] * work[e_c]
            mufu = mu["rejected"] * work[e_r] + mu["contributing"] * work[e_c]
            tf = flop / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            tm = mufu / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            rl_all[nm] = {"ms": round(ms, 4), "bound": "fp32-issue", "evaluations_rejected": work[e_r],
                          "evaluations_contributing": work[e_c], "algorithmic_flop": flop, "algorithmic_mufu": mufu,
                          "achieved_tflops": round(tf, 3), "frac_of_fp32_peak": round(tf / fp32_peak, 4),
                          "achieved_tmufu": round(tm, 4), "frac_of_mufu_peak": round(tm / mufu_peak, 4),
                          "hbm_bytes": (40 * P + 20 * W * H) + (36 * P if nm == "blend_bwd" else 0)}
        if exchange_ms is not None:
            m_union = int((buf.touch_mask != 0).sum())
            ex_floats = exchange.get("floats") or ((59 * m_union + 2 * n) if p2p is not None else buf.grad_arena.numel())
            rl_all["exchange"] = {"ms": round(exchange_ms, 4), "bound": "nvlink", "bytes_summed": 4 * int(ex_floats),
                                  "touched_union": m_union,
                                  "bytes_max_reduced": 8 * n, "mode": exchange.get("mode"),
                                  "note": "whole exchange of one step, timed alone after a finished step (max over ranks)"}
        # `roofline`: the DOMINANT kernel of the step, k_blend_bwd, against the FP32-issue roofline of SURVEY 8(d)
        bwd = rl_all["blend_bwd"]
        roofline = {"kernel": "k_blend_bwd", "bound": "fp32-issue", "achieved": bwd["achieved_tflops"],
                    "peak": round(fp32_peak, 2), "unit": "TFLOP/s", "frac": bwd["frac_of_fp32_peak"],
                    "traffic": TRAFFIC.get("k_blend_bwd"), "ms": bwd["ms"], "share_of_step": round(bwd["ms"] / max(sum(stages_ms.values()), 1e-9), 3),
                    "work": "E_bwd of view 0 by the reference traversal (backward.cu:117-145): 15 FLOP + 1 EX2 per "
                            "alpha-rejected, 55 FLOP + 1 EX2 + 1 RCP per contributing evaluation (SURVEY 8d)",
                    "mufu": {"achieved": bwd["achieved_tmufu"], "peak": round(mufu_peak, 3), "unit": "T op/s",
                             "frac": bwd["frac_of_mufu_peak"]},
                    "peak_source": f"{sms.value} SMs x 128 FP32 lanes x 2 x {clock_ghz:.3f} GHz (cudaDeviceGetAttribute)",
                    "hbm_peak_gbs": peak, "hbm_peak_source": peak_src}
        ref_sort_gbs = reference_sort_bytes(P, W, H) / (stages_ms["sort"] * 1e-3) / 1e9 if stages_ms["sort"] > 0 else 0.0
        res = {
            "metric": METRIC, "value": round(value, 3), "unit": "views/s", "n_gpus": world, "steps": steps,
            "steps_requested": args.steps, "warmup": warm, "ms_per_step": round(total_ms / steps, 4),
            "ms_per_view": round(total_ms / steps / V, 4), "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "views_per_gpu_per_step": V, "views_per_step": world * V,
                       "frames_in_flight": 1 if args.no_overlap else 2,
                       "driver": "cugs_b200_trainer_step(phases = views): C++ step driver, device-side pair count, "
                                 + ("one CUDA graph per step" if not args.no_graph else "eager launches"),
                       "P_pairs_view0": P, "pair_capacity": p_cap, "sort_passes": passes, "sort_key_bits": key_bits,
                       "densification_stats_fused": with_stats,
                       "collective": "none" if world == 1 else (
                           "one NCCL all-reduce(sum) of the 61N-float arena per step (+ MAX of max_radii)" if args.dense_allreduce else
                           ("per step, peer to peer over NVLink on symmetric memory (no NCCL): k_xchg_masks (MAX of [touch mask | "
                            "max_radii], SUM of the statistics) + k_xchg_rows (each rank sums its slice of the touched gradient rows "
                            "across all arenas and writes it back to all), 3 device-side barriers") if p2p is not None else
                           "per step: int32 MAX all-reduce of [touch mask | max_radii] (8 B/Gaussian) + ONE NCCL all-reduce(sum) of "
                           "the touched gradient rows and the additive statistics"),
                       "gradient_exchange": exchange, "touched_fraction": touched_frac,
                       "mask_allreduce_overlapped": overlap is not None,
                       "min_timed_ms": MIN_TIMED_MS,
                       "l2": "no flush needed: per-step inputs (708 MB of Gaussian parameters at 3M) exceed the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 3), "unit": "views/s", "ms_per_view": round(e2e_ms / e2e_steps / V, 4),
                    "steps": e2e_steps,
                    "h2d_bytes_per_step": V * H * W * 3 * 4, "d2h_bytes_per_step": 48,
                    "loss": round(e2e_scalars[0], 6),
                    "what": "NativeTrainer.step_views (cugs_b200_trainer_step), per view: H2D target from pinned host memory "
                            "(copy stream, under the rendering of the same view) -> render -> fused L1+SSIM loss+grad -> "
                            "render_backward; per step: D2H {loss, l1, ssim, ok, largest P} + host wait"},
            "gpu_launches": launches,
            "roofline": roofline,
            "work_view0": work,
            "stages_ms": {k: round(v_, 4) for k, v_ in stages_ms.items()},
            "stages_sum_ms": round(sum(stages_ms.values()), 4),
            "roofline_all": rl_all,
            "sort_gbs": round(ref_sort_gbs, 1),
            "sort_gbs_note": "bytes of the reference formulation (12-B pairs, 6 onesweep passes at 1080p = 152 B/pair) "
                             "divided by the time of this library's whole sort stage (depth sort of N + tile sort of P): a "
                             "cross-implementation throughput, NOT a bandwidth; the traffic this design has is "
                             "roofline_all.sort.design_gbs; CUB on the same pairs: profiles/r02/sort_bench*.jsonl",
            "sort_mpairs_per_s": round(P / (stages_ms["sort"] * 1e-3) / 1e6, 1) if stages_ms["sort"] > 0 else None,
        }
        if world == 1 and not args.no_cpu_baseline:
            res["cpu_baseline"] = cpu_baseline(scene, args.workload)
    nat.close()
    nat_e2e.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return res


def run_b200_train_step(args, cugs, torch, dist, scene, model, cams, rank, world, local, dev, desc):
    """BASELINE.json configs[2] / [3]: the full training step (render fwd + fused L1/SSIM loss + render bwd
    with fused densification statistics + gradient exchange + ONE fused Adam launch) through the C++ step
    driver (cugs_b200_trainer_step; Trainer::train_step, training/trainer.cpp:178-316), replayed as CUDA graphs."""
    import numpy as np
    from cuda_gaussian_splatting_b200 import _lib
    n, W, H = scene.n, scene.camera.width, scene.camera.height
    V = len(cams)
    rng = np.random.default_rng(4321 + rank)
    targets = [torch.from_numpy(rng.uniform(size=(H, W, 3)).astype(np.float32)).to(dev) for _ in range(V)]
    use_p2p = world > 1 and args.exchange == "p2p"
    gbuf, p2p, p2p_note = exchange_buffers(cugs, torch, dist, n, W, H, int(model.sh_coeffs.shape[2]), dev, use_p2p)
    trainer = cugs.NativeTrainer(model, cams, targets, cugs.TrainConfig(), total_views_per_step=world * V,
                                 frames_in_flight=1 if args.no_overlap else 2, use_graph=not args.no_graph,
                                 grad_buffers=gbuf)
    lib, h = _lib.load_library(), _lib.handle(local)
    step_no = [3000]  # SH degree 3 active (lr_schedule.hpp:70-72)
    exch_state = {}
    overlap = cugs.MaskOverlap(dev) if (world > 1 and p2p is None and not args.no_mask_overlap) else None

    def step():
        if world == 1:
            trainer.train_step(step_no[0])
        else:
            if overlap is not None:
                trainer.step_views_until_mask(step_no[0])
                overlap.start(trainer.buffers, exch_state)
                trainer.step_views_rest(step_no[0])
                overlap.finish()
            else:
                trainer.step_views(step_no[0])
            if p2p is not None:
                p2p.exchange(with_stats=True)
            else:
                cugs.sparse_allreduce_step(trainer.buffers, with_stats=True, state=exch_state,
                                           mask_reduced=overlap is not None)
            b = trainer.buffers
            cugs.fold_step_stats(b.step_grad_accum, b.step_grad_count, b.step_max_radii, trainer.stats.grad_accum,
                                 trainer.stats.grad_count, trainer.stats.max_radii_2d)
            trainer.step_update(step_no[0])
        step_no[0] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    k = torch.tensor([calibrated_steps(args.steps, timed(5) / 5)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(k, op=dist.ReduceOp.MAX)
    steps = int(k.item())
    l0 = int(lib.cugs_b200_launch_count(h))
    sampler.mark_begin()
    total_ms = timed(steps)
    sampler.mark_end()
    launches = int(lib.cugs_b200_launch_count(h)) - l0
    clocks = sampler.stop() if rank == 0 else {}
    scalars, ok, pmax, _ = trainer.result()
    assert ok, f"a frame overflowed the pair capacity ({pmax} > {trainer.pair_capacity})"
    trainer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None
    views = world * V * steps
    return {"metric": "training views/s (render fwd+bwd + L1/SSIM loss + densification stats + fused Adam)",
            "value": round(views / (total_ms * 1e-3), 3), "unit": "views/s", "n_gpus": world, "steps": steps,
            "steps_requested": args.steps, "warmup": warm, "ms_per_step": round(total_ms / steps, 4),
            "ms_per_view": round(total_ms / steps / V, 4), "higher_is_better": True,
            "scaling": "strong" if args.views_total > 0 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc.replace("fwd+bwd", "full training step"), "views_per_gpu_per_step": V,
                       "views_per_step": world * V, "adam_elements": 59 * n, "l2": "inputs exceed L2",
                       "driver": "cugs_b200_trainer_step: C++ step driver, no host synchronisation, "
                                 + ("CUDA graph replay" if not args.no_graph else "eager launches"),
                       "pair_capacity": trainer.pair_capacity, "largest_pair_count": pmax,
                       "gradient_exchange": ("single" if world == 1 else
                                             (p2p_note or ("p2p" + ("+multicast" if p2p.multicast else "")
                                                           if p2p is not None else "nccl rows")))},
            "clocks": clocks, "gpu_launches": launches, "final_loss": scalars[0]}


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
def run_reference(args, module: str = "cugs_ref") -> dict:
    """module = "cugs_ref": the UNMODIFIED reference compiled for sm_100 (the reference arm);
    module = "cugs_dropin" (--impl dropin): the SAME harness and the same calls -- cugs::render / render_backward /
    combined_loss with the reference's own C++ signatures -- linked against wrapper/cugs_b200_dropin.cpp, i.e.
    this library behind the reference's API exactly as the reference's C++ training loop would use it."""
    rank, world, local = dist_env()
    if rank != 0:
        return None
    import numpy as np
    import torch

    import cuda_gaussian_splatting_b200 as cugs  # synth() only: the scene generator, no kernels
    n, W, H, seed, desc = WORKLOADS[args.workload]
    scene = cugs.synth(n, W, H, seed=seed)
    impl = "reference" if module == "cugs_ref" else "dropin"
    base = {"impl": impl, "metric": METRIC, "unit": "views/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "views_per_gpu_per_step": args.views_per_gpu}}
    ref = None
    why = ""
    try:
        sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
        import importlib
        ref = importlib.import_module(module)  # cugs_ref: the unmodified reference, compiled from /root/reference for sm_100
        assert torch.cuda.is_available()
    except Exception as e:  # noqa: BLE001
        ref, why = None, f"{type(e).__name__}: {e}"
    if ref is None and impl == "dropin":
        base["unavailable"] = f"oracle/_ref/cugs_dropin*.so not built ({why})"
        return base
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
    steps = calibrated_steps(args.steps, timed(step_resident, 3) / 3)   # >= 1 s timed region
    sampler.mark_begin()
    ms = timed(step_resident, steps)
    sampler.mark_end()
    clocks = sampler.stop()
    for _ in range(2):
        step_e2e()
    e2e_steps = calibrated_steps(args.steps, timed(step_e2e, 3) / 3)
    e2e_ms = timed(step_e2e, e2e_steps)
    value = steps * V / (ms * 1e-3)
    base["steps"], base["steps_requested"] = steps, args.steps
    base.update({
        "value": round(value, 3), "ms_per_step": round(ms / steps, 4), "ms_per_view": round(ms / steps / V, 4),
        "clocks": clocks,
        "cpu_baseline": {"value": round(value, 3), "unit": "views/s", "cores": 0, "kind": "reference",
                         "sample": ("the unmodified reference's own CUDA render()+render_backward() (oracle/_ref, built "
                                    "for sm_100), full workload, run on the GPU: the reference has no CPU rasterizer")
                         if impl == "reference" else
                         "this library behind the reference's C++ API (wrapper/cugs_b200_dropin.cpp), same harness calls"},
        "e2e": {"value": round(e2e_steps * V / (e2e_ms * 1e-3), 3), "unit": "views/s",
                "ms_per_view": round(e2e_ms / e2e_steps / V, 4), "steps": e2e_steps,
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
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "dropin"])
    ap.add_argument("--workload", default="B", choices=sorted(WORKLOADS))
    ap.add_argument("--views-per-gpu", type=int, default=2,
                    help="views rendered fwd+bwd per GPU per step (2 = BASELINE config[3]: 16 views/step on 8 GPUs)")
    ap.add_argument("--views-total", type=int, default=0,
                    help="STRONG scaling: a fixed number of views per step split over the GPUs (16 = BASELINE config[3])")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA graph replay")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: gradient exchange as peer-to-peer kernels over NVLink symmetric memory (default) or as NCCL "
                         "all-reduces around gather / scatter copies")
    ap.add_argument("--no-mask-overlap", action="store_true",
                    help="N > 1: run the MAX all-reduce of the touch mask after the backward instead of under it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dense-allreduce", action="store_true",
                    help="N > 1: all-reduce the whole gradient arena instead of only the touched rows")
    ap.add_argument("--no-overlap", action="store_true", help="one frame in flight (no 2-stream view pipelining)")
    ap.add_argument("--mode", default="fwd_bwd", choices=["fwd_bwd", "train_step"],
                    help="fwd_bwd = the headline metric; train_step = BASELINE config[2] (full step incl. loss, Adam, stats)")
    args = ap.parse_args()
    if args.impl == "reference":
        res = run_reference(args)
    elif args.impl == "dropin":
        res = run_reference(args, module="cugs_dropin")
    else:
        res = run_b200(args)
    if res is not None:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
