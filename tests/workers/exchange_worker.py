"""Worker of tests/test_multi_gpu_exchange.py (run under `python -m torch.distributed.run`, one rank per GPU).

Checks, on real NCCL over NVLink:
  1. sparse exchange == dense all-reduce == single-rank accumulation of ALL views (every rank recomputes the
     whole batch locally as the reference value): untouched rows exactly 0, touched rows <= 1e-6 norm-rel;
  2. the densification statistics travel with the exchange (grad_accum / grad_count summed, max_radii maxed);
  3. after Adam + one MCMC noise step the parameters are bit-identical on every rank.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import cuda_gaussian_splatting_b200 as cugs  # noqa: E402
from cuda_gaussian_splatting_b200 import parallel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, W, H, V = 60_000, 640, 360, 2 * world + 1           # odd view count: ranks hold different numbers of views
    scene = cugs.synth(n, W, H, seed=77)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    mk = lambda: cugs.GaussianModel(t(scene.positions), t(scene.sh_coeffs), t(scene.opacities), t(scene.rotations),
                                    t(scene.scales))
    model = mk()
    cams = [scene.camera] + cugs.ring_cameras(scene, V - 1, radius_frac=0.05)
    rng = np.random.default_rng(5)
    dLs = [t(rng.uniform(-1, 1, size=(H, W, 3)).astype(np.float32)) for _ in range(V)]
    settings = cugs.RenderSettings((0.1, 0.1, 0.1), 3, 1.0)

    def accumulate(views, buf, sparse, with_stats):
        stats = (buf.step_grad_accum, buf.step_grad_count, buf.step_max_radii) if with_stats else None
        if with_stats:
            buf.step_grad_accum.zero_(); buf.step_grad_count.zero_(); buf.step_max_radii.zero_()
        for k, v in enumerate(views):
            out = cugs.render(model, cams[v], settings, buf)
            cugs.render_backward(dLs[v], out, model, cams[v], settings, buf, stats=stats, accumulate=(k > 0),
                                 touch_mask=buf.touch_mask if sparse else None, sparse_rows=sparse)

    mine = parallel.shard_views(V, world, rank)
    # (a) single-rank reference: every rank accumulates ALL views densely
    ref_buf = cugs.FrameBuffers(n, W, H, 16, dev)
    accumulate(list(range(V)), ref_buf, sparse=False, with_stats=True)
    ref_arena = ref_buf.grad_arena.clone()
    ref_maxr = ref_buf.step_max_radii.clone()
    # (b) dense all-reduce of the local shard
    dense = cugs.FrameBuffers(n, W, H, 16, dev)
    accumulate(mine, dense, sparse=False, with_stats=True)
    parallel.allreduce_step(dense.grad_arena, dense.step_max_radii)
    # (c) sparse exchange of the local shard (two steps in a row: the mask / zero-row invariant must survive)
    # the first exchange blocks once on M and sets the row capacity; the next two read M on the device
    sp = cugs.FrameBuffers(n, W, H, 16, dev)
    infos, state = [], {}
    for _ in range(3):
        accumulate(mine, sp, sparse=True, with_stats=True)
        # (dense_threshold 0.99: this small scene is mostly visible; the sparse path is what is under test)
        infos.append(parallel.sparse_allreduce_step(sp, with_stats=True, state=state, dense_threshold=0.99))
    torch.cuda.synchronize()

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    errs = []
    layout, _ = parallel.arena_layout(n, 16)
    for nm, (off, sz) in layout.items():
        r_, d_, s_ = ref_arena[off:off + sz], dense.grad_arena[off:off + sz], sp.grad_arena[off:off + sz]
        if rel(d_, r_) > 1e-6:
            errs.append(f"dense vs single-rank {nm}: {rel(d_, r_):.3e}")
        if rel(s_, r_) > 1e-6:
            errs.append(f"sparse vs single-rank {nm}: {rel(s_, r_):.3e}")
        if rel(s_, d_) > 1e-6:
            errs.append(f"sparse vs dense {nm}: {rel(s_, d_):.3e}")
    untouched = sp.touch_mask == 0
    rows = torch.cat([sp.dL_dpositions, sp.dL_dsh_coeffs.reshape(n, -1), sp.dL_dopacities, sp.dL_dscales,
                      sp.dL_drotations], dim=1)
    if float(rows[untouched].abs().max() if bool(untouched.any()) else 0.0) != 0.0:
        errs.append("untouched rows are not exactly zero after the sparse exchange")
    ref_rows = torch.cat([ref_buf.dL_dpositions, ref_buf.dL_dsh_coeffs.reshape(n, -1), ref_buf.dL_dopacities,
                          ref_buf.dL_dscales, ref_buf.dL_drotations], dim=1)
    if bool(untouched.any()) and float(ref_rows[untouched].abs().max()) != 0.0:
        errs.append("a row outside the union mask has a non-zero single-rank gradient")
    if not torch.equal(sp.step_max_radii, ref_maxr) or not torch.equal(dense.step_max_radii, ref_maxr):
        errs.append("max_radii differs from the single-rank maximum")
    if not torch.equal(sp.step_grad_count, ref_buf.step_grad_count):
        errs.append("grad_count differs")
    if infos[0].get("mode") != "sparse":
        errs.append(f"expected the sparse mode, got {infos[0]}")
    if infos[0].get("host_sync") is not True or infos[2].get("host_sync") is not False:
        errs.append(f"expected a blocking first exchange and a non-blocking third one, got {infos}")
    if state.get("overflow"):
        errs.append("the row capacity overflowed")

    # (c2) the C++ step driver's views phase, split at the last view's mask, with the mask collective hidden under the
    # rest of that backward (parallel.MaskOverlap): same sums as the plain schedule and as the single-rank reference
    tcfg = cugs.TrainConfig(densify=True, background=tuple(settings.background))
    nat_bufs = cugs.FrameBuffers(n, W, H, 16, dev)
    nat = cugs.NativeTrainer(model, [cams[v] for v in mine], [None] * len(mine), tcfg, total_views_per_step=V,
                             grad_buffers=nat_bufs, dL_dcolors=[dLs[v] for v in mine], use_graph=True)
    ov, st2 = parallel.MaskOverlap(dev), {}
    for it in range(4):   # eager, capture, replay, replay
        nat.step_views_until_mask(3000 + it)
        ov.start(nat_bufs, st2)
        nat.step_views_rest(3000 + it)
        ov.finish()
        parallel.sparse_allreduce_step(nat_bufs, with_stats=True, state=st2, dense_threshold=0.99, mask_reduced=True)
    torch.cuda.synchronize()
    _, ok, _, _ = nat.result()
    if not ok or st2.get("overflow"):
        errs.append("native trainer / overlapped exchange overflowed")
    for nm, (off, sz) in layout.items():
        if rel(nat_bufs.grad_arena[off:off + sz], ref_arena[off:off + sz]) > 1e-6:
            errs.append(f"overlapped native schedule vs single-rank {nm}: {rel(nat_bufs.grad_arena[off:off + sz], ref_arena[off:off + sz]):.3e}")
    if not torch.equal(nat_bufs.step_max_radii, ref_maxr) or not torch.equal(nat_bufs.step_grad_count, ref_buf.step_grad_count):
        errs.append("overlapped native schedule: statistics differ")
    nat.close()

    # (c3) the peer-to-peer exchange over symmetric memory (no NCCL, no compact buffers): same sums, identical bits
    # on every rank
    pbuf = cugs.FrameBuffers(n, W, H, 16, dev, symmetric=True)
    modes = []
    for use_mc in (False, True):   # unicast peer loads / stores, then the NVSwitch multicast mapping (if there is one)
        p2p = parallel.P2PExchange(pbuf, use_multicast=use_mc)
        if use_mc and not p2p.multicast:
            continue
        modes.append("multicast" if p2p.multicast else "unicast")
        tag = modes[-1]
        for _ in range(3):   # consecutive steps: the mask / zero-row invariant must survive the exchange
            accumulate(mine, pbuf, sparse=True, with_stats=True)
            p2p.exchange(with_stats=True)
        torch.cuda.synchronize()
        for nm, (off, sz) in layout.items():
            if rel(pbuf.grad_arena[off:off + sz], ref_arena[off:off + sz]) > 1e-6:
                errs.append(f"p2p {tag} vs single-rank {nm}: {rel(pbuf.grad_arena[off:off + sz], ref_arena[off:off + sz]):.3e}")
        if not torch.equal(pbuf.step_max_radii, ref_maxr) or not torch.equal(pbuf.step_grad_count, ref_buf.step_grad_count):
            errs.append(f"p2p {tag}: statistics differ")
        if not torch.equal(pbuf.touch_mask, sp.touch_mask):
            errs.append(f"p2p {tag}: union mask differs from the NCCL path")
        root = pbuf.grad_arena.clone()
        dist.broadcast(root, 0)
        if not torch.equal(root, pbuf.grad_arena):
            errs.append(f"p2p {tag}: the arenas of the ranks are not bit-identical")
    if rank == 0:
        print(f"p2p exchange modes tested: {modes}", flush=True)

    # (d) replicas stay identical: Adam (grad_scale = 1/V) + MCMC noise on the exchanged gradients
    opt = cugs.FusedAdam(model)
    opt.grad_scale = 1.0 / V
    opt.apply_gradients(cugs.BackwardOutput(sp.dL_dpositions, sp.dL_drotations, sp.dL_dscales, sp.dL_dopacities,
                                            sp.dL_dsh_coeffs, sp.dL_dmeans_2d))
    opt.step()
    cugs.mcmc_inject_noise(model, 600, cugs.MCMCConfig())
    for nm in ("positions", "sh_coeffs", "opacities", "rotations", "scales"):
        mine_t = getattr(model, nm).contiguous()
        root = mine_t.clone()
        dist.broadcast(root, 0)
        if not torch.equal(root, mine_t):
            errs.append(f"{nm} differs from rank 0 after Adam + noise")
        if not bool(torch.isfinite(mine_t).all()):
            errs.append(f"{nm} not finite")
    flag = torch.tensor([len(errs)], device=dev)
    dist.all_reduce(flag)
    for e in errs:
        print(f"[rank {rank}] FAIL {e}", flush=True)
    if rank == 0:
        print(f"exchange worker: world {world}, views {V}, union touched {infos[0].get('touched')} of {n}, "
              f"total failures {int(flag.item())}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
