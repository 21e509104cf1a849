"""Times the phases of the sparse gradient exchange against the dense all-reduce (N GPUs, torchrun):

    python -m torch.distributed.run --nproc-per-node N tools/exchange_probe.py [touched fraction]

Every phase alone (CUDA events, 10 iterations after 3 warm-ups, max over ranks is what bounds a step), then the
whole exchange in its two modes: blocking on M (round 1) and with the row count read on the device (`state`).
Synthetic arena of 3 M Gaussians with a random touch mask of the given density per rank."""
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import cuda_gaussian_splatting_b200 as cugs  # noqa: E402
from cuda_gaussian_splatting_b200.parallel import _CudaRowOps  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 3_000_000
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.16
b = cugs.FrameBuffers(n, 64, 64, 16, dev)
# the ranks' views overlap heavily: a common core plus a small private part, like neighbouring cameras
g_common = torch.Generator(device=dev).manual_seed(1)
g_own = torch.Generator(device=dev).manual_seed(100 + rank)
common = torch.rand(n, device=dev, generator=g_common) < frac * 0.85
own = torch.rand(n, device=dev, generator=g_own) < frac * 0.15
local_mask = (common | own).int()
b.grad_arena.normal_()
ops = _CudaRowOps()


def reset_mask():
    b.touch_mask.copy_(local_mask)


def run(fn, iters=10, pre=None):
    ts = []
    for it in range(iters + 3):
        if pre:
            pre()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        if it >= 3:
            ts.append((e0.elapsed_time(e1), wall))
    g = torch.tensor([sum(t[0] for t in ts) / len(ts), sum(t[1] for t in ts) / len(ts)], device=dev, dtype=torch.float64)
    dist.all_reduce(g, op=dist.ReduceOp.MAX)
    return round(float(g[0]), 4), round(float(g[1]), 4)


res = {}
res["dense all-reduce(sum) of the 61N arena"] = run(lambda: dist.all_reduce(b.grad_arena))
res["MAX all-reduce of [mask | max_radii] (2N i32)"] = run(lambda: dist.all_reduce(b.max_buf, op=dist.ReduceOp.MAX), pre=reset_mask)
reset_mask()
dist.all_reduce(b.max_buf, op=dist.ReduceOp.MAX)
res["scan + host read of M"] = run(lambda: ops.scan(b))
res["scan, M stays on the device"] = run(lambda: ops.scan_dev(b))
off, m = ops.scan(b)
cap = int(m * 1.1) + 1024
_, m_dev = ops.scan_dev(b)
rows = ops.compact_floats(m, 16)
rows_cap = ops.compact_floats(cap, 16)
compact = torch.zeros((rows_cap + 2 * n,), device=dev)
res["gather (exact M)"] = run(lambda: ops.gather(b, off, m, compact[:rows]))
res["gather (capacity 1.1 M, device count)"] = run(lambda: ops.gather(b, off, cap, compact[:rows_cap], m_dev))
res["all-reduce(sum) of the M rows"] = run(lambda: dist.all_reduce(compact[:rows]))
res["all-reduce(sum) of 1.1 M rows + 2N statistics"] = run(lambda: dist.all_reduce(compact[:rows_cap + 2 * n]))
res["scatter (exact M)"] = run(lambda: ops.scatter(b, off, m, compact[:rows]))
res["whole exchange, blocking on M, with statistics"] = run(lambda: cugs.sparse_allreduce_step(b, with_stats=True), pre=reset_mask)
state = {}
reset_mask()
cugs.sparse_allreduce_step(b, with_stats=True, state=state)
res["whole exchange, M on the device, with statistics"] = run(
    lambda: cugs.sparse_allreduce_step(b, with_stats=True, state=state), pre=reset_mask)
res["whole exchange, M on the device, no statistics"] = run(
    lambda: cugs.sparse_allreduce_step(b, with_stats=False, state=state), pre=reset_mask)
if rank == 0:
    print(json.dumps({"world": world, "n": n, "touched_union": m, "touched_fraction": round(m / n, 4),
                      "row_capacity": state.get("m_cap"),
                      "phases_ms_gpu_wall": res}, indent=1))
dist.barrier()
dist.destroy_process_group()
