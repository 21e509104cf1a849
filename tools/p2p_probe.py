"""Times the pieces of the peer-to-peer gradient exchange (N GPUs, torchrun): the symmetric-memory barrier alone,
k_xchg_masks, scan + index list, k_xchg_rows, and the whole exchange; max over ranks, CUDA events."""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import cuda_gaussian_splatting_b200 as cugs  # noqa: E402
from cuda_gaussian_splatting_b200 import _lib  # noqa: E402
from cuda_gaussian_splatting_b200.rasterizer import _stream  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 3_000_000
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.16
b = cugs.FrameBuffers(n, 64, 64, 16, dev, symmetric=True)
g_common = torch.Generator(device=dev).manual_seed(1)
g_own = torch.Generator(device=dev).manual_seed(100 + rank)
local_mask = ((torch.rand(n, device=dev, generator=g_common) < frac * 0.85) |
              (torch.rand(n, device=dev, generator=g_own) < frac * 0.15)).int()
b.grad_arena.normal_()
b.grad_arena.mul_(1e-3)
x = cugs.P2PExchange(b, use_multicast=(os.environ.get("P2P_MULTICAST", "1") == "1"))
lib, h = _lib.load_library(), _lib.handle(local)


def reset():
    b.touch_mask.copy_(local_mask)


def run(fn, iters=10, pre=None):
    ts = []
    for it in range(iters + 3):
        if pre:
            pre()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1))
    g = torch.tensor([sum(ts) / len(ts)], device=dev, dtype=torch.float64)
    dist.all_reduce(g, op=dist.ReduceOp.MAX)
    return round(float(g[0]), 4)


res = {}
res["symmetric-memory barrier alone"] = run(lambda: x.h_max.barrier(channel=0))
res["3 barriers"] = run(lambda: (x.h_max.barrier(channel=0), x.h_max.barrier(channel=1), x.h_arena.barrier(channel=0)))
s = _stream(dev)
res["k_xchg_masks with statistics (no barrier)"] = run(
    lambda: lib.cugs_b200_p2p_reduce_masks(h, s, n, x.world, x.rank, x._maxbuf, x._accum, x._count,
                                           x._maxbuf_mc if x.multicast else None, x._accum_mc if x.multicast else None,
                                           x._count_mc if x.multicast else None), pre=reset)
reset()
x.h_max.barrier(channel=0)
lib.cugs_b200_p2p_reduce_masks(h, s, n, x.world, x.rank, x._maxbuf, None, None,
                               x._maxbuf_mc if x.multicast else None, None, None)
x.h_max.barrier(channel=1)
res["scan (M on the device) + index list"] = run(lambda: (x.ops.scan_dev(b), lib.cugs_b200_build_touch_index(
    h, s, n, b.touch_mask.data_ptr(), b.touch_offsets.data_ptr(), b._touch_idx.data_ptr())))
offsets, m_dev = x.ops.scan_dev(b)
m = int(m_dev.item())
res["k_xchg_rows (no barrier)"] = run(lambda: lib.cugs_b200_p2p_reduce_rows(h, s, n, 16, x.world, x.rank,
                                                                            b._touch_idx.data_ptr(), m_dev.data_ptr(), x._grads,
                                                                            x._grads_mc if x.multicast else None))
res["whole exchange with statistics"] = run(lambda: x.exchange(with_stats=True), pre=reset)
if rank == 0:
    rows_bytes = m * 59 * 4
    print(json.dumps({"world": world, "n": n, "multicast": x.multicast, "touched_union": m, "row_bytes_MB": round(rows_bytes / 1e6, 1),
                      "remote_MB_each_way_per_rank": round(rows_bytes * (world - 1) / world / 1e6, 1),
                      "phases_ms": res}, indent=1))
dist.barrier()
dist.destroy_process_group()
