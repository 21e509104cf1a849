"""Times an all-reduce(sum) of the 3M-Gaussian gradient arena (61 N floats = 732 MB) on N GPUs:
NCCL (whatever algorithm NCCL_ALGO selects) and, if available, torch symmetric-memory multimem."""
import os, sys, time
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 61 * 3_000_000
x = torch.randn(n, device=dev)

def timed(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

ms = timed(lambda: dist.all_reduce(x))
if rank == 0: print(f"NCCL all_reduce algo={os.environ.get('NCCL_ALGO','default')} {n*4/1e6:.0f} MB: {ms:.3f} ms  algbw {n*4/ms/1e6:.0f} GB/s", flush=True)
ms = timed(lambda: dist.reduce_scatter_tensor(x[: n // world], x[: n // world * world]))
if rank == 0: print(f"NCCL reduce_scatter: {ms:.3f} ms", flush=True)
if os.environ.get("PROBE_SYMM", "0") == "1":
    try:
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(n, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
        t.copy_(x)
        if rank == 0: print("symm_mem ok; multicast_ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, flush=True)
        for name in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
            if not hasattr(torch.ops.symm_mem, name): 
                if rank == 0: print("no op", name)
                continue
            op = getattr(torch.ops.symm_mem, name)
            try:
                ms = timed(lambda: op(t, "sum", dist.group.WORLD.group_name))
                if rank == 0: print(f"symm_mem.{name}: {ms:.3f} ms", flush=True)
            except Exception as e:
                if rank == 0: print(name, "failed:", str(e)[:200], flush=True)
    except Exception as e:
        if rank == 0: print("symm_mem unavailable:", str(e)[:300], flush=True)
dist.barrier(); dist.destroy_process_group()
