"""Times the phases of the sparse gradient exchange against the dense all-reduce (N GPUs, torchrun)."""
import os, sys, time
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import cuda_gaussian_splatting_b200 as cugs
from cuda_gaussian_splatting_b200.parallel import _CudaRowOps

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 3_000_000
b = cugs.FrameBuffers(n, 64, 64, 16, dev)
g = torch.Generator(device=dev).manual_seed(rank)
frac = float(os.environ.get("TOUCH", "0.15"))
b.touch_mask.copy_((torch.rand(n, device=dev, generator=g) < frac).int())
b.grad_arena.normal_()
ops = _CudaRowOps()

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def run(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); e0 = ev()
    for _ in range(iters): fn()
    e1 = ev(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (time.perf_counter() - t0) * 1e3 / iters

res = {}
res["dense allreduce"] = run(lambda: dist.all_reduce(b.grad_arena))
res["max allreduce (2N i32)"] = run(lambda: dist.all_reduce(b.max_buf, op=dist.ReduceOp.MAX))
res["scan + host sync"] = run(lambda: ops.scan(b))
off, m = ops.scan(b)
compact = torch.zeros((ops.compact_floats(m, 16),), device=dev)
res["gather"] = run(lambda: ops.gather(b, off, m, compact))
res["compact allreduce"] = run(lambda: dist.all_reduce(compact))
res["scatter"] = run(lambda: ops.scatter(b, off, m, compact))
res["sparse_allreduce_step"] = run(lambda: cugs.sparse_allreduce_step(b, with_stats=False))
if rank == 0:
    print(f"world {world} n {n} touched-union {m} ({m/n:.2%})")
    for k, (gpu, wall) in res.items(): print(f"  {k:28s} gpu {gpu:7.3f} ms   wall {wall:7.3f} ms")
dist.barrier(); dist.destroy_process_group()
