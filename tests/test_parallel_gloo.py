"""Host logic of view-parallel training on CPU: world_size-2 gloo processes shard a batch of views,
each builds a deterministic stand-in for its views' gradients in the arena layout, and ONE
all-reduce per step must give every rank the sum over all views (bit-identical on both ranks),
with max_radii max-reduced and the step statistics folded into the persistent accumulators."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

N, C, VIEWS = 1000, 16, 7


def fake_view_arena(view: int, total: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(1000 + view)
    return torch.randn(total, generator=g, dtype=torch.float32)


def fake_view_radii(view: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(2000 + view)
    return torch.randint(0, 50, (N,), generator=g).float()


def worker(rank: int, world: int, port: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cuda_gaussian_splatting_b200 import parallel
    layout, total = parallel.arena_layout(N, C)
    arena = torch.zeros(total)
    max_radii = torch.zeros(N)
    mine = parallel.shard_views(VIEWS, world, rank)
    for k, v in enumerate(mine):  # accumulate = (k > 0), as render_backward(accumulate=...) does
        arena += fake_view_arena(v, total) if k else 0
        if k == 0:
            arena.copy_(fake_view_arena(v, total))
        max_radii = torch.maximum(max_radii, fake_view_radii(v))
    parallel.allreduce_step(arena, max_radii)
    acc, cnt, mx = torch.zeros(N), torch.zeros(N), torch.zeros(N)
    o, s = layout["grad_accum"]
    o2, s2 = layout["grad_count"]
    parallel.fold_step_stats(arena[o:o + s], arena[o2:o2 + s2], max_radii, acc, cnt, mx)
    torch.save({"arena": arena, "max_radii": max_radii, "acc": acc, "mx": mx, "views": mine},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_views_partition():
    from cuda_gaussian_splatting_b200 import parallel
    for world in (1, 2, 4, 8):
        allv = sorted(v for r in range(world) for v in parallel.shard_views(16, world, r))
        assert allv == list(range(16))
        assert all(len(parallel.shard_views(16, world, r)) == 16 // world for r in range(world))
    assert parallel.shard_views(7, 2, 0) == [0, 2, 4, 6] and parallel.shard_views(7, 2, 1) == [1, 3, 5]
    with pytest.raises(ValueError):
        parallel.shard_views(4, 2, 2)


def test_arena_layout_matches_adam_group_order():
    from cuda_gaussian_splatting_b200 import parallel
    layout, total = parallel.arena_layout(1003, 16)
    names = list(layout)
    assert names[:5] == ["positions", "sh_coeffs", "opacities", "scales", "rotations"]  # fused_adam.cu:94-97
    assert [layout[n][1] for n in names] == [3009, 48144, 1003, 3009, 4012, 1003, 1003]
    assert all(layout[n][0] % 64 == 0 for n in names)  # 256-byte aligned segments (float4 accesses)
    ends = [layout[n][0] + layout[n][1] for n in names]
    assert all(ends[i] <= layout[names[i + 1]][0] for i in range(len(names) - 1)) and ends[-1] <= total


def test_allreduce_world_size_2_gloo(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from cuda_gaussian_splatting_b200 import parallel
    layout, total = parallel.arena_layout(N, C)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    assert sorted(r0["views"] + r1["views"]) == list(range(VIEWS))
    assert torch.equal(r0["arena"], r1["arena"]), "replicas must stay bit-identical after the all-reduce"
    # per-rank partial sums in the rank's own accumulation order, then the cross-rank sum
    partial = []
    for views in (r0["views"], r1["views"]):
        a = fake_view_arena(views[0], total).clone()
        for v in views[1:]:
            a += fake_view_arena(v, total)
        partial.append(a)
    assert torch.equal(r0["arena"], partial[0] + partial[1])
    expect_max = torch.stack([fake_view_radii(v) for v in range(VIEWS)]).max(dim=0).values
    assert torch.equal(r0["max_radii"], expect_max) and torch.equal(r1["max_radii"], expect_max)
    o, s = layout["grad_accum"]
    assert torch.equal(r0["acc"], r0["arena"][o:o + s]) and torch.equal(r0["mx"], expect_max)


# ---- sparse gradient exchange: host logic with torch stand-ins for the three CUDA row operations ----
def _a4(x):
    return (x + 3) // 4 * 4


class TorchRowOps:
    def compact_floats(self, m, num_coeffs):
        return _a4(3 * m) + _a4(3 * num_coeffs * m) + _a4(m) + _a4(3 * m) + _a4(4 * m)

    def scan(self, b):
        mask = b.touch_mask.to(torch.int64)
        inc = torch.cumsum(mask, 0)
        return (inc - mask).to(torch.int32), int(inc[-1])

    def _groups(self, b):
        return [b.dL_dpositions, b.dL_dsh_coeffs.view(b.n, -1), b.dL_dopacities, b.dL_dscales, b.dL_drotations]

    status_dev = None

    def scan_dev(self, b):  # the total stays "on the device": a tensor, never an int
        offsets, m = self.scan(b)
        return offsets, torch.tensor([m], dtype=torch.int64)

    def _rows(self, b, m, m_dev):
        """(selected rows, count): with m_dev, m is the row capacity (rows beyond it are dropped)."""
        idx = b.touch_mask.bool().nonzero().reshape(-1)
        if m_dev is not None:
            real = int(m_dev[0])
            if self.status_dev is not None:
                self.status_dev[0], self.status_dev[1] = real, int(real > m)
            idx = idx[:m]
        return idx, idx.numel()

    def gather(self, b, offsets, m, compact, m_dev=None):
        idx, k = self._rows(b, m, m_dev)
        off = 0
        for g in self._groups(b):
            w = g.shape[1]
            compact[off:off + k * w].copy_(g[idx].reshape(-1))
            compact[off + k * w:off + _a4(m * w)] = 0
            off += _a4(m * w)

    def scatter(self, b, offsets, m, compact, m_dev=None):
        idx, k = self._rows(b, m, m_dev)
        off = 0
        for g in self._groups(b):
            w = g.shape[1]
            g[idx] = compact[off:off + k * w].view(k, w)
            off += _a4(m * w)


class FakeBuffers:
    def __init__(self, n, C):
        from cuda_gaussian_splatting_b200 import parallel
        layout, total = parallel.arena_layout(n, C)
        self.n = n
        self.grad_arena = torch.zeros(total)
        seg = lambda nm: self.grad_arena[layout[nm][0]:layout[nm][0] + layout[nm][1]]
        self.dL_dpositions = seg("positions").view(n, 3)
        self.dL_dsh_coeffs = seg("sh_coeffs").view(n, 3, C)
        self.dL_dopacities = seg("opacities").view(n, 1)
        self.dL_dscales = seg("scales").view(n, 3)
        self.dL_drotations = seg("rotations").view(n, 4)
        self.step_grad_accum, self.step_grad_count = seg("grad_accum"), seg("grad_count")
        self.max_buf = torch.zeros(2 * n, dtype=torch.int32)
        self.touch_mask = self.max_buf[:n]
        self.step_max_radii = self.max_buf[n:].view(torch.float32)
        self.grad_compact = None
        self.touch_offsets = None


def fill_rank(b, rank):
    g = torch.Generator().manual_seed(77 + rank)
    touched = torch.rand(b.n, generator=g) < (0.2 if rank == 0 else 0.3)
    for t in (b.dL_dpositions, b.dL_dsh_coeffs, b.dL_dopacities, b.dL_dscales, b.dL_drotations):
        t.copy_(torch.randn(t.shape, generator=g) * touched.view(-1, *([1] * (t.dim() - 1))))
    b.touch_mask.copy_(touched.to(torch.int32))
    b.step_grad_accum.copy_(torch.rand(b.n, generator=g))
    b.step_grad_count.copy_((torch.rand(b.n, generator=g) < 0.5).float())
    b.step_max_radii.copy_(torch.randint(0, 40, (b.n,), generator=g).float())


def sparse_worker(rank, world, port, out_dir, threshold):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cuda_gaussian_splatting_b200 import parallel
    b = FakeBuffers(N, C)
    fill_rank(b, rank)
    info = parallel.sparse_allreduce_step(b, with_stats=True, dense_threshold=threshold, ops=TorchRowOps())
    torch.save({"arena": b.grad_arena.clone(), "max_buf": b.max_buf.clone(), "info": info}, os.path.join(out_dir, f"s{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("threshold,mode", [(0.6, "sparse"), (0.1, "dense")])
def test_sparse_exchange_equals_dense_sum(tmp_path, threshold, mode):
    world, port = 2, 31500 + (os.getpid() % 2000) + (7 if mode == "dense" else 0)
    mp.spawn(sparse_worker, args=(world, port, str(tmp_path), threshold), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "s0.pt"), torch.load(tmp_path / "s1.pt")
    assert r0["info"]["mode"] == mode
    assert torch.equal(r0["arena"], r1["arena"]) and torch.equal(r0["max_buf"], r1["max_buf"])
    a, b = FakeBuffers(N, C), FakeBuffers(N, C)
    fill_rank(a, 0)
    fill_rank(b, 1)
    assert torch.equal(r0["arena"], a.grad_arena + b.grad_arena), "sparse exchange must equal the dense sum"
    assert torch.equal(r0["max_buf"][:N], torch.maximum(a.touch_mask, b.touch_mask))
    assert torch.equal(r0["max_buf"][N:].view(torch.float32), torch.maximum(a.step_max_radii, b.step_max_radii))


def nosync_worker(rank, world, port, out_dir):
    """Three consecutive exchanges with `state`: the first blocks on M and sets the row capacity, the next two
    size the collective on that capacity and read M "on the device"; the last one is made to overflow."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cuda_gaussian_splatting_b200 import parallel
    state, outs = {}, []
    for step in range(3):
        b = FakeBuffers(N, C)
        fill_rank(b, rank + 10 * step)
        info = parallel.sparse_allreduce_step(b, with_stats=True, ops=TorchRowOps(), state=state)
        outs.append({"arena": b.grad_arena.clone(), "info": dict(info)})
    # force an overflow: a capacity far below the union
    state["m_cap"], state["pending"] = 8, None
    b = FakeBuffers(N, C)
    fill_rank(b, rank)
    parallel.sparse_allreduce_step(b, with_stats=True, ops=TorchRowOps(), state=state)
    b2 = FakeBuffers(N, C)
    fill_rank(b2, rank)
    parallel.sparse_allreduce_step(b2, with_stats=True, ops=TorchRowOps(), state=state)   # sees the previous status
    torch.save({"outs": outs, "overflow": bool(state.get("overflow")), "m_cap": state["m_cap"]},
               os.path.join(out_dir, f"n{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_without_host_round_trip_for_m(tmp_path):
    world, port = 2, 33500 + (os.getpid() % 2000)
    mp.spawn(nosync_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "n0.pt"), torch.load(tmp_path / "n1.pt")
    assert r0["outs"][0]["info"]["host_sync"] is True
    assert r0["outs"][1]["info"]["host_sync"] is False and r0["outs"][2]["info"]["host_sync"] is False
    for step in range(3):
        a, b = FakeBuffers(N, C), FakeBuffers(N, C)
        fill_rank(a, 0 + 10 * step)
        fill_rank(b, 1 + 10 * step)
        assert torch.equal(r0["outs"][step]["arena"], r1["outs"][step]["arena"])
        assert torch.equal(r0["outs"][step]["arena"], a.grad_arena + b.grad_arena), f"step {step}"
    assert r0["overflow"] and r1["overflow"], "an exchange beyond its row capacity must be flagged one step later"
    assert r0["m_cap"] == r1["m_cap"] > 8, "every rank grows the capacity to the same value"
