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
