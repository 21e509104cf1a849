"""View-parallel training plumbing: one process per GPU, Gaussian parameters
replicated, the step's batch of camera views sharded over the ranks, parameter gradients and the
additive densification statistics combined by ONE all-reduce(sum) of a contiguous FP32 arena per
step (plus a max-reduction of `max_radii`, which is not additive). The reference has no multi-GPU
code at all (single process, single GPU: training/trainer.cpp:83, :186-188); this is the data-parallel
schedule BASELINE.json names.

Everything here is host logic on `torch.distributed`; it works with NCCL (GPU) and gloo (CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, world_size: int, rank: int) -> List[int]:
    """Round-robin assignment of the step's views: rank g takes views g, g+G, g+2G, ... so that
    every rank gets floor/ceil(num_views / G) views (8/4/2 views per GPU for 16 views at G = 2/4/8)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    return list(range(rank, num_views, world_size))


def arena_layout(n: int, num_coeffs: int, align: int = 64):
    """Element offsets of the gradient arena: the five Adam groups in the reference's order
    (optimizer/fused_adam.cu:94-97: positions, sh_coeffs, opacities, scales, rotations) followed by
    the two additive densification statistics of the step (grad_accum, grad_count;
    optimizer/densification.cpp:59-88). Segment starts are aligned to `align` floats."""
    sizes = [3 * n, 3 * num_coeffs * n, n, 3 * n, 4 * n, n, n]
    names = ["positions", "sh_coeffs", "opacities", "scales", "rotations", "grad_accum", "grad_count"]
    out, off = {}, 0
    for nm, sz in zip(names, sizes):
        out[nm] = (off, sz)
        off += (sz + align - 1) // align * align
    return out, max(off, align)


def allreduce_step(arena: torch.Tensor, max_radii: Optional[torch.Tensor] = None,
                   group: Optional[dist.ProcessGroup] = None, async_op: bool = False):
    """The step's collective: sum the arena over the ranks (in place); `max_radii` (if given) is
    max-reduced. Returns the work handles when async_op. Single-process runs are a no-op."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return []
    works = [dist.all_reduce(arena, op=dist.ReduceOp.SUM, group=group, async_op=async_op)]
    if max_radii is not None:
        works.append(dist.all_reduce(max_radii, op=dist.ReduceOp.MAX, group=group, async_op=async_op))
    return works if async_op else []


def grad_scale_for(num_views_total: int) -> float:
    """Adam consumes the SUM over the step's views scaled by 1/views (mean gradient)."""
    return 1.0 / float(max(num_views_total, 1))


def fold_step_stats(step_grad_accum: torch.Tensor, step_grad_count: torch.Tensor, step_max_radii: torch.Tensor,
                    grad_accum: torch.Tensor, grad_count: torch.Tensor, max_radii: torch.Tensor) -> None:
    """Fold the (all-reduced) per-step statistics into the persistent densification accumulators
    (densification.cpp:77-86 applied once per view, summed over the step's views)."""
    grad_accum.add_(step_grad_accum)
    grad_count.add_(step_grad_count)
    torch.maximum(max_radii, step_max_radii, out=max_radii)
