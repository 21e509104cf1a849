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


class P2PExchange:
    """The step's gradient exchange as two kernels over NVLink peer memory instead of NCCL collectives around
    local gather / scatter copies (csrc/grad_exchange.cu, "Peer-to-peer exchange"): the buffers must come from
    ``FrameBuffers(..., symmetric=True)``. Per step:

        barrier -> k_xchg_masks  (MAX of [touch mask | max_radii], SUM of the two statistics, sliced over the ranks)
        barrier -> scan of the union mask + index list (local, identical on every rank; M stays on the device)
                -> k_xchg_rows   (each rank sums ITS slice of the union rows across all arenas, in rank order,
                                  and writes the result into all arenas)
        barrier

    Same sums as the all-reduce (each one computed exactly once, so every rank holds the same bits), the bytes of
    one all-reduce on NVLink, no compact buffers, no row-capacity, nothing for the host to wait for."""

    def __init__(self, buffers, group: Optional[dist.ProcessGroup] = None, use_multicast: Optional[bool] = None):
        """``use_multicast``: None = use the NVSwitch multicast mapping (multimem.ld_reduce / multimem.st: the switch
        reduces and replicates, about half the bytes per NVLink direction) when the symmetric buffers have one,
        False = unicast peer loads / stores only."""
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from .rasterizer import _check
        _check(getattr(buffers, "symmetric", False), "P2PExchange needs FrameBuffers(..., symmetric=True)")
        self.b = buffers
        self.group = group if group is not None else dist.group.WORLD
        name = self.group.group_name
        self.h_arena = symm_mem.rendezvous(buffers.grad_arena, name)
        self.h_max = symm_mem.rendezvous(buffers.max_buf, name)
        self.world, self.rank = int(self.h_arena.world_size), int(self.h_arena.rank)
        _check(self.world <= 8, "P2PExchange supports up to 8 ranks (one NVSwitch box)")
        b = buffers
        base = b.grad_arena.data_ptr()
        groups = (b.dL_dpositions, b.dL_dsh_coeffs, b.dL_dopacities, b.dL_dscales, b.dL_drotations)
        offs = [t.data_ptr() - base for t in groups]
        self._grads = (C.c_void_p * (self.world * 5))()
        self._accum = (C.c_void_p * self.world)()
        self._count = (C.c_void_p * self.world)()
        self._maxbuf = (C.c_void_p * self.world)()
        for p in range(self.world):
            pb = int(self.h_arena.buffer_ptrs[p])
            for k in range(5):
                self._grads[p * 5 + k] = pb + offs[k]
            self._accum[p] = pb + (b.step_grad_accum.data_ptr() - base)
            self._count[p] = pb + (b.step_grad_count.data_ptr() - base)
            self._maxbuf[p] = int(self.h_max.buffer_ptrs[p])
        mc_a = int(getattr(self.h_arena, "multicast_ptr", 0) or 0)
        mc_m = int(getattr(self.h_max, "multicast_ptr", 0) or 0)
        # default: multicast on a full 8-GPU box. The number of multimem operations of a rank falls with 1 / R while
        # its unicast traffic grows with (R - 1) / R: measured (profiles/r02/) unicast is faster at 2 ranks (0.31 vs
        # 0.56 ms for 531 k rows), multicast at 8 (0.70 vs 0.79 ms for 867 k rows, masks 0.15 vs 0.25 ms)
        self.multicast = (bool(mc_a and mc_m) and self.world >= 8) if use_multicast is None \
            else bool(use_multicast and mc_a and mc_m)
        self._grads_mc = (C.c_void_p * 5)()
        self._accum_mc = self._count_mc = self._maxbuf_mc = None
        if self.multicast:
            for k in range(5):
                self._grads_mc[k] = mc_a + offs[k]
            self._accum_mc = mc_a + (b.step_grad_accum.data_ptr() - base)
            self._count_mc = mc_a + (b.step_grad_count.data_ptr() - base)
            self._maxbuf_mc = mc_m
        self.ops = _CudaRowOps()

    def exchange(self, with_stats: bool = True) -> dict:
        from . import _lib
        from .rasterizer import _lib_and_handle, _stream
        b = self.b
        dev = b.grad_arena.device
        lib, h = _lib_and_handle(dev)
        n, C_ = int(b.n), int(b.dL_dsh_coeffs.shape[2])
        self.h_max.barrier(channel=0)      # every rank's mask / statistics / gradient rows of this step are final
        mc = self.multicast
        st = lib.cugs_b200_p2p_reduce_masks(h, _stream(dev), n, self.world, self.rank, self._maxbuf,
                                            self._accum if with_stats else None, self._count if with_stats else None,
                                            self._maxbuf_mc if mc else None,
                                            self._accum_mc if (mc and with_stats) else None,
                                            self._count_mc if (mc and with_stats) else None)
        _lib.check(h, st, "cugs_b200_p2p_reduce_masks")
        self.h_max.barrier(channel=1)      # all slices of the union mask have landed
        offsets, m_dev = self.ops.scan_dev(b)
        st = lib.cugs_b200_build_touch_index(h, _stream(dev), n, b.touch_mask.data_ptr(), offsets.data_ptr(),
                                             b._touch_idx.data_ptr())
        _lib.check(h, st, "cugs_b200_build_touch_index")
        st = lib.cugs_b200_p2p_reduce_rows(h, _stream(dev), n, C_, self.world, self.rank, b._touch_idx.data_ptr(),
                                           m_dev.data_ptr(), self._grads, self._grads_mc if mc else None)
        _lib.check(h, st, "cugs_b200_p2p_reduce_rows")
        self.h_arena.barrier(channel=0)    # all sums written everywhere; nobody reads these rows any more
        return {"mode": "p2p-multicast" if mc else "p2p", "host_sync": False}


class MaskOverlap:
    """Hides the first collective of the sparse exchange -- the MAX all-reduce of [touch mask | max_radii] and
    the scan of the union mask -- under the tail of the step's last backward: the mask is final after the
    classification pass of the last view (NativeTrainer.step_views_until_mask), the chain rule of that view
    (step_views_rest) does not touch it, so the collective runs on a side stream meanwhile.

        trainer.step_views_until_mask(step); overlap.start(buffers, state)
        trainer.step_views_rest(step);       overlap.finish()
        sparse_allreduce_step(buffers, ..., state=state, mask_reduced=True)
    """

    def __init__(self, device, group: Optional[dist.ProcessGroup] = None):
        self.device = torch.device(device)
        self.group = group
        self.stream = torch.cuda.Stream(self.device)
        self.ready = torch.cuda.Event()
        self.done = torch.cuda.Event()

    def start(self, buffers, state: Optional[dict] = None, ops=None) -> None:
        cur = torch.cuda.current_stream(self.device)
        self.ready.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.ready)
            dist.all_reduce(buffers.max_buf, op=dist.ReduceOp.MAX, group=self.group)
            if state is not None and state.get("m_cap"):
                ops = ops or _CudaRowOps()
                state["pre_scanned"] = ops.scan_dev(buffers)   # offsets + device-side M of the union mask
            self.done.record(self.stream)

    def finish(self) -> None:
        torch.cuda.current_stream(self.device).wait_event(self.done)


def sparse_allreduce_step(buffers, with_stats: bool = True, dense_threshold: float = 0.6,
                          group: Optional[dist.ProcessGroup] = None, ops=None, state: Optional[dict] = None,
                          mask_reduced: bool = False) -> dict:
    """The step's gradient exchange, exploiting that a view leaves most Gaussians' gradients exactly
    zero: (1) ONE int32 MAX all-reduce of [touch mask | max_radii bits] (8 B/Gaussian), (2) exclusive
    scan of the union mask -> M touched Gaussians, (3) the M gradient rows are gathered into a dense
    buffer, followed by the two additive statistics, (4) ONE all-reduce(sum) of (59 M + 2 N) floats
    instead of 61 N, (5) scatter back. Falls back to the dense all-reduce when M > dense_threshold * N.
    Numerically a plain sum either way.

    ``mask_reduced``: the MAX all-reduce (and, with ``state``, the scan) was already done by :class:`MaskOverlap`.

    ``state`` (a dict the caller keeps between steps) removes the host round trip for M: the collective is
    sized on a row CAPACITY derived from the previous steps' M (identical on every rank, because M is the
    size of the all-reduced union), the real count is read on the device, and {M, M > capacity} reaches the
    host asynchronously; it is looked at one step later (``state["overflow"]`` is then set and the caller must
    repeat that step -- the capacity is ``state["headroom"]`` (default 1.15) x the M that set it and grows as
    soon as M comes within 5 % of it, which makes an overflow a non-event for a view set that changes
    gradually). Without ``state`` the call blocks once on M, as in round 1.

    ``buffers``: FrameBuffers (grad_arena, max_buf, touch_mask, dL_d* views). ``ops``: object with
    scan(mask)->(offsets, M), gather(...), scatter(...); defaults to the CUDA library (tests inject
    a torch implementation to exercise the host logic on CPU with gloo)."""
    b = buffers
    n = int(b.n)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return {"mode": "single", "touched": None}
    if not mask_reduced:   # (MaskOverlap already reduced it under the last backward)
        dist.all_reduce(b.max_buf, op=dist.ReduceOp.MAX, group=group)
    ops = ops or _CudaRowOps()
    num_coeffs = int(b.dL_dsh_coeffs.shape[2])

    def finish(m_layout, offsets, m_for_ops, m_dev):
        rows = ops.compact_floats(m_layout, num_coeffs)  # the gradient rows, group-major
        need = rows + (2 * n if with_stats else 0)
        if b.grad_compact is None or b.grad_compact.numel() < need:
            # zeros: the <= 3 pad floats between the group blocks are never written by the gather
            b.grad_compact = torch.zeros((int(need * 1.25) + 1024,), dtype=torch.float32, device=b.grad_arena.device)
        compact = b.grad_compact[:need]
        ops.gather(b, offsets, m_for_ops, compact, m_dev)
        if with_stats:
            compact[rows:rows + n].copy_(b.step_grad_accum)
            compact[rows + n:].copy_(b.step_grad_count)
        dist.all_reduce(compact, op=dist.ReduceOp.SUM, group=group)
        ops.scatter(b, offsets, m_for_ops, compact, m_dev)
        if with_stats:
            b.step_grad_accum.copy_(compact[rows:rows + n])
            b.step_grad_count.copy_(compact[rows + n:])
        return need

    if state is not None and state.get("m_cap") and hasattr(ops, "scan_dev"):
        # look at the status of the PREVIOUS exchange (its copy was queued a whole step ago)
        prev = state.get("pending")
        if prev is not None:
            prev["event"].synchronize()
            m_prev, over = int(prev["host"][0]), bool(int(prev["host"][1]))
            state["m_seen"] = max(state.get("m_seen", 0), m_prev)
            if over:
                state["overflow"] = True
            if m_prev > dense_threshold * n:
                state["m_cap"] = None          # go back to the blocking path, which switches to dense
            elif m_prev * 1.05 > state["m_cap"]:
                state["m_cap"] = min(n, int(m_prev * state.get("headroom", 1.15)) + 1024)
            state["pending"] = None
    if state is not None and state.get("m_cap") and hasattr(ops, "scan_dev"):
        m_cap = int(state["m_cap"])
        pre = state.pop("pre_scanned", None) if mask_reduced else None
        offsets, m_dev = pre if pre is not None else ops.scan_dev(b)
        if "status_dev" not in state:
            dev = b.grad_arena.device
            state["status_dev"] = torch.zeros((2,), dtype=torch.int64, device=dev)
            pin = (lambda t: t.pin_memory()) if dev.type == "cuda" else (lambda t: t)
            state["status_host"] = [pin(torch.zeros((2,), dtype=torch.int64)) for _ in range(2)]
            state["flip"] = 0
        ops.status_dev = state["status_dev"]
        need = finish(m_cap, offsets, m_cap, m_dev)
        host = state["status_host"][state["flip"]]
        state["flip"] ^= 1
        host.copy_(state["status_dev"], non_blocking=True)
        ev = _DoneEvent()
        if b.grad_arena.device.type == "cuda":
            ev = torch.cuda.Event()
            ev.record()
        state["pending"] = {"host": host, "event": ev}
        return {"mode": "sparse", "touched": state.get("m_seen"), "row_capacity": m_cap, "floats": need,
                "host_sync": False}

    offsets, m = ops.scan(b)
    if m > dense_threshold * n:
        dist.all_reduce(b.grad_arena, op=dist.ReduceOp.SUM, group=group)
        return {"mode": "dense", "touched": m}
    need = finish(m, offsets, m, None)
    if state is not None:
        state["m_cap"] = min(n, int(m * state.get("headroom", 1.15)) + 1024)
        state["m_seen"] = max(state.get("m_seen", 0), m)
    return {"mode": "sparse", "touched": m, "floats": need, "host_sync": True}


class _DoneEvent:
    """Stand-in for torch.cuda.Event on the CPU (gloo) test path: copies there are synchronous."""

    def synchronize(self):
        return None


class _CudaRowOps:
    """scan / gather / scatter on the CUDA library (cugs_b200_scan, cugs_b200_{gather,scatter}_grad_rows)."""

    def _groups(self, b):
        import ctypes as C
        arr = (C.c_void_p * 5)()
        for k, t in enumerate((b.dL_dpositions, b.dL_dsh_coeffs, b.dL_dopacities, b.dL_dscales, b.dL_drotations)):
            arr[k] = t.data_ptr()
        return arr

    def compact_floats(self, m, num_coeffs):
        from . import _lib
        return int(_lib.load_library().cugs_b200_compact_grad_floats(int(m), int(num_coeffs)))

    def scan(self, b):
        import ctypes as C
        from . import _lib
        from .rasterizer import _lib_and_handle, _stream
        dev = b.grad_arena.device
        lib, h = _lib_and_handle(dev)
        n = int(b.n)
        if b.touch_offsets is None:
            b.touch_offsets = torch.empty((n,), dtype=torch.int32, device=dev)
            b._scan_tmp = torch.empty((lib.cugs_b200_scan_temp_bytes(n),), dtype=torch.uint8, device=dev)
            b._touch_idx = torch.empty((n,), dtype=torch.int32, device=dev)
        total = C.c_int64(0)
        st = lib.cugs_b200_scan(h, _stream(dev), n, b.touch_mask.data_ptr(), b.touch_offsets.data_ptr(), None,
                                C.byref(total), b._scan_tmp.data_ptr(), b._scan_tmp.numel())
        _lib.check(h, st, "cugs_b200_scan")
        return b.touch_offsets, int(total.value)

    status_dev = None  # set by sparse_allreduce_step in the no-host-sync mode: receives {M, M > capacity}

    def scan_dev(self, b):
        """Scan of the union mask with the total left on the device (no host round trip)."""
        from . import _lib
        from .rasterizer import _lib_and_handle, _stream
        dev = b.grad_arena.device
        lib, h = _lib_and_handle(dev)
        n = int(b.n)
        if b.touch_offsets is None:
            b.touch_offsets = torch.empty((n,), dtype=torch.int32, device=dev)
            b._scan_tmp = torch.empty((lib.cugs_b200_scan_temp_bytes(n),), dtype=torch.uint8, device=dev)
            b._touch_idx = torch.empty((n,), dtype=torch.int32, device=dev)
        if getattr(b, "_touch_total", None) is None:
            b._touch_total = torch.zeros((1,), dtype=torch.int64, device=dev)
        st = lib.cugs_b200_scan(h, _stream(dev), n, b.touch_mask.data_ptr(), b.touch_offsets.data_ptr(),
                                b._touch_total.data_ptr(), None, b._scan_tmp.data_ptr(), b._scan_tmp.numel())
        _lib.check(h, st, "cugs_b200_scan")
        return b.touch_offsets, b._touch_total

    def gather(self, b, offsets, m, compact, m_dev=None):
        from . import _lib
        from .rasterizer import _lib_and_handle, _stream
        dev = b.grad_arena.device
        lib, h = _lib_and_handle(dev)
        st = lib.cugs_b200_gather_grad_rows(h, _stream(dev), int(b.n), int(b.dL_dsh_coeffs.shape[2]),
                                            b.touch_mask.data_ptr(), offsets.data_ptr(), int(m), self._groups(b),
                                            compact.data_ptr(), b._touch_idx.data_ptr(),
                                            None if m_dev is None else m_dev.data_ptr(),
                                            None if (m_dev is None or self.status_dev is None) else self.status_dev.data_ptr())
        _lib.check(h, st, "cugs_b200_gather_grad_rows")

    def scatter(self, b, offsets, m, compact, m_dev=None):
        from . import _lib
        from .rasterizer import _lib_and_handle, _stream
        dev = b.grad_arena.device
        lib, h = _lib_and_handle(dev)
        # touch = offsets = None: the index list built by this exchange's gather is still in b._touch_idx
        st = lib.cugs_b200_scatter_grad_rows(h, _stream(dev), int(b.n), int(b.dL_dsh_coeffs.shape[2]),
                                             None, None, int(m), compact.data_ptr(),
                                             self._groups(b), b._touch_idx.data_ptr(),
                                             None if m_dev is None else m_dev.data_ptr())
        _lib.check(h, st, "cugs_b200_scatter_grad_rows")


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
