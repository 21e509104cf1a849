"""Host side of the model-resizing steps that follow the hot path on a schedule:
``DensificationController`` (optimizer/densification.{hpp,cpp}) and MCMC relocation
(optimizer/mcmc_densification.cpp:56-138). The policy is the reference's, written against the C ABI:
classification and the moves are CUDA kernels (csrc/density_control.cu); only the rarely taken
``max_gaussians`` budget cap stays a host decision on the flag array (torch.topk, as the reference).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import CugsDensifyConfig
from .rasterizer import GaussianModel, _lib_and_handle, _ptr, _stream
from .training import DensificationStats, FusedAdam, MCMCConfig

KEEP, CLONE, SPLIT = 1, 2, 4  # CUGS_DENSIFY_* of include/cugs_b200.h
RESET_OPACITY = -4.59511985013459  # inverse_sigmoid(0.01), densification.cpp:26


@dataclass
class DensificationConfig:  # optimizer/densification.hpp:23-44 (without the 6 GB-laptop VRAM guard)
    densify_from: int = 500
    densify_until: int = 15000
    densify_every: int = 100
    opacity_reset_every: int = 3000
    grad_threshold: float = 0.0002
    opacity_threshold: float = 0.005
    percent_dense: float = 0.01
    max_screen_size: int = 20
    max_gaussians: int = 0


@dataclass
class DensificationResult:  # the reference's `DensificationStats`, densification.hpp:47-54
    num_cloned: int = 0
    num_split: int = 0
    num_pruned: int = 0
    num_before: int = 0
    num_after: int = 0


def _f32(x) -> float:
    return float(np.float32(x))


class DensificationController(DensificationStats):
    """optimizer/densification.hpp:66-160. The accumulators (``accumulate_gradients``) are inherited
    from :class:`DensificationStats`; ``densify`` = clone / split / prune in the reference's order."""

    def __init__(self, config: Optional[DensificationConfig] = None, scene_extent: float = 1.0, n: int = 0,
                 device="cuda", seed: int = 0xD5171F):
        super().__init__(n, device)
        self.config = config or DensificationConfig()
        self.scene_extent = float(scene_extent)
        self.seed = int(seed)
        self.last_split_normals = None

    # ---- schedule (densification.cpp:41-51) ----
    def should_densify(self, step: int) -> bool:
        c = self.config
        return c.densify_from <= step <= c.densify_until and step % c.densify_every == 0

    def should_reset_opacity(self, step: int) -> bool:
        c = self.config
        return c.opacity_reset_every > 0 and step >= c.densify_from and step % c.opacity_reset_every == 0

    def reset_opacity(self, model: GaussianModel) -> None:  # densification.cpp:335-338
        model.opacities.fill_(RESET_OPACITY)

    def reset_accumulators(self, n: int, device) -> None:  # densification.cpp:344-349
        self._reset(n, device)

    # ---- classification ----
    def _native_config(self, step: int) -> CugsDensifyConfig:
        c = self.config
        return CugsDensifyConfig(
            grad_threshold=c.grad_threshold,
            size_threshold=_f32(np.float32(c.percent_dense) * np.float32(self.scene_extent)),   # :355 / :393
            opacity_threshold=c.opacity_threshold,
            apply_size_pruning=int(c.opacity_reset_every > 0 and step > c.opacity_reset_every),  # :416-417
            max_screen_size=float(c.max_screen_size),
            ws_threshold=_f32(np.float32(0.1) * np.float32(self.scene_extent)))                  # :438

    def classify(self, model: GaussianModel, step: int):
        """-> (flags [N] u8, (kept originals, clones, splits), temp). One kernel, one host sync."""
        dev = model.positions.device
        n = model.num_gaussians()
        if self.grad_accum.shape[0] != n:
            # compute_split_mask pads missing accumulators with zeros (:380-385); the clone mask would
            # throw in the reference, here both see zeros
            self.reset_accumulators(n, dev)
        lib, h = _lib_and_handle(dev)
        flags = torch.empty((max(n, 1),), dtype=torch.uint8, device=dev)
        temp = torch.empty((lib.cugs_b200_densify_temp_bytes(n),), dtype=torch.uint8, device=dev)
        counts = (C.c_int64 * 3)()
        cfg = self._native_config(step)
        st = lib.cugs_b200_densify_classify(h, _stream(dev), n, _ptr(model.scales), _ptr(model.opacities),
                                            _ptr(self.grad_accum), _ptr(self.grad_count), _ptr(self.max_radii_2d),
                                            C.byref(cfg), _ptr(flags), counts, _ptr(temp), temp.numel())
        _lib.check(h, st, "cugs_b200_densify_classify")
        return flags[:n], (int(counts[0]), int(counts[1]), int(counts[2])), temp

    def _cap(self, flags: torch.Tensor, bit: int, budget: int) -> None:
        """Keep only the `budget` highest-average-gradient candidates of `bit` (densification.cpp:128-137,
        :196-213): avg_grad.masked_fill(~mask, -1).topk(budget)."""
        mask = (flags & bit) != 0
        flags &= ~bit & 0xFF
        if budget <= 0:
            return
        avg = self.grad_accum / self.grad_count.clamp_min(1)
        idx = avg.masked_fill(~mask, -1.0).topk(budget).indices
        flags[idx] |= bit

    # ---- densify (densification.cpp:94-329) ----
    def densify(self, model: GaussianModel, step: int, optimizer: Optional[FusedAdam] = None,
                carry_optimizer_state: bool = False, return_normals: bool = False) -> DensificationResult:
        """Clone, split and prune ``model`` in place (its five tensors are replaced). ``optimizer``: re-bound
        to the new tensors when the model changed -- with fresh moments and step count like the reference's
        rebuild (trainer.cpp:281-289), or, with ``carry_optimizer_state``, with the moments of the kept rows
        carried over by the same kernel (new rows start at zero)."""
        n = model.num_gaussians()
        res = DensificationResult(num_before=n, num_after=n)
        if n == 0:
            return res
        dev = model.positions.device
        c = self.config
        flags, (kept, n_clone, n_split), temp = self.classify(model, step)
        capped = False
        if c.max_gaussians > 0:  # budget caps: host policy on the flag array
            if n_clone > 0 and n_clone > c.max_gaussians - n:
                self._cap(flags, CLONE, c.max_gaussians - n)
                n_clone = max(min(n_clone, c.max_gaussians - n), 0)
                capped = True
            if n_split > 0:
                budget = int((c.max_gaussians - (n + n_clone)) / 2)  # C++ int division truncates (:191-192)
                if n_split > budget:
                    self._cap(flags, SPLIT, budget)
                    n_split = max(min(n_split, budget), 0)
                    capped = True
        if capped:
            kept = int((((flags & KEEP) != 0) & ((flags & SPLIT) == 0)).sum().item())
        n_out = kept + n_clone + 2 * n_split
        res.num_cloned, res.num_split = n_clone, n_split
        res.num_pruned = n + n_clone + 2 * n_split - n_out  # :317-319 (split originals count as pruned)
        res.num_after = n_out
        changed = n_clone > 0 or n_split > 0 or n_out != n
        if changed:
            lib, h = _lib_and_handle(dev)
            src = [model.positions, model.sh_coeffs, model.opacities, model.scales, model.rotations]
            num_coeffs = int(model.sh_coeffs.shape[2])
            shapes = [(n_out, 3), (n_out, 3, num_coeffs), (n_out, 1), (n_out, 3), (n_out, 4)]
            dst = [torch.empty(s, dtype=torch.float32, device=dev) for s in shapes]
            arr = lambda ts: (C.c_void_p * 5)(*[t.data_ptr() for t in ts])
            carry = carry_optimizer_state and optimizer is not None
            dm = [torch.empty_like(t) for t in dst] if carry else None
            dv = [torch.empty_like(t) for t in dst] if carry else None
            normals = (torch.empty((2 * n_split, 3), dtype=torch.float32, device=dev)
                       if return_normals and n_split > 0 else None)
            st = lib.cugs_b200_densify_apply(h, _stream(dev), n, n_out, num_coeffs, _ptr(flags), arr(src), arr(dst),
                                             arr(optimizer.m) if carry else None, arr(optimizer.v) if carry else None,
                                             arr(dm) if carry else None, arr(dv) if carry else None,
                                             self.seed + step, _ptr(normals), _ptr(temp), temp.numel())
            _lib.check(h, st, "cugs_b200_densify_apply")
            self.last_split_normals = normals
            (model.positions, model.sh_coeffs, model.opacities, model.scales, model.rotations) = dst
            if optimizer is not None:
                optimizer.rebind(model, dm, dv)
                optimizer.update_lr(step)  # trainer.cpp:283
        self.reset_accumulators(n_out, dev)  # :326
        return res


# ----------------------------------------------------------------------------------------------
# MCMC relocation (optimizer/mcmc_densification.cpp:29-33, :56-138)
# ----------------------------------------------------------------------------------------------
@dataclass
class MCMCStats:  # mcmc_densification.hpp:54-59
    num_relocated: int = 0
    num_dead: int = 0
    num_total: int = 0


def mcmc_should_relocate(step: int, config: MCMCConfig) -> bool:
    return config.relocate_from <= step <= config.relocate_until and step % config.relocate_every == 0


def mcmc_relocate(model: GaussianModel, step: int, config: MCMCConfig, scene_extent: float = 1.0,
                  want_stats: bool = True, debug: Optional[dict] = None) -> MCMCStats:
    """MCMCController::relocate, in place, four small kernels; N stays constant so the optimizer is not
    touched (trainer.cpp:265). ``want_stats=False`` skips the host read of the counts (no sync).
    ``debug``: filled with the chosen sources [N] (-1 = untouched) and the jitter normals [N,3]."""
    n = model.num_gaussians()
    stats = MCMCStats(num_total=n)
    if n == 0:
        return stats
    dev = model.positions.device
    lib, h = _lib_and_handle(dev)
    max_relocate = int(np.float32(config.relocate_cap) * np.float32(n))  # static_cast<int>(cap * n), :95
    temp = torch.empty((lib.cugs_b200_mcmc_relocate_temp_bytes(n),), dtype=torch.uint8, device=dev)
    counts = (C.c_int64 * 2)() if want_stats else None
    source = normals = None
    if debug is not None:
        source = torch.empty((n,), dtype=torch.int32, device=dev)
        normals = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    st = lib.cugs_b200_mcmc_relocate(h, _stream(dev), n, int(model.sh_coeffs.shape[2]), _ptr(model.positions),
                                     _ptr(model.sh_coeffs), _ptr(model.opacities), _ptr(model.scales),
                                     _ptr(model.rotations), config.dead_opacity_threshold, max_relocate,
                                     float(scene_extent), int(config.seed), int(step), _ptr(source), _ptr(normals),
                                     counts, _ptr(temp), temp.numel())
    _lib.check(h, st, "cugs_b200_mcmc_relocate")
    if debug is not None:
        debug["source"], debug["normals"] = source, normals
    if want_stats:
        stats.num_dead, stats.num_relocated = int(counts[0]), int(counts[1])
    return stats
