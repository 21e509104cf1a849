"""Python face of the C++ training step (include/cugs_b200.h ``cugs_b200_trainer_*``; csrc/trainer.cu):
``Trainer::train_step`` of the reference (training/trainer.cpp:178-316) without the Dataset, run as ONE
captured CUDA graph per step with no host synchronisation. torch is used for device memory only; the step
sequence, the learning-rate / SH-degree schedules and the graph live in the library.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import CugsTrainConfig, CugsTrainTensors, CugsView
from .rasterizer import CameraInfo, GaussianModel, RenderSettings, _check, _lib_and_handle, make_view
from .training import DensificationStats, TrainConfig


class NativeTrainer:
    """The C++ step driver. Same step as :class:`training.SyntheticTrainer` (and as the reference's
    ``Trainer::train_step``), but sequenced by the library: no per-view host round trip, one CUDA graph per
    step. ``pair_capacity`` bounds the (tile, Gaussian) pairs of one frame; ``None`` measures it with one
    blocking render per view and adds 25 % head-room. ``train_step`` raises nothing when a frame overflows:
    ``result()`` reports ``ok = False`` (the update was skipped on the device) and ``grow_and_retry`` handles it.

    View-parallel training: construct with ``total_views_per_step`` = views of all ranks and call
    ``step_views(step)`` -> gradient exchange -> ``step_update(step)``."""

    VIEWS, UPDATE, BOTH, VIEWS_UNTIL_MASK, VIEWS_REST = 1, 2, 3, 4, 8

    def __init__(self, model: GaussianModel, cameras: Sequence[CameraInfo], targets: Sequence[torch.Tensor],
                 config: Optional[TrainConfig] = None, total_views_per_step: Optional[int] = None,
                 pair_capacity: Optional[int] = None, frames_in_flight: int = 2, use_graph: bool = True,
                 sparse_rows: bool = True, grad_buffers=None, dL_dcolors=None):
        from .rasterizer import FrameBuffers, render
        self.model, self.cameras, self.targets = model, list(cameras), [None if t is None else t.contiguous() for t in targets]
        self.config = config or TrainConfig()
        cfg = self.config
        dev = model.positions.device
        self.device = dev
        self.lib, self.h = _lib_and_handle(dev)
        n = model.num_gaussians()
        cam0 = self.cameras[0]
        W, H, Cn = cam0.width, cam0.height, int(model.sh_coeffs.shape[2])
        self.n, self.W, self.H, self.C = n, W, H, Cn
        for p in (model.positions, model.sh_coeffs, model.opacities, model.scales, model.rotations):
            _check(p.is_cuda and p.is_contiguous() and p.dtype == torch.float32, "params must be contiguous f32 CUDA tensors")
        # gradient arena / statistics / mask: the FrameBuffers layout, so the view-parallel exchange
        # (parallel.sparse_allreduce_step) works on the same object
        self.buffers = grad_buffers if grad_buffers is not None else FrameBuffers(n, W, H, Cn, dev)
        b = self.buffers
        if pair_capacity is None:
            settings = RenderSettings(cfg.background, cfg.max_sh_degree, 1.0)
            pmax = 0
            for cam in self.cameras:
                pmax = max(pmax, int(render(model, cam, settings, b).gaussian_indices.numel()))
            pair_capacity = int(pmax * 1.25) + 4096
        self.pair_capacity = int(pair_capacity)
        self.m = [torch.zeros_like(p) for p in self._params()]
        self.v = [torch.zeros_like(p) for p in self._params()]
        self.multi_rank = total_views_per_step is not None and total_views_per_step > len(self.cameras)
        self.stats = DensificationStats(n, dev)
        self.sparse_rows = sparse_rows
        self.total_views = int(total_views_per_step or len(self.cameras))
        self.frames_in_flight, self.use_graph = int(frames_in_flight), bool(use_graph)
        self._t = None
        self._dLs = list(dL_dcolors) if dL_dcolors is not None else None
        self._create()

    def _params(self):
        m = self.model
        return [m.positions, m.sh_coeffs, m.opacities, m.scales, m.rotations]

    def _native_config(self) -> CugsTrainConfig:
        cfg, a = self.config, self.config.adam
        mc = cfg.mcmc
        c = CugsTrainConfig()
        c.lambda_ssim, c.max_sh_degree = float(cfg.lambda_ssim), int(cfg.max_sh_degree)
        for k in range(3):
            c.background[k] = float(cfg.background[k])
        c.lr_position_init, c.lr_position_final = a.position_lr_config.lr_init, a.position_lr_config.lr_final
        c.lr_position_max_steps = a.position_lr_config.max_steps
        c.lr_sh_coeffs, c.lr_opacities, c.lr_scales, c.lr_rotations = a.lr_sh_coeffs, a.lr_opacities, a.lr_scales, a.lr_rotations
        c.beta1, c.beta2, c.eps = a.beta1, a.beta2, a.eps
        c.accumulate_stats = int(bool(cfg.densify) and mc is None)   # the ADC statistics are unused in MCMC mode
        c.mcmc = int(mc is not None)
        if mc is not None:
            c.lambda_opacity, c.lambda_scale = mc.lambda_opacity, mc.lambda_scale
            c.noise_lr_init, c.noise_lr_final, c.noise_lr_max_steps = mc.noise_lr_init, mc.noise_lr_final, mc.noise_lr_max_steps
            c.noise_gate_k, c.noise_gate_t, c.noise_seed = mc.noise_gate_k, mc.noise_gate_t, int(mc.seed)
        c.frames_in_flight, c.use_graph = self.frames_in_flight, int(self.use_graph)
        return c

    def _create(self, adam_steps: int = 0):
        lib, b = self.lib, self.buffers
        if self._t is not None:
            lib.cugs_b200_trainer_destroy(self._t)
            self._t = None
        need = lib.cugs_b200_trainer_workspace_bytes(self.n, self.C, self.W, self.H, self.pair_capacity,
                                                     self.frames_in_flight)
        self.workspace = torch.empty((need,), dtype=torch.uint8, device=self.device)
        t = CugsTrainTensors()
        grads = [b.dL_dpositions, b.dL_dsh_coeffs, b.dL_dopacities, b.dL_dscales, b.dL_drotations]
        for k in range(5):
            t.params[k] = self._params()[k].data_ptr()
            t.adam_m[k], t.adam_v[k] = self.m[k].data_ptr(), self.v[k].data_ptr()
            t.grads[k] = grads[k].data_ptr()
        t.dL_dmeans_2d = b.dL_dmeans_2d.data_ptr()
        cfgn = self._native_config()
        if cfgn.accumulate_stats:
            if self.multi_rank:   # per-step statistics live in the arena and are summed by the exchange
                t.grad_accum, t.grad_count = b.step_grad_accum.data_ptr(), b.step_grad_count.data_ptr()
                t.max_radii = b.step_max_radii.data_ptr()
            else:
                t.grad_accum, t.grad_count = self.stats.grad_accum.data_ptr(), self.stats.grad_count.data_ptr()
                t.max_radii = self.stats.max_radii_2d.data_ptr()
        t.touch_mask = b.touch_mask.data_ptr() if self.sparse_rows else None
        out = C.c_void_p()
        st = lib.cugs_b200_trainer_create(self.h, self.n, self.C, self.W, self.H, self.pair_capacity, C.byref(cfgn),
                                          C.byref(t), self.workspace.data_ptr(), self.workspace.numel(), C.byref(out))
        _lib.check(self.h, st, "cugs_b200_trainer_create")
        self._t = out.value
        lib.cugs_b200_trainer_set_adam_steps(self._t, int(adam_steps))
        self.set_views(self.cameras, self.targets, self._dLs, getattr(self, "_targets_host", None))

    def set_views(self, cameras: Sequence[CameraInfo], targets: Sequence[torch.Tensor],
                  dL_dcolors: Optional[Sequence[Optional[torch.Tensor]]] = None,
                  targets_host: Optional[Sequence[Optional[torch.Tensor]]] = None) -> None:
        """``dL_dcolors`` (optional, per view): a given dL/dcolor replaces the loss of that view (forward +
        backward only -- the headline benchmark's step). ``targets_host`` (optional, per view): PINNED host images
        copied into ``targets[v]`` inside every step, under the rendering of the view (end-to-end mode)."""
        self.cameras, self.targets = list(cameras), [None if t is None else t.contiguous() for t in targets]
        V = len(self.cameras)
        views = (CugsView * V)()
        tg = (C.c_void_p * V)()
        dl = (C.c_void_p * V)()
        th = (C.c_void_p * V)()
        self._targets_host = list(targets_host) if targets_host is not None else [None] * V
        for k, ht in enumerate(self._targets_host):
            if ht is not None:
                _check(tuple(ht.shape) == (self.H, self.W, 3) and ht.dtype == torch.float32 and ht.is_pinned()
                       and ht.is_contiguous(), "targets_host must be pinned contiguous [H, W, 3] float32 tensors")
                th[k] = ht.data_ptr()
        self._dLs = list(dL_dcolors) if dL_dcolors is not None else [None] * V
        for k, g in enumerate(self._dLs):
            if g is not None:
                _check(tuple(g.shape) == (self.H, self.W, 3) and g.is_cuda and g.dtype == torch.float32 and g.is_contiguous(),
                       "dL_dcolor must be a contiguous [H, W, 3] float32 CUDA tensor")
                dl[k] = g.data_ptr()
        settings = RenderSettings(self.config.background, self.config.max_sh_degree, 1.0)
        for k, (cam, tgt) in enumerate(zip(self.cameras, self.targets)):
            _check(tgt is not None or self._dLs[k] is not None, "a view needs a target image or a given dL/dcolor")
            views[k] = make_view(cam, settings, self.config.max_sh_degree, self.C)
            if tgt is not None:
                _check(tuple(tgt.shape) == (self.H, self.W, 3) and tgt.is_cuda and tgt.dtype == torch.float32,
                       "targets must be [H, W, 3] float32 CUDA tensors")
                tg[k] = tgt.data_ptr()
        st = self.lib.cugs_b200_trainer_set_views(self._t, V, views, tg, dl, th, self.total_views)
        _lib.check(self.h, st, "cugs_b200_trainer_set_views")

    def _step(self, step: int, phases: int) -> None:
        s = torch.cuda.current_stream(self.device).cuda_stream
        st = self.lib.cugs_b200_trainer_step(self._t, s, int(step), int(phases))
        _lib.check(self.h, st, "cugs_b200_trainer_step")

    def train_step(self, step: int) -> None:
        """One whole step (single GPU): views + update, nothing blocks."""
        self._step(step, self.BOTH)

    def step_views(self, step: int) -> None:
        if self.multi_rank and self.config.densify and self.config.mcmc is None:
            b = self.buffers
            b.step_grad_accum.zero_(); b.step_grad_count.zero_(); b.step_max_radii.zero_()
        self._step(step, self.VIEWS)

    def step_views_until_mask(self, step: int) -> None:
        """The views phase up to and including the classification pass of the last view's backward: after it the
        touch mask, dL/dmeans_2d and the statistics are final (parallel.MaskOverlap starts the mask collective)."""
        _check(self.sparse_rows, "step_views_until_mask needs sparse_rows=True")
        if self.multi_rank and self.config.densify and self.config.mcmc is None:
            b = self.buffers
            b.step_grad_accum.zero_(); b.step_grad_count.zero_(); b.step_max_radii.zero_()
        self._step(step, self.VIEWS_UNTIL_MASK)

    def step_views_rest(self, step: int) -> None:
        """The rest of the views phase: the chain rule of the last view (does not touch the mask / statistics)."""
        self._step(step, self.VIEWS_REST)

    def step_update(self, step: int) -> None:
        self._step(step, self.UPDATE)

    def result(self):
        """Blocking: ({loss, l1, ssim} of the last step, ok, largest pair count, views)."""
        sc = (C.c_float * 3)()
        stt = (C.c_int64 * 3)()
        s = torch.cuda.current_stream(self.device).cuda_stream
        st = self.lib.cugs_b200_trainer_result(self._t, s, sc, stt)
        _lib.check(self.h, st, "cugs_b200_trainer_result")
        return [float(x) for x in sc], bool(stt[0]), int(stt[1]), int(stt[2])

    def grow(self, pairs_seen: int) -> None:
        """Re-create the native trainer with room for ``pairs_seen`` pairs per frame (after ok = False). A
        skipped step left no trace: the Adam step counter lives on the device and only advances with a step
        that ran."""
        steps = int(self.lib.cugs_b200_trainer_adam_steps(self._t))
        self.pair_capacity = int(pairs_seen * 1.25) + 4096
        self._create(adam_steps=max(steps, 0))

    def train_step_checked(self, step: int):
        """One step followed by a (blocking) look at its status; an overflowed step is repeated with larger
        buffers, so the caller always gets a completed step. Returns the step's {loss, l1, ssim}."""
        self.train_step(step)
        scalars, ok, pmax, _ = self.result()
        if not ok:
            self.grow(pmax)
            self.train_step(step)
            scalars, ok, pmax, _ = self.result()
            _check(ok, "step overflowed again after growing the pair capacity")
        return scalars

    @property
    def adam_steps(self) -> int:
        return int(self.lib.cugs_b200_trainer_adam_steps(self._t))

    def close(self) -> None:
        if self._t is not None:
            torch.cuda.synchronize(self.device)
            self.lib.cugs_b200_trainer_destroy(self._t)
            self._t = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
