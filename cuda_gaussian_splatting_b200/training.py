"""Host-side mirror of the two per-step pieces either side of the rasterizer, on the C ABI:

* ``combined_loss`` / ``l1_loss`` / ``ssim_loss`` (reference training/loss.hpp:22-53) — here the
  fused kernel also returns the gradient the reference obtains from libtorch autograd
  (training/trainer.cpp:214-217): ``combined_loss_with_grad``.
* ``FusedAdam`` / ``AdamConfig`` / ``position_lr`` / ``active_sh_degree_for_step``
  (optimizer/fused_adam.hpp:29-106, optimizer/adam.hpp:30-41, training/lr_schedule.hpp:36-80).
* ``DensificationStats.accumulate_gradients`` (optimizer/densification.cpp:59-88).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import _lib
from .rasterizer import BackwardOutput, GaussianModel, _check, _lib_and_handle, _ptr, _stream


# ----------------------------------------------------------------------------------------------
# loss
# ----------------------------------------------------------------------------------------------
def _validate_image(img: torch.Tensor, name: str) -> None:  # loss.cpp:15-25
    _check(img.dim() == 3, f"{name} must be 3-dimensional [H, W, 3], got {img.dim()} dims")
    _check(img.shape[2] == 3, f"{name} must have 3 channels, got {img.shape[2]}")
    _check(img.dtype == torch.float32, f"{name} must be float32, got {img.dtype}")
    _check(img.is_cuda, f"{name} must be on a CUDA device")


def _validate_pair(rendered: torch.Tensor, target: torch.Tensor) -> None:  # loss.cpp:28-34
    _validate_image(rendered, "rendered")
    _validate_image(target, "target")
    _check(rendered.shape == target.shape,
           f"rendered and target must have the same shape, got {tuple(rendered.shape)} vs {tuple(target.shape)}")


_loss_ws: dict = {}


def combined_loss_with_grad(rendered: torch.Tensor, target: torch.Tensor, lambda_: float = 0.2,
                            want_grad: bool = True, ssim_map: Optional[torch.Tensor] = None,
                            window_size: int = 11):
    """One fused pass: returns (scalars[3] = {loss, l1, mean ssim} on device, dL/d(rendered)).
    ``ssim_map`` (optional [H,W] output) receives what the reference's ``ssim()`` returns.
    ``window_size``: any odd size >= 3 (loss.cpp:91-92); 11 is the default of every reference caller."""
    _validate_pair(rendered, target)
    _check(window_size % 2 == 1, f"window_size must be odd, got {window_size}")      # loss.cpp:91
    _check(window_size >= 3, f"window_size must be >= 3, got {window_size}")          # loss.cpp:92
    dev = rendered.device
    lib, h = _lib_and_handle(dev)
    H, W = int(rendered.shape[0]), int(rendered.shape[1])
    key = (dev.index, W, H, _stream(dev))  # one scratch per stream: frames of two streams may overlap
    ws = _loss_ws.get(key)
    if ws is None:
        ws = torch.empty((lib.cugs_b200_loss_workspace_bytes(W, H),), dtype=torch.uint8, device=dev)
        _loss_ws[key] = ws
    scalars = torch.empty((3,), dtype=torch.float32, device=dev)
    grad = torch.empty_like(rendered, memory_format=torch.contiguous_format) if want_grad else None
    st = lib.cugs_b200_loss_l1_ssim(h, _stream(dev), W, H, float(lambda_), int(window_size), _ptr(rendered.contiguous()),
                                    _ptr(target.contiguous()), _ptr(grad), _ptr(scalars), _ptr(ws), ws.numel(),
                                    _ptr(ssim_map))
    _lib.check(h, st, "cugs_b200_loss_l1_ssim")
    return scalars, grad


def combined_loss(rendered, target, lambda_: float = 0.2) -> torch.Tensor:  # loss.hpp:52
    return combined_loss_with_grad(rendered, target, lambda_, want_grad=False)[0][0]


def l1_loss(rendered, target) -> torch.Tensor:  # loss.hpp:22
    return combined_loss_with_grad(rendered, target, 0.0, want_grad=False)[0][1]


def ssim_loss(rendered, target, window_size: int = 11) -> torch.Tensor:  # loss.hpp:43
    return 1.0 - combined_loss_with_grad(rendered, target, 1.0, want_grad=False, window_size=window_size)[0][2]


def ssim(rendered, target, window_size: int = 11) -> torch.Tensor:  # loss.hpp:33 -> [H,W] map
    _validate_pair(rendered, target)
    m = torch.empty(rendered.shape[:2], dtype=torch.float32, device=rendered.device)
    combined_loss_with_grad(rendered, target, 1.0, want_grad=False, ssim_map=m, window_size=window_size)
    return m


def ssim_mean(rendered, target) -> torch.Tensor:
    return combined_loss_with_grad(rendered, target, 1.0, want_grad=False)[0][2]


# ----------------------------------------------------------------------------------------------
# learning-rate schedule (training/lr_schedule.hpp)
# ----------------------------------------------------------------------------------------------
@dataclass
class PositionLRConfig:
    lr_init: float = 1.6e-4
    lr_final: float = 1.6e-6
    max_steps: int = 30000


def position_lr(step: int, config: PositionLRConfig) -> float:  # lr_schedule.hpp:49-57
    if step >= config.max_steps:
        return float(np.float32(config.lr_final))
    if step <= 0:
        return float(np.float32(config.lr_init))
    t = np.float32(step) / np.float32(config.max_steps)
    log_ratio = np.float32(math.log(np.float32(config.lr_final) / np.float32(config.lr_init)))
    return float(np.float32(config.lr_init) * np.float32(math.exp(np.float32(t * log_ratio))))


def active_sh_degree_for_step(step: int, max_degree: int) -> int:  # lr_schedule.hpp:70-72
    return min(step // 1000, max_degree)


@dataclass
class AdamConfig:  # optimizer/adam.hpp:30-41
    position_lr_config: PositionLRConfig = field(default_factory=PositionLRConfig)
    lr_sh_coeffs: float = 2.5e-3
    lr_opacities: float = 0.05
    lr_scales: float = 5e-3
    lr_rotations: float = 1e-3
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-15


@dataclass
class MCMCConfig:  # optimizer/mcmc_densification.hpp:27-51 (without the VRAM guard)
    relocate_from: int = 500
    relocate_until: int = 15000
    relocate_every: int = 100
    dead_opacity_threshold: float = 0.005
    relocate_cap: float = 0.05
    noise_lr_init: float = 5e5
    noise_lr_final: float = 1e3
    noise_lr_max_steps: int = 30000
    noise_gate_k: float = 100.0
    noise_gate_t: float = 0.995
    lambda_opacity: float = 0.01
    lambda_scale: float = 0.01
    seed: int = 0x5EED


def mcmc_noise_lr(step: int, config: MCMCConfig) -> float:
    """MCMCController::noise_lr (mcmc_densification.cpp:38-47): log-linear decay in float32."""
    if step >= config.noise_lr_max_steps:
        return float(np.float32(config.noise_lr_final))
    if step <= 0:
        return float(np.float32(config.noise_lr_init))
    t = np.float32(step) / np.float32(config.noise_lr_max_steps)
    log_ratio = np.float32(math.log(np.float32(config.noise_lr_final) / np.float32(config.noise_lr_init)))
    return float(np.float32(config.noise_lr_init) * np.float32(math.exp(np.float32(t * log_ratio))))


def mcmc_inject_noise(model: GaussianModel, step: int, config: MCMCConfig, return_normals: bool = False):
    """MCMCController::inject_noise (mcmc_densification.cpp:144-161), one kernel, in place."""
    dev = model.positions.device
    lib, h = _lib_and_handle(dev)
    n = model.num_gaussians()
    normals = torch.empty((n, 3), dtype=torch.float32, device=dev) if return_normals else None
    st = lib.cugs_b200_mcmc_inject_noise(h, _stream(dev), n, _ptr(model.positions), _ptr(model.scales),
                                         _ptr(model.opacities), mcmc_noise_lr(step, config), config.noise_gate_k,
                                         config.noise_gate_t, int(config.seed), int(step), _ptr(normals))
    _lib.check(h, st, "cugs_b200_mcmc_inject_noise")
    return normals


class FusedAdam:
    """optimizer/fused_adam.hpp:29-106. Group order positions, sh_coeffs, opacities, scales,
    rotations (fused_adam.cu:94-97); all five groups are updated by ONE kernel launch."""

    K_POSITIONS, K_SH, K_OPACITIES, K_SCALES, K_ROTATIONS = range(5)

    def __init__(self, model: GaussianModel, config: Optional[AdamConfig] = None):
        self.model = model
        self.config = config or AdamConfig()
        self.step_count = 0
        self._params = [model.positions, model.sh_coeffs, model.opacities, model.scales, model.rotations]
        for p in self._params:
            _check(p.is_cuda, "FusedAdam: param must be on CUDA")
            _check(p.is_contiguous() and p.dtype == torch.float32, "FusedAdam: params must be contiguous f32")
        self.m = [torch.zeros_like(p) for p in self._params]
        self.v = [torch.zeros_like(p) for p in self._params]
        self.grads = [None] * 5
        c = self.config
        self.learning_rates = [c.position_lr_config.lr_init, c.lr_sh_coeffs, c.lr_opacities, c.lr_scales,
                               c.lr_rotations]
        self.grad_scale = 1.0
        self.mcmc_lambda_opacity = 0.0  # > 0: MCMC regulariser gradient added inside the launch
        self.mcmc_lambda_scale = 0.0

    def rebind(self, model: GaussianModel, m=None, v=None) -> None:
        """Point the optimizer at the (resized) tensors of ``model``. Without moments this equals the
        reference's rebuild after densification (trainer.cpp:281-283: a new FusedAdam, zero moments, step
        count 0); with ``m``/``v`` (five tensors each, e.g. from cugs_b200_densify_apply) the state and the
        step count are carried over."""
        self.model = model
        self._params = [model.positions, model.sh_coeffs, model.opacities, model.scales, model.rotations]
        for p in self._params:
            _check(p.is_cuda and p.is_contiguous() and p.dtype == torch.float32,
                   "FusedAdam: params must be contiguous f32 CUDA tensors")
        if m is None or v is None:
            self.m = [torch.zeros_like(p) for p in self._params]
            self.v = [torch.zeros_like(p) for p in self._params]
            self.step_count = 0
        else:
            for p, mi, vi in zip(self._params, m, v):
                _check(mi.shape == p.shape and vi.shape == p.shape, "FusedAdam.rebind: moment shape mismatch")
            self.m, self.v = list(m), list(v)
        self.grads = [None] * 5

    def apply_gradients(self, grads: BackwardOutput) -> None:  # fused_adam.cu:113-120
        self.grads = [grads.dL_dpositions, grads.dL_dsh_coeffs, grads.dL_dopacities, grads.dL_dscales,
                      grads.dL_drotations]

    def update_lr(self, step: int) -> None:  # fused_adam.cu:122-124
        self.learning_rates[0] = position_lr(step, self.config.position_lr_config)

    def zero_grad(self) -> None:  # fused_adam.cu:126-138
        self.grads = [None] * 5

    def get_lr(self, group: int) -> float:
        return self.learning_rates[group]

    def step(self) -> None:  # fused_adam.cu:140-164
        self.step_count += 1
        b1, b2 = float(np.float32(self.config.beta1)), float(np.float32(self.config.beta2))
        bc1 = 1.0 / (1.0 - b1 ** self.step_count)  # double precision on the host (:145-149)
        bc2 = 1.0 / (1.0 - b2 ** self.step_count)
        dev = self._params[0].device
        lib, h = _lib_and_handle(dev)
        P = (C.c_void_p * 5)()
        G = (C.c_void_p * 5)()
        M = (C.c_void_p * 5)()
        V = (C.c_void_p * 5)()
        cnt = (C.c_int64 * 5)()
        lr = (C.c_float * 5)()
        keep = []
        for k in range(5):
            g = self.grads[k]
            if g is None:  # "if (!grads_[i].defined()) continue" (:157)
                cnt[k] = 0
                continue
            _check(g.is_cuda, "FusedAdam: grad must be on CUDA")
            _check(g.numel() == self._params[k].numel(),
                   f"FusedAdam: param/grad size mismatch: {self._params[k].numel()} vs {g.numel()}")
            g = g.contiguous()
            keep.append(g)
            P[k], G[k], M[k], V[k] = (self._params[k].data_ptr(), g.data_ptr(), self.m[k].data_ptr(),
                                      self.v[k].data_ptr())
            cnt[k] = self._params[k].numel()
            lr[k] = self.learning_rates[k]
        st = lib.cugs_b200_adam_step_mcmc(h, _stream(dev), P, G, M, V, cnt, lr, self.config.beta1, self.config.beta2,
                                          self.config.eps, bc1, bc2, float(self.grad_scale),
                                          float(self.mcmc_lambda_opacity), float(self.mcmc_lambda_scale))
        _lib.check(h, st, "cugs_b200_adam_step_mcmc")


class DensificationStats:
    """The per-step accumulators of DensificationController (optimizer/densification.cpp:59-88)."""

    def __init__(self, n: int, device):
        self._reset(n, device)

    def _reset(self, n: int, device) -> None:
        """(Re)allocate the three accumulators as zeros. A plain method, NOT __init__: subclasses
        (density.DensificationController) have a different constructor signature, so the lazy
        re-initialisation below must not dispatch to it."""
        self.grad_accum = torch.zeros((n,), dtype=torch.float32, device=device)
        self.grad_count = torch.zeros((n,), dtype=torch.float32, device=device)
        self.max_radii_2d = torch.zeros((n,), dtype=torch.float32, device=device)

    def as_tuple(self):
        return (self.grad_accum, self.grad_count, self.max_radii_2d)

    def accumulate_gradients(self, dL_dmeans_2d: torch.Tensor, radii: torch.Tensor) -> None:
        n = dL_dmeans_2d.shape[0]
        if self.grad_accum.shape[0] != n:  # lazy re-init (:66-68)
            self._reset(n, dL_dmeans_2d.device)
        if n == 0:
            return
        dev = dL_dmeans_2d.device
        lib, h = _lib_and_handle(dev)
        st = lib.cugs_b200_accumulate_stats(h, _stream(dev), n, _ptr(dL_dmeans_2d.contiguous()),
                                            _ptr(radii.contiguous()), _ptr(self.grad_accum),
                                            _ptr(self.grad_count), _ptr(self.max_radii_2d))
        _lib.check(h, st, "cugs_b200_accumulate_stats")


# ----------------------------------------------------------------------------------------------
# training-step driver on synthetic views (the caller of the hot path: training/trainer.cpp:178-316)
# ----------------------------------------------------------------------------------------------
@dataclass
class TrainConfig:  # the fields of training/trainer.hpp:38-75 that reach the hot path
    lambda_ssim: float = 0.2
    max_sh_degree: int = 3
    background: tuple = (0.0, 0.0, 0.0)
    adam: AdamConfig = field(default_factory=AdamConfig)
    densify: bool = True  # accumulate the ADC statistics every step (trainer.cpp:269)
    mcmc: Optional["MCMCConfig"] = None  # MCMC mode (trainer.cpp:232-237, :246-266): regulariser + noise, no ADC stats
    # model resizing on schedule (trainer.cpp:253-303); None = fixed N (the measured configurations)
    densification: Optional[object] = None   # density.DensificationConfig: ADC clone / split / prune + opacity reset
    mcmc_relocation: bool = False            # MCMC mode: relocate dead Gaussians on the MCMCConfig schedule
    scene_extent: float = 1.0
    carry_optimizer_state: bool = False      # False = the reference's optimizer rebuild after densification


class SyntheticTrainer:
    """``Trainer::train_step`` (training/trainer.cpp:178-316) without the Dataset: per step
    update_lr -> active SH degree -> for each of the rank's views: render -> fused L1+SSIM loss and
    its gradient -> render_backward (gradients summed in place over the views, densification
    statistics fused into the same launch) -> ONE all-reduce over the ranks -> ONE fused Adam launch
    with grad_scale = 1 / total views. No host synchronisation except the per-view pair-count read
    (the reference has 3-6 `.item()` syncs per step: trainer.cpp:205-207, :223-224, :308)."""

    def __init__(self, model: GaussianModel, cameras, targets, config: Optional[TrainConfig] = None,
                 total_views_per_step: Optional[int] = None):
        from .rasterizer import FrameBuffers, RenderSettings
        self.model, self.cameras, self.targets = model, list(cameras), list(targets)
        self.config = config or TrainConfig()
        self.total_views = int(total_views_per_step or len(self.cameras))
        n = model.num_gaussians()
        cam0 = self.cameras[0]
        self.buffers = FrameBuffers(n, cam0.width, cam0.height, int(model.sh_coeffs.shape[2]), model.positions.device)
        self.optimizer = FusedAdam(model, self.config.adam)
        self.optimizer.grad_scale = 1.0 / float(self.total_views)
        if self.config.mcmc is not None:
            self.optimizer.mcmc_lambda_opacity = self.config.mcmc.lambda_opacity
            self.optimizer.mcmc_lambda_scale = self.config.mcmc.lambda_scale
        if self.config.densification is not None and self.config.mcmc is None:
            from .density import DensificationController
            self.stats = DensificationController(self.config.densification, self.config.scene_extent, n,
                                                 model.positions.device)
        else:
            self.stats = DensificationStats(n, model.positions.device)
        self._RenderSettings = RenderSettings
        self._FrameBuffers = FrameBuffers
        self.last_scalars = None
        self.last_density_event = None

    def train_step(self, step: int) -> torch.Tensor:
        from .parallel import fold_step_stats, sparse_allreduce_step
        from .rasterizer import BackwardOutput, render, render_backward
        cfg, b = self.config, self.buffers
        self.optimizer.update_lr(step)                                   # trainer.cpp:180
        degree = active_sh_degree_for_step(step, cfg.max_sh_degree)      # :183
        settings = self._RenderSettings(cfg.background, degree, 1.0)
        multi = torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size() > 1
        densify = cfg.densify and cfg.mcmc is None  # the ADC statistics are unused in MCMC mode (trainer.cpp:246-266)
        if densify:
            if multi:  # per-step statistics live in the arena and are summed by the all-reduce
                b.step_grad_accum.zero_(); b.step_grad_count.zero_(); b.step_max_radii.zero_()
                stats = (b.step_grad_accum, b.step_grad_count, b.step_max_radii)
            else:
                stats = self.stats.as_tuple()
        else:
            stats = None
        scal_sum = None
        for k, (cam, target) in enumerate(zip(self.cameras, self.targets)):
            out = render(self.model, cam, settings, b)                   # :211
            scalars, dL = combined_loss_with_grad(out.color, target, cfg.lambda_ssim)  # :214-225 in one pass
            render_backward(dL, out, self.model, cam, settings, b, stats=stats, accumulate=(k > 0),
                            touch_mask=b.touch_mask, sparse_rows=True)  # :228, :269
            scal_sum = scalars if scal_sum is None else scal_sum + scalars
        if multi:
            self.last_exchange = sparse_allreduce_step(b, with_stats=densify)
            if densify:
                fold_step_stats(b.step_grad_accum, b.step_grad_count, b.step_max_radii, self.stats.grad_accum,
                                self.stats.grad_count, self.stats.max_radii_2d)
        self.optimizer.zero_grad()                                       # :240-242
        self.optimizer.apply_gradients(BackwardOutput(b.dL_dpositions, b.dL_drotations, b.dL_dscales,
                                                      b.dL_dopacities, b.dL_dsh_coeffs, b.dL_dmeans_2d))
        self.optimizer.step()
        self.last_density_event = None
        if cfg.mcmc is not None:  # position noise every iteration, after the optimizer step (trainer.cpp:251)
            mcmc_inject_noise(self.model, step, cfg.mcmc)
            if cfg.mcmc_relocation:
                from .density import mcmc_relocate, mcmc_should_relocate
                if mcmc_should_relocate(step, cfg.mcmc):                 # :254-265; N constant, optimizer untouched
                    mcmc_relocate(self.model, step, cfg.mcmc, cfg.scene_extent, want_stats=False)
                    self.last_density_event = "relocate"
        elif cfg.densification is not None:
            # every rank holds the same statistics (all-reduced) and draws the same Philox numbers, so the
            # replicas take identical decisions and stay bit-identical without any extra exchange
            ctrl = self.stats
            if ctrl.should_densify(step):                                # :272-296
                res = ctrl.densify(self.model, step, optimizer=self.optimizer,
                                   carry_optimizer_state=cfg.carry_optimizer_state)
                self.last_density_event = res
                if self.model.num_gaussians() != b.n:
                    cam0 = self.cameras[0]
                    self.buffers = self._FrameBuffers(self.model.num_gaussians(), cam0.width, cam0.height,
                                                      int(self.model.sh_coeffs.shape[2]), self.model.positions.device)
            if ctrl.should_reset_opacity(step):                          # :299-302
                ctrl.reset_opacity(self.model)
        self.last_scalars = scal_sum / float(len(self.cameras))
        return self.last_scalars


class TargetUploader:
    """Host -> device upload of the per-view target images on a side stream, double-buffered, so the
    PCIe copy of view v+1 overlaps the rendering of view v (the reference's trainer loads and uploads
    the image synchronously every step: training/trainer.cpp:186-198)."""

    def __init__(self, height: int, width: int, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.bufs = [torch.empty((height, width, 3), dtype=torch.float32, device=self.device) for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.cur = None      # buffer handed out by get() and possibly still in use
        self.pending = None  # buffer being filled by prefetch()

    def prefetch(self, host_image: torch.Tensor) -> None:
        """Start the copy of a pinned host image into the buffer that is not in use."""
        s = 0 if self.cur is None else self.cur ^ 1
        self.stream.wait_event(self.consumed[s])  # the previous consumer of this buffer must be done
        with torch.cuda.stream(self.stream):
            self.bufs[s].copy_(host_image, non_blocking=True)
            self.ready[s].record(self.stream)
        self.pending = s

    def get(self) -> torch.Tensor:
        """The most recently prefetched image, usable on the current stream."""
        assert self.pending is not None, "prefetch() first"
        s = self.pending
        torch.cuda.current_stream(self.device).wait_event(self.ready[s])
        self.cur, self.pending = s, None
        return self.bufs[s]

    def release(self) -> None:
        """Mark the current buffer as consumed by everything enqueued so far on the current stream."""
        self.consumed[self.cur].record(torch.cuda.current_stream(self.device))
