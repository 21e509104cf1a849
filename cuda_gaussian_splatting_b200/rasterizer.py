"""Host-side mirror of the reference's rasterizer interface, on top of the C ABI.

Same names, argument meaning and error behaviour as /root/reference/src/rasterizer/*.hpp:
``render`` / ``render_backward`` (rasterizer.hpp:57-60, :88-93), the public stage functions
``project_gaussians`` (projection.hpp:39), ``sort_gaussians`` (sorting.hpp:41),
``rasterize_forward`` (forward.hpp:41), ``rasterize_backward`` (backward.hpp:39),
``project_backward`` (projection_backward.hpp:44), ``evaluate_sh_cuda`` (core/sh.hpp:29) and
``evaluate_sh_backward_cuda`` (core/sh_backward.hpp:25). torch is used for device memory and
streams only; all compute goes through libcugs_b200.so (sm_100a CUDA) — no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import CugsView

TILE = 16  # rasterizer/sorting.hpp:16


# ----------------------------------------------------------------------------------------------
# boundary types
# ----------------------------------------------------------------------------------------------
@dataclass
class CameraInfo:
    """core/types.hpp:78-109 (pinhole intrinsics + world->camera rotation/translation)."""

    width: int
    height: int
    fx: float
    fy: float
    cx: float
    cy: float
    rotation: np.ndarray = field(default_factory=lambda: np.eye(3, dtype=np.float32))
    translation: np.ndarray = field(default_factory=lambda: np.zeros(3, dtype=np.float32))

    def world_to_camera(self) -> np.ndarray:
        """Row-major 4x4 (types.hpp:103-108, projection.cu:227-233)."""
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = np.asarray(self.rotation, dtype=np.float32)
        m[:3, 3] = np.asarray(self.translation, dtype=np.float32)
        return m

    def camera_center(self) -> np.ndarray:
        """-R^T t in float32 (types.hpp:98-100)."""
        r = np.asarray(self.rotation, dtype=np.float32)
        t = np.asarray(self.translation, dtype=np.float32)
        out = np.zeros(3, dtype=np.float32)
        for i in range(3):  # same left-to-right float accumulation as the fixed-size product
            acc = np.float32(0.0)
            for k in range(3):
                acc = np.float32(acc + np.float32(-r[k, i]) * t[k])
            out[i] = acc
        return out

    def as_ref_list(self) -> list:
        """18 numbers consumed by oracle/ref_harness.cpp."""
        return ([float(self.width), float(self.height), float(self.fx), float(self.fy), float(self.cx),
                 float(self.cy)] + [float(x) for x in np.asarray(self.rotation, np.float32).reshape(-1)]
                + [float(x) for x in np.asarray(self.translation, np.float32).reshape(-1)])


@dataclass
class RenderSettings:
    """rasterizer.hpp:17-21."""

    background: Sequence[float] = (0.0, 0.0, 0.0)
    active_sh_degree: int = 3
    scale_modifier: float = 1.0


@dataclass
class GaussianModel:
    """core/gaussian.hpp:34-102 (SoA tensors, f32, one CUDA device)."""

    positions: torch.Tensor   # [N,3]
    sh_coeffs: torch.Tensor   # [N,3,C]
    opacities: torch.Tensor   # [N,1] logit
    rotations: torch.Tensor   # [N,4] wxyz
    scales: torch.Tensor      # [N,3] log

    def num_gaussians(self) -> int:
        return int(self.positions.shape[0])

    def max_sh_degree(self) -> int:
        return int(math.sqrt(float(self.sh_coeffs.shape[2]))) - 1 if self.sh_coeffs.dim() == 3 else 0

    def is_valid(self) -> bool:
        n = self.positions.shape[0]
        ok = (self.positions.dim() == 2 and self.positions.shape[1] == 3
              and self.sh_coeffs.dim() == 3 and self.sh_coeffs.shape[0] == n and self.sh_coeffs.shape[1] == 3
              and self.opacities.dim() == 2 and tuple(self.opacities.shape) == (n, 1)
              and self.rotations.dim() == 2 and tuple(self.rotations.shape) == (n, 4)
              and self.scales.dim() == 2 and tuple(self.scales.shape) == (n, 3))
        dev = self.positions.device
        return bool(ok and all(t.device == dev for t in
                               (self.sh_coeffs, self.opacities, self.rotations, self.scales)))


@dataclass
class ProjectionOutput:  # projection.hpp:15-24
    means_2d: torch.Tensor
    depths: torch.Tensor
    cov_2d_inv: torch.Tensor
    radii: torch.Tensor
    tiles_touched: torch.Tensor
    rgb: torch.Tensor
    opacities_act: torch.Tensor


@dataclass
class SortingOutput:  # sorting.hpp:19-24
    gaussian_keys_sorted: torch.Tensor
    gaussian_values_sorted: torch.Tensor
    tile_ranges: torch.Tensor
    total_pairs: int


@dataclass
class ForwardOutput:  # forward.hpp:11-15
    color: torch.Tensor
    final_T: torch.Tensor
    n_contrib: torch.Tensor


@dataclass
class RasterizeBackwardOutput:  # backward.hpp:13-18
    dL_drgb: torch.Tensor
    dL_dopacity_act: torch.Tensor
    dL_dmeans_2d: torch.Tensor
    dL_dcov_2d_inv: torch.Tensor


@dataclass
class ProjectionBackwardOutput:  # projection_backward.hpp:15-21
    dL_dpositions: torch.Tensor
    dL_drotations: torch.Tensor
    dL_dscales: torch.Tensor
    dL_dopacities: torch.Tensor
    dL_dsh_coeffs: torch.Tensor


@dataclass
class RenderOutput:  # rasterizer.hpp:27-46
    color: torch.Tensor
    final_T: torch.Tensor
    n_contrib: torch.Tensor
    means_2d: torch.Tensor
    depths: torch.Tensor
    cov_2d_inv: torch.Tensor
    radii: torch.Tensor
    rgb: torch.Tensor
    opacities_act: torch.Tensor
    gaussian_indices: torch.Tensor
    tile_ranges: torch.Tensor
    # private: the frame workspace (packed blend records) reused by render_backward
    _workspace: Optional[torch.Tensor] = None

    def as_list(self):
        return [self.color, self.final_T, self.n_contrib, self.means_2d, self.depths, self.cov_2d_inv,
                self.radii, self.rgb, self.opacities_act, self.gaussian_indices, self.tile_ranges]


@dataclass
class BackwardOutput:  # rasterizer.hpp:65-72
    dL_dpositions: torch.Tensor
    dL_drotations: torch.Tensor
    dL_dscales: torch.Tensor
    dL_dopacities: torch.Tensor
    dL_dsh_coeffs: torch.Tensor
    dL_dmeans_2d: torch.Tensor


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def _check(cond: bool, msg: str) -> None:
    if not cond:  # TORCH_CHECK -> c10::Error -> RuntimeError in Python
        raise RuntimeError(msg)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else (t.data_ptr() if t.numel() > 0 else None)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous().to(torch.float32)  # projection.cu:240-243


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def make_view(camera: CameraInfo, settings: RenderSettings, active_degree: int, num_coeffs: int) -> CugsView:
    v = CugsView()
    v.width, v.height = int(camera.width), int(camera.height)
    v.fx, v.fy, v.cx, v.cy = float(camera.fx), float(camera.fy), float(camera.cx), float(camera.cy)
    w2c = camera.world_to_camera().reshape(-1)
    for i in range(16):
        v.view[i] = float(w2c[i])
    c = camera.camera_center()
    for i in range(3):
        v.cam_center[i] = float(c[i])
        v.bg[i] = float(settings.background[i])
    v.active_sh_degree = int(active_degree)
    v.num_coeffs = int(num_coeffs)
    v.scale_modifier = float(settings.scale_modifier)
    return v


def _lib_and_handle(dev: torch.device):
    _check(dev.type == "cuda", "tensors must be on a CUDA device (there is no CPU path)")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return _lib.load_library(), _lib.handle(idx)


def num_tiles(width: int, height: int) -> int:
    return ((width + TILE - 1) // TILE) * ((height + TILE - 1) // TILE)


# ----------------------------------------------------------------------------------------------
# stage functions
# ----------------------------------------------------------------------------------------------
def project_gaussians(positions, rotations, scales, opacities, sh_coeffs, camera: CameraInfo,
                      active_sh_degree: int, scale_modifier: float = 1.0,
                      _packed: Optional[torch.Tensor] = None,
                      _depth_minmax: Optional[torch.Tensor] = None) -> ProjectionOutput:
    """projection.hpp:39-47 / projection.cu:195-289."""
    _check(positions.is_cuda, "positions must be on CUDA")
    _check(positions.dim() == 2 and positions.shape[1] == 3, "positions must be [N,3]")
    _check(0 <= active_sh_degree <= 3, f"SH degree must be 0..3, got {active_sh_degree}")
    _check(sh_coeffs.dim() == 3 and sh_coeffs.shape[1] == 3, "sh_coeffs must be [N, 3, C]")
    _check(sh_coeffs.shape[2] >= (active_sh_degree + 1) ** 2,
           f"Need at least {(active_sh_degree + 1) ** 2} coefficients for degree {active_sh_degree}")
    dev = positions.device
    n = positions.shape[0]
    f = dict(dtype=torch.float32, device=dev)
    i = dict(dtype=torch.int32, device=dev)
    out = ProjectionOutput(torch.empty((n, 2), **f), torch.empty((n,), **f), torch.empty((n, 3), **f),
                           torch.empty((n,), **i), torch.empty((n,), **i), torch.empty((n, 3), **f),
                           torch.empty((n,), **f))
    if n == 0:
        return out
    lib, h = _lib_and_handle(dev)
    pos, rot, scl, opa, sh = map(_f32c, (positions, rotations, scales, opacities, sh_coeffs))
    settings = RenderSettings(scale_modifier=scale_modifier)
    v = make_view(camera, settings, active_sh_degree, sh.shape[2])
    st = lib.cugs_b200_preprocess_fwd(h, _stream(dev), n, C.byref(v), _ptr(pos), _ptr(rot), _ptr(scl),
                                      _ptr(opa), _ptr(sh), _ptr(out.means_2d), _ptr(out.depths),
                                      _ptr(out.cov_2d_inv), _ptr(out.radii), _ptr(out.tiles_touched),
                                      _ptr(out.rgb), _ptr(out.opacities_act), _ptr(_packed),
                                      _ptr(_depth_minmax))
    _lib.check(h, st, "cugs_b200_preprocess_fwd")
    return out


def sort_gaussians(means_2d, depths, radii, tiles_touched, img_w: int, img_h: int,
                   depth_bits: int = 32, tile_bits: Optional[int] = None) -> SortingOutput:
    """sorting.hpp:41-46 / sorting.cu:115-227. depth_bits/tile_bits select the key bits that are
    sorted (default: every bit that can differ for tile|depth keys)."""
    _check(means_2d.is_cuda, "means_2d must be on CUDA")
    dev = means_2d.device
    n = means_2d.shape[0]
    nt = num_tiles(img_w, img_h)
    i32 = dict(dtype=torch.int32, device=dev)
    if n == 0:
        return SortingOutput(torch.empty((0,), dtype=torch.int64, device=dev), torch.empty((0,), **i32),
                             torch.zeros((nt, 2), **i32), 0)
    lib, h = _lib_and_handle(dev)
    s = _stream(dev)
    tiles = tiles_touched.contiguous().to(torch.int32)
    offsets = torch.empty((n,), **i32)
    tmp = torch.empty((lib.cugs_b200_scan_temp_bytes(n),), dtype=torch.uint8, device=dev)
    total = C.c_int64(0)
    st = lib.cugs_b200_scan(h, s, n, _ptr(tiles), _ptr(offsets), None, C.byref(total), _ptr(tmp), tmp.numel())
    _lib.check(h, st, "cugs_b200_scan")
    p = int(total.value)
    if p == 0:
        return SortingOutput(torch.empty((0,), dtype=torch.int64, device=dev), torch.empty((0,), **i32),
                             torch.zeros((nt, 2), **i32), 0)
    keys = torch.empty((p,), dtype=torch.int64, device=dev)
    vals = torch.empty((p,), **i32)
    m2d, dep, rad = means_2d.contiguous(), depths.contiguous(), radii.contiguous()
    st = lib.cugs_b200_duplicate_with_keys(h, s, n, img_w, img_h, _ptr(m2d), _ptr(dep), _ptr(rad), _ptr(tiles),
                                           _ptr(offsets), p, _ptr(keys), _ptr(vals))
    _lib.check(h, st, "cugs_b200_duplicate_with_keys")
    keys_sorted = torch.empty_like(keys)
    vals_sorted = torch.empty_like(vals)
    stmp = torch.empty((lib.cugs_b200_sort_temp_bytes(p),), dtype=torch.uint8, device=dev)
    if tile_bits is None:
        tile_bits = max(0, math.ceil(math.log2(nt))) if nt > 1 else 0
    st = lib.cugs_b200_sort_pairs(h, s, p, int(depth_bits), int(tile_bits), _ptr(keys), _ptr(vals),
                                  _ptr(keys_sorted), _ptr(vals_sorted), _ptr(stmp), stmp.numel())
    _lib.check(h, st, "cugs_b200_sort_pairs")
    ranges = torch.empty((nt, 2), **i32)
    st = lib.cugs_b200_tile_ranges(h, s, p, _ptr(keys_sorted), nt, _ptr(ranges))
    _lib.check(h, st, "cugs_b200_tile_ranges")
    return SortingOutput(keys_sorted, vals_sorted, ranges, p)


def rasterize_forward(means_2d, cov_2d_inv, rgb, opacities, tile_ranges, gaussian_indices, img_w: int,
                      img_h: int, background: Sequence[float]) -> ForwardOutput:
    """forward.hpp:41-49 / forward.cu:180-240."""
    _check(means_2d.is_cuda, "means_2d must be on CUDA")
    dev = means_2d.device
    f = dict(dtype=torch.float32, device=dev)
    color = torch.empty((img_h, img_w, 3), **f)
    final_T = torch.empty((img_h, img_w), **f)
    n_contrib = torch.empty((img_h, img_w), dtype=torch.int32, device=dev)
    if img_w == 0 or img_h == 0:
        return ForwardOutput(color, final_T, n_contrib)
    lib, h = _lib_and_handle(dev)
    cam = CameraInfo(img_w, img_h, 1.0, 1.0, 0.0, 0.0)
    v = make_view(cam, RenderSettings(background=background), 0, 1)
    args = [t.contiguous() for t in (tile_ranges, gaussian_indices, means_2d, cov_2d_inv, rgb, opacities)]
    st = lib.cugs_b200_blend_fwd(h, _stream(dev), C.byref(v), *[_ptr(a) for a in args], None, _ptr(color),
                                 _ptr(final_T), _ptr(n_contrib))
    _lib.check(h, st, "cugs_b200_blend_fwd")
    return ForwardOutput(color, final_T, n_contrib)


def rasterize_backward(dL_dcolor, means_2d, cov_2d_inv, rgb, opacities, tile_ranges, gaussian_indices,
                       final_T, n_contrib, img_w: int, img_h: int, background: Sequence[float],
                       n_gaussians: int) -> RasterizeBackwardOutput:
    """backward.hpp:39-51 / backward.cu:239-306."""
    _check(dL_dcolor.is_cuda, "dL_dcolor must be on CUDA")
    dev = dL_dcolor.device
    n = int(n_gaussians)
    f = dict(dtype=torch.float32, device=dev)
    out = RasterizeBackwardOutput(torch.zeros((n, 3), **f), torch.zeros((n,), **f), torch.zeros((n, 2), **f),
                                  torch.zeros((n, 3), **f))
    if n == 0 or img_w == 0 or img_h == 0:
        return out
    lib, h = _lib_and_handle(dev)
    cam = CameraInfo(img_w, img_h, 1.0, 1.0, 0.0, 0.0)
    v = make_view(cam, RenderSettings(background=background), 0, 1)
    acc = torch.empty((n, 12), **f)
    args = [t.contiguous() for t in (tile_ranges, gaussian_indices, means_2d, cov_2d_inv, rgb, opacities)]
    st = lib.cugs_b200_blend_bwd(h, _stream(dev), n, C.byref(v), *[_ptr(a) for a in args], None,
                                 _ptr(dL_dcolor.contiguous()), _ptr(final_T.contiguous()),
                                 _ptr(n_contrib.contiguous()), _ptr(out.dL_drgb), _ptr(out.dL_dopacity_act),
                                 _ptr(out.dL_dmeans_2d), _ptr(out.dL_dcov_2d_inv), _ptr(acc))
    _lib.check(h, st, "cugs_b200_blend_bwd")
    return out


def project_backward(dL_dmeans_2d, dL_dcov_2d_inv, dL_drgb, dL_dopacity_act, positions, rotations, scales,
                     opacities, sh_coeffs, radii, camera: CameraInfo, active_sh_degree: int,
                     scale_modifier: float = 1.0, rgb: Optional[torch.Tensor] = None,
                     stats: Optional[Sequence[torch.Tensor]] = None) -> ProjectionBackwardOutput:
    """projection_backward.hpp:44-57 / projection_backward.cu:253-344. ``rgb`` is the forward colour
    (its sign is the ReLU gate); when omitted it is recomputed with project_gaussians."""
    _check(positions.is_cuda, "positions must be on CUDA")
    dev = positions.device
    n = positions.shape[0]
    f = dict(dtype=torch.float32, device=dev)
    pos, rot, scl, opa, sh = map(_f32c, (positions, rotations, scales, opacities, sh_coeffs))
    out = ProjectionBackwardOutput(torch.empty((n, 3), **f), torch.empty((n, 4), **f), torch.empty((n, 3), **f),
                                   torch.empty((n, 1), **f), torch.empty_like(sh))
    if n == 0:
        return out
    lib, h = _lib_and_handle(dev)
    if rgb is None:
        rgb = project_gaussians(pos, rot, scl, opa, sh, camera, active_sh_degree, scale_modifier).rgb
    v = make_view(camera, RenderSettings(scale_modifier=scale_modifier), active_sh_degree, sh.shape[2])
    sp = [None, None, None] if stats is None else [_ptr(t) for t in stats]
    st = lib.cugs_b200_preprocess_bwd(h, _stream(dev), n, C.byref(v), _ptr(pos), _ptr(rot), _ptr(scl), _ptr(opa),
                                      _ptr(sh), _ptr(radii.contiguous()), _ptr(rgb.contiguous()),
                                      _ptr(dL_dmeans_2d.contiguous()), _ptr(dL_dcov_2d_inv.contiguous()),
                                      _ptr(dL_drgb.contiguous()), _ptr(dL_dopacity_act.contiguous()),
                                      _ptr(out.dL_dpositions), _ptr(out.dL_drotations), _ptr(out.dL_dscales),
                                      _ptr(out.dL_dopacities), _ptr(out.dL_dsh_coeffs), *sp)
    _lib.check(h, st, "cugs_b200_preprocess_bwd")
    return out


def evaluate_sh_cuda(degree: int, sh_coeffs, directions) -> torch.Tensor:
    """core/sh.hpp:29 / sh.cu:81-123 (same validation, no clamp)."""
    _check(0 <= degree <= 3, f"SH degree must be 0..3, got {degree}")
    _check(sh_coeffs.is_cuda, "sh_coeffs must be on CUDA device")
    _check(directions.is_cuda, "directions must be on CUDA device")
    _check(sh_coeffs.dim() == 3 and sh_coeffs.shape[1] == 3, "sh_coeffs must be [N, 3, C]")
    _check(directions.dim() == 2 and directions.shape[1] == 3, "directions must be [N, 3]")
    _check(sh_coeffs.shape[0] == directions.shape[0], "Batch size mismatch")
    _check(sh_coeffs.shape[2] >= (degree + 1) ** 2, f"Need at least {(degree + 1) ** 2} coefficients for degree {degree}")
    dev = sh_coeffs.device
    n = sh_coeffs.shape[0]
    out = torch.empty((n, 3), dtype=torch.float32, device=dev)
    if n == 0:
        return out
    lib, h = _lib_and_handle(dev)
    sh, d = _f32c(sh_coeffs), _f32c(directions)
    st = lib.cugs_b200_sh_forward(h, _stream(dev), n, degree, sh.shape[2], _ptr(sh), _ptr(d), _ptr(out))
    _lib.check(h, st, "cugs_b200_sh_forward")
    return out


def evaluate_sh_backward_cuda(degree: int, sh_coeffs, directions, dL_dcolor) -> torch.Tensor:
    """core/sh_backward.hpp:25 / sh_backward.cu:114-156."""
    _check(0 <= degree <= 3, f"SH degree must be 0..3, got {degree}")
    _check(sh_coeffs.is_cuda and directions.is_cuda and dL_dcolor.is_cuda, "inputs must be on CUDA device")
    _check(sh_coeffs.dim() == 3 and sh_coeffs.shape[1] == 3, "sh_coeffs must be [N, 3, C]")
    _check(directions.dim() == 2 and directions.shape[1] == 3, "directions must be [N, 3]")
    _check(dL_dcolor.dim() == 2 and dL_dcolor.shape[1] == 3, "dL_dcolor must be [N, 3]")
    dev = sh_coeffs.device
    n = sh_coeffs.shape[0]
    sh, d, g = _f32c(sh_coeffs), _f32c(directions), _f32c(dL_dcolor)
    out = torch.empty_like(sh)
    if n == 0:
        return out
    lib, h = _lib_and_handle(dev)
    st = lib.cugs_b200_sh_backward(h, _stream(dev), n, degree, sh.shape[2], _ptr(sh), _ptr(d), _ptr(g), _ptr(out))
    _lib.check(h, st, "cugs_b200_sh_backward")
    return out


# ----------------------------------------------------------------------------------------------
# render / render_backward
# ----------------------------------------------------------------------------------------------
class FrameBuffers:
    """Reusable device buffers for one in-flight frame of a fixed (N, W, H): everything render()
    would otherwise allocate per call (the reference allocates ~45 tensors per frame)."""

    def __init__(self, n: int, width: int, height: int, num_coeffs: int, device, p_capacity: int = 0,
                 share_grads_with: Optional["FrameBuffers"] = None, symmetric: bool = False):
        f = dict(dtype=torch.float32, device=device)
        i = dict(dtype=torch.int32, device=device)
        self.n, self.width, self.height = n, width, height
        self.means_2d = torch.empty((n, 2), **f)
        self.depths = torch.empty((n,), **f)
        self.cov_2d_inv = torch.empty((n, 3), **f)
        self.radii = torch.empty((n,), **i)
        self.rgb = torch.empty((n, 3), **f)
        self.opacities_act = torch.empty((n,), **f)
        self.color = torch.empty((height, width, 3), **f)
        self.final_T = torch.empty((height, width), **f)
        self.n_contrib = torch.empty((height, width), **i)
        self.tile_ranges = torch.empty((num_tiles(width, height), 2), **i)
        self.gaussian_indices = torch.empty((max(p_capacity, 1),), **i)
        # frame arena: the N-sized head (packed blend records, binning state, 2-D gradient accumulator) has a
        # fixed size; the P-sized pair scratch is a separate block that grows without touching the head
        lib = _lib.load_library()
        self.workspace = torch.empty((lib.cugs_b200_render_workspace_bytes(n, 0),), dtype=torch.uint8, device=device)
        self.pair_scratch = torch.empty((1,), dtype=torch.uint8, device=device)
        self.p_capacity = 0
        self.ensure_capacity(max(p_capacity, 1))
        # gradients: ONE contiguous arena laid out as the five Adam groups (fused_adam.cu:94-97:
        # positions, sh_coeffs, opacities, scales, rotations) followed by the two additive
        # densification statistics of the step, so that view-parallel training needs exactly one
        # all-reduce(sum) per step. Segment starts are 256-byte aligned (float4 accesses).
        # (a second in-flight frame of the same step shares the arena of the first: share_grads_with)
        from .parallel import arena_layout
        layout, total = arena_layout(n, num_coeffs)
        if share_grads_with is not None:
            _check(share_grads_with.n == n and share_grads_with.grad_arena.numel() == total,
                   "share_grads_with: incompatible FrameBuffers")
        # symmetric = True (view-parallel training, parallel.P2PExchange): the arena and the MAX buffer are
        # allocated in torch symmetric memory, i.e. at the same offsets on every rank and mappable by the peers
        self.symmetric = bool(symmetric) if share_grads_with is None else share_grads_with.symmetric
        if share_grads_with is not None:
            self.grad_arena = share_grads_with.grad_arena
        elif symmetric:
            import torch.distributed._symmetric_memory as symm_mem
            self.grad_arena = symm_mem.empty(total, dtype=torch.float32, device=torch.device(device))
            self.grad_arena.zero_()
        else:
            self.grad_arena = torch.zeros((total,), **f)
        seg = lambda nm: self.grad_arena[layout[nm][0]:layout[nm][0] + layout[nm][1]]
        self.dL_dpositions = seg("positions").view(n, 3)
        self.dL_dsh_coeffs = seg("sh_coeffs").view(n, 3, num_coeffs)
        self.dL_dopacities = seg("opacities").view(n, 1)
        self.dL_dscales = seg("scales").view(n, 3)
        self.dL_drotations = seg("rotations").view(n, 4)
        self.step_grad_accum = seg("grad_accum")   # sum over the step's views of ||dL/dmeans_2d|| (visible only)
        self.step_grad_count = seg("grad_count")   # number of the step's views in which the Gaussian was visible
        # quantities that need a MAX reduction live in one int32 buffer [touch mask | max_radii bits]
        # (non-negative floats order like their bit patterns, so one int32 MAX all-reduce serves both)
        if share_grads_with is not None:
            self.max_buf = share_grads_with.max_buf
        elif symmetric:
            import torch.distributed._symmetric_memory as symm_mem
            self.max_buf = symm_mem.empty(2 * max(n, 1), dtype=torch.int32, device=torch.device(device))
            self.max_buf.zero_()
        else:
            self.max_buf = torch.zeros((2 * max(n, 1),), **i)
        self.touch_mask = self.max_buf[:n]                       # 1 where some view gave a non-zero gradient
        self.step_max_radii = self.max_buf[n:2 * n].view(torch.float32)
        self.grad_compact = None                                  # lazily allocated by parallel.sparse_allreduce_step
        self.touch_offsets = None
        self.dL_dmeans_2d = torch.empty((n, 2), **f)
        # frames rendered without the host round trip (render(..., sync=False)): {P, overflowed} on the device
        # and its pinned host copy, refreshed by fetch_status()
        self.status_dev = torch.zeros((2,), dtype=torch.int64, device=device)
        self.status_host = torch.zeros((2,), dtype=torch.int64).pin_memory() if torch.device(device).type == "cuda" \
            else torch.zeros((2,), dtype=torch.int64)

    def fetch_status(self, non_blocking: bool = True) -> None:
        """Queue the device->host copy of {P, overflowed} of the last sync=False frame on the current stream."""
        self.status_host.copy_(self.status_dev, non_blocking=non_blocking)

    def last_pairs(self):
        """(P, overflowed) of the last fetched frame; valid once the stream that ran fetch_status() is synchronised."""
        return int(self.status_host[0]), bool(int(self.status_host[1]))

    def ensure_capacity(self, p: int) -> None:
        if p <= self.p_capacity:
            return
        lib = _lib.load_library()
        cap = int(p * 1.25) + 1024
        dev = self.workspace.device
        self.pair_scratch = torch.empty((lib.cugs_b200_render_pair_scratch_bytes(cap),), dtype=torch.uint8, device=dev)
        self.gaussian_indices = torch.empty((cap,), dtype=torch.int32, device=dev)
        self.p_capacity = cap


def render(model: GaussianModel, camera: CameraInfo, settings: RenderSettings,
           buffers: Optional[FrameBuffers] = None, sync: bool = True) -> RenderOutput:
    """rasterizer.hpp:57-60 / rasterizer.cpp:22-113.

    ``sync=False`` (needs ``buffers`` whose pair capacity was established by an earlier frame or by
    ``ensure_capacity``): nothing blocks -- the pair count P stays on the device, every launch is sized on the
    buffers' capacity (cugs_b200_render_forward), ``gaussian_indices`` is returned at full capacity (entries
    beyond P are not written; the kernels never read them) and ``buffers.status_dev`` receives {P, overflowed}.
    Call ``buffers.fetch_status()`` and check ``buffers.last_pairs()`` after a synchronisation point (e.g. once per
    step): an overflowed frame must be re-rendered (``sync=True`` grows the capacity)."""
    _check(model.is_valid(), "GaussianModel is not valid")
    _check(model.positions.is_cuda, "GaussianModel must be on CUDA device")
    dev = model.positions.device
    n = model.num_gaussians()
    f = dict(dtype=torch.float32, device=dev)
    i = dict(dtype=torch.int32, device=dev)
    H, W = camera.height, camera.width
    if n == 0:  # rasterizer.cpp:36-55
        color = torch.zeros((H, W, 3), **f)
        for ch in range(3):
            color[..., ch] = settings.background[ch]
        return RenderOutput(color, torch.ones((H, W), **f), torch.zeros((H, W), **i), torch.empty((0, 2), **f),
                            torch.empty((0,), **f), torch.empty((0, 3), **f), torch.empty((0,), **i),
                            torch.empty((0, 3), **f), torch.empty((0,), **f), torch.empty((0,), **i),
                            torch.empty((0, 2), **i))
    lib, h = _lib_and_handle(dev)
    s = _stream(dev)
    active = min(settings.active_sh_degree, model.max_sh_degree())  # rasterizer.cpp:60
    pos, rot, scl, opa, sh = map(_f32c, (model.positions, model.rotations, model.scales, model.opacities,
                                         model.sh_coeffs))
    v = make_view(camera, settings, active, sh.shape[2])
    b = buffers if buffers is not None else FrameBuffers(n, W, H, sh.shape[2], dev)
    _check(b.n == n and b.width == W and b.height == H, "FrameBuffers do not match the model / camera")
    if not sync:
        _check(buffers is not None and b.p_capacity > 1, "sync=False needs FrameBuffers with an established pair capacity")
        st = lib.cugs_b200_render_forward(
            h, s, n, b.p_capacity, C.byref(v), _ptr(pos), _ptr(rot), _ptr(scl), _ptr(opa), _ptr(sh), _ptr(b.means_2d),
            _ptr(b.depths), _ptr(b.cov_2d_inv), _ptr(b.radii), _ptr(b.rgb), _ptr(b.opacities_act),
            b.gaussian_indices.data_ptr(), _ptr(b.tile_ranges), _ptr(b.color), _ptr(b.final_T), _ptr(b.n_contrib),
            _ptr(b.workspace), b.workspace.numel(), _ptr(b.pair_scratch), b.pair_scratch.numel(),
            b.status_dev.data_ptr())
        _lib.check(h, st, "cugs_b200_render_forward")
        return RenderOutput(b.color, b.final_T, b.n_contrib, b.means_2d, b.depths, b.cov_2d_inv, b.radii, b.rgb,
                            b.opacities_act, b.gaussian_indices[:b.p_capacity], b.tile_ranges, _workspace=b.workspace)
    p = C.c_int64(0)
    st = lib.cugs_b200_render_plan(h, s, n, C.byref(v), _ptr(pos), _ptr(rot), _ptr(scl), _ptr(opa), _ptr(sh),
                                   _ptr(b.means_2d), _ptr(b.depths), _ptr(b.cov_2d_inv), _ptr(b.radii),
                                   _ptr(b.rgb), _ptr(b.opacities_act), _ptr(b.workspace), b.workspace.numel(),
                                   C.byref(p))
    _lib.check(h, st, "cugs_b200_render_plan")
    P = int(p.value)
    b.ensure_capacity(P)
    st = lib.cugs_b200_render_finish(h, s, n, P, C.byref(v), _ptr(b.means_2d), _ptr(b.depths), _ptr(b.cov_2d_inv),
                                     _ptr(b.radii), _ptr(b.rgb), _ptr(b.opacities_act),
                                     b.gaussian_indices.data_ptr(), _ptr(b.tile_ranges), _ptr(b.color),
                                     _ptr(b.final_T), _ptr(b.n_contrib), _ptr(b.workspace), b.workspace.numel(),
                                     _ptr(b.pair_scratch), b.pair_scratch.numel())
    _lib.check(h, st, "cugs_b200_render_finish")
    return RenderOutput(b.color, b.final_T, b.n_contrib, b.means_2d, b.depths, b.cov_2d_inv, b.radii, b.rgb,
                        b.opacities_act, b.gaussian_indices[:P], b.tile_ranges, _workspace=b.workspace)


def count_evaluations(render_out: RenderOutput, camera: CameraInfo) -> dict:
    """MEASUREMENT ONLY: the (pixel, Gaussian) evaluations the reference's traversal performs on this frame,
    split into alpha-rejected and contributing ones, forward (rasterizer/forward.cu:121-157) and backward
    (rasterizer/backward.cu:117-145) -- the work units of the blend kernels' FP32 / MUFU roofline.
    Blocks (one device->host copy of four counters)."""
    r = render_out
    dev = r.color.device
    lib, h = _lib_and_handle(dev)
    v = make_view(camera, RenderSettings(), 0, 1)
    counts = torch.zeros((4,), dtype=torch.int64, device=dev)
    st = lib.cugs_b200_count_evaluations(h, _stream(dev), C.byref(v), _ptr(r.tile_ranges),
                                         r.gaussian_indices.data_ptr(), _ptr(r.means_2d), _ptr(r.cov_2d_inv),
                                         _ptr(r.opacities_act), _ptr(r.n_contrib), counts.data_ptr())
    _lib.check(h, st, "cugs_b200_count_evaluations")
    c = [int(x) for x in counts.cpu()]
    return {"fwd_rejected": c[0], "fwd_contributing": c[1], "bwd_rejected": c[2], "bwd_contributing": c[3],
            "E_fwd": c[0] + c[1], "E_bwd": c[2] + c[3]}


class ImageBuffers:
    """Device buffers of a render-only frame (no backward-only arrays): what render_image reuses."""

    def __init__(self, n: int, width: int, height: int, device):
        f = dict(dtype=torch.float32, device=device)
        i = dict(dtype=torch.int32, device=device)
        self.n, self.width, self.height = n, width, height
        self.means_2d = torch.empty((n, 2), **f)
        self.radii = torch.empty((n,), **i)
        self.color = torch.empty((height, width, 3), **f)
        self.final_T = torch.empty((height, width), **f)
        self.n_contrib = torch.empty((height, width), **i)
        self.tile_ranges = torch.empty((num_tiles(width, height), 2), **i)
        lib = _lib.load_library()
        self.workspace = torch.empty((lib.cugs_b200_render_workspace_bytes(n, 0),), dtype=torch.uint8, device=device)
        self.pair_scratch = torch.empty((1,), dtype=torch.uint8, device=device)
        self.gaussian_indices = torch.empty((1,), **i)
        self.p_capacity = 0

    ensure_capacity = FrameBuffers.ensure_capacity


def render_image(model: GaussianModel, camera: CameraInfo, settings: RenderSettings,
                 buffers: Optional[ImageBuffers] = None):
    """Render-only entry point for the callers that never run the backward pass -- evaluate()
    (training/metrics.cpp:131) and Viewer::render_frame (viewer/viewer.cpp:645-669) consume only
    ``color``, ``final_T`` and ``n_contrib``. Same kernels and same pixels as render(); depths,
    cov_2d_inv, rgb and opacities_act are not materialised. Returns (color, final_T, n_contrib)."""
    _check(model.is_valid(), "GaussianModel is not valid")
    _check(model.positions.is_cuda, "GaussianModel must be on CUDA device")
    n = model.num_gaussians()
    if n == 0:
        out = render(model, camera, settings)
        return out.color, out.final_T, out.n_contrib
    dev = model.positions.device
    lib, h = _lib_and_handle(dev)
    s = _stream(dev)
    active = min(settings.active_sh_degree, model.max_sh_degree())
    pos, rot, scl, opa, sh = map(_f32c, (model.positions, model.rotations, model.scales, model.opacities,
                                         model.sh_coeffs))
    v = make_view(camera, settings, active, sh.shape[2])
    b = buffers if buffers is not None else ImageBuffers(n, camera.width, camera.height, dev)
    _check(b.n == n and b.width == camera.width and b.height == camera.height,
           "ImageBuffers do not match the model / camera")
    p = C.c_int64(0)
    st = lib.cugs_b200_render_plan(h, s, n, C.byref(v), _ptr(pos), _ptr(rot), _ptr(scl), _ptr(opa), _ptr(sh),
                                   _ptr(b.means_2d), None, None, _ptr(b.radii), None, None, _ptr(b.workspace),
                                   b.workspace.numel(), C.byref(p))
    _lib.check(h, st, "cugs_b200_render_plan")
    P = int(p.value)
    b.ensure_capacity(P)
    st = lib.cugs_b200_render_finish(h, s, n, P, C.byref(v), _ptr(b.means_2d), None, None, _ptr(b.radii), None, None,
                                     b.gaussian_indices.data_ptr(), _ptr(b.tile_ranges), _ptr(b.color),
                                     _ptr(b.final_T), _ptr(b.n_contrib), _ptr(b.workspace), b.workspace.numel(),
                                     _ptr(b.pair_scratch), b.pair_scratch.numel())
    _lib.check(h, st, "cugs_b200_render_finish")
    return b.color, b.final_T, b.n_contrib


def render_backward(dL_dcolor: torch.Tensor, render_out: RenderOutput, model: GaussianModel,
                    camera: CameraInfo, settings: RenderSettings, buffers: Optional[FrameBuffers] = None,
                    stats: Optional[Sequence[torch.Tensor]] = None, accumulate: bool = False,
                    touch_mask: Optional[torch.Tensor] = None, sparse_rows: bool = False) -> BackwardOutput:
    """rasterizer.hpp:88-93 / rasterizer.cpp:115-186. ``stats`` = (grad_accum, grad_count,
    max_radii) fuses DensificationController::accumulate_gradients into the same launch;
    ``accumulate`` adds the parameter gradients to ``buffers`` instead of overwriting them
    (gradient of a batch of views); ``touch_mask`` ([N] int32) records which Gaussians received a
    non-zero gradient (input of the sparse gradient exchange, parallel.sparse_allreduce_step);
    ``sparse_rows`` additionally skips the gradient rows that are not touched, relying on the invariant
    that rows with mask 0 are zero (FrameBuffers allocates zeros; use the buffers' own touch_mask and
    do not write the gradient tensors by other means in between)."""
    _check(dL_dcolor.is_cuda, "dL_dcolor must be on CUDA device")
    _check(dL_dcolor.dim() == 3 and dL_dcolor.shape[2] == 3, "dL_dcolor must be [H, W, 3]")
    dev = dL_dcolor.device
    n = model.num_gaussians()
    f = dict(dtype=torch.float32, device=dev)
    if n == 0:  # rasterizer.cpp:130-139
        return BackwardOutput(torch.zeros((0, 3), **f), torch.zeros((0, 4), **f), torch.zeros((0, 3), **f),
                              torch.zeros((0, 1), **f), torch.zeros_like(model.sh_coeffs), torch.zeros((0, 2), **f))
    lib, h = _lib_and_handle(dev)
    active = min(settings.active_sh_degree, model.max_sh_degree())
    pos, rot, scl, opa, sh = map(_f32c, (model.positions, model.rotations, model.scales, model.opacities,
                                         model.sh_coeffs))
    v = make_view(camera, settings, active, sh.shape[2])
    ws = render_out._workspace
    _check(ws is not None, "render_out does not come from render() of this library")
    _check(not accumulate or buffers is not None, "accumulate=True needs persistent FrameBuffers")
    if buffers is not None:
        g = (buffers.dL_dpositions, buffers.dL_drotations, buffers.dL_dscales, buffers.dL_dopacities,
             buffers.dL_dsh_coeffs, buffers.dL_dmeans_2d)
    else:
        g = (torch.empty((n, 3), **f), torch.empty((n, 4), **f), torch.empty((n, 3), **f),
             torch.empty((n, 1), **f), torch.empty_like(sh), torch.empty((n, 2), **f))
    sp = [None, None, None] if stats is None else [_ptr(t) for t in stats]
    r = render_out
    _check(not sparse_rows or (touch_mask is not None and buffers is not None),
           "sparse_rows=True needs persistent FrameBuffers and their touch_mask")
    flags = (1 if accumulate else 0) | (2 if sparse_rows else 0)  # CUGS_BWD_ACCUMULATE | CUGS_BWD_SPARSE_ROWS
    st = lib.cugs_b200_render_backward(
        h, _stream(dev), n, C.byref(v), _ptr(pos), _ptr(rot), _ptr(scl), _ptr(opa), _ptr(sh), _ptr(r.means_2d),
        _ptr(r.cov_2d_inv), _ptr(r.radii), _ptr(r.rgb), _ptr(r.opacities_act), _ptr(r.gaussian_indices),
        _ptr(r.tile_ranges), _ptr(r.final_T), _ptr(r.n_contrib), _ptr(dL_dcolor.contiguous()),
        *[_ptr(t) for t in g], *sp, _ptr(touch_mask), flags, _ptr(ws), ws.numel())
    _lib.check(h, st, "cugs_b200_render_backward")
    return BackwardOutput(*g)
