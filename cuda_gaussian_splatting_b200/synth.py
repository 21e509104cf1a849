"""Deterministic synthetic scenes and cameras: COLMAP datasets are unavailable
offline, so every test / benchmark input is generated here with numpy ``default_rng(seed)`` and
is byte-identical for the CUDA path, the CPU oracle and the compiled reference."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .rasterizer import CameraInfo


@dataclass
class Scene:
    positions: np.ndarray  # [N,3] f32
    sh_coeffs: np.ndarray  # [N,3,C] f32
    opacities: np.ndarray  # [N,1] f32 logit
    rotations: np.ndarray  # [N,4] f32 wxyz, un-normalised
    scales: np.ndarray     # [N,3] f32 log
    camera: CameraInfo

    @property
    def n(self) -> int:
        return int(self.positions.shape[0])


def default_camera(width: int, height: int) -> CameraInfo:
    """camera 0: R = I, t = 0, fx = fy = 0.75 W, principal point at the image centre."""
    return CameraInfo(width, height, 0.75 * width, 0.75 * width, width / 2.0, height / 2.0,
                      np.eye(3, dtype=np.float32), np.zeros(3, dtype=np.float32))


def synth(n: int, width: int, height: int, seed: int = 1234, num_coeffs: int = 16,
          adversarial: bool = False, sigma_px: float = 2.0) -> Scene:
    """Random Gaussians inside the frustum of camera 0.

    depth z = exp(U[ln 1, ln 30]); centre pixel uniform over the image (``adversarial`` widens it
    to |ndc| <= 3 so that off-screen Gaussians, the near-plane cull and the A.2 tile-count quirk all
    fire); pixel-space sigma ~ LogNormal(ln sigma_px, 0.6) clipped to [0.3, 40]; per-axis anisotropy
    exp(U[-0.7, 0.7]); rotations ~ N(0,1)^4 un-normalised; opacity logit ~ N(0, 2);
    SH DC ~ N(0, 1), higher bands ~ N(0, 0.15).
    """
    rng = np.random.default_rng(seed)
    cam = default_camera(width, height)
    z = np.exp(rng.uniform(np.log(1.0), np.log(30.0), size=n))
    if adversarial:
        u = cam.cx + rng.uniform(-3.0, 3.0, size=n) * (width / 2.0)
        v = cam.cy + rng.uniform(-3.0, 3.0, size=n) * (height / 2.0)
        z = np.where(rng.uniform(size=n) < 0.05, rng.uniform(-1.0, 0.3, size=n), z)  # some behind / near
    else:
        u = rng.uniform(0.0, width, size=n)
        v = rng.uniform(0.0, height, size=n)
    pos = np.stack([(u - cam.cx) * z / cam.fx, (v - cam.cy) * z / cam.fy, z], axis=1)
    spx = np.clip(rng.lognormal(np.log(sigma_px), 0.6, size=n), 0.3, 40.0)
    s_world = spx * np.abs(z) / cam.fx
    aniso = np.exp(rng.uniform(-0.7, 0.7, size=(n, 3)))
    scales = np.log(np.maximum(s_world[:, None] * aniso, 1e-8))
    rot = rng.normal(size=(n, 4))
    opa = rng.normal(0.0, 2.0, size=(n, 1))
    sh = rng.normal(0.0, 0.15, size=(n, 3, num_coeffs))
    sh[:, :, 0] = rng.normal(0.0, 1.0, size=(n, 3))
    f = np.float32
    return Scene(pos.astype(f), sh.astype(f), opa.astype(f), rot.astype(f), scales.astype(f), cam)


def ring_cameras(scene: Scene, count: int, seed: int = 99, radius_frac: float = 0.35) -> list:
    """Extra views: cameras on a ring around camera 0's optical axis, all looking at the centroid
    of the frustum contents (used for view-batched training)."""
    rng = np.random.default_rng(seed)
    base = scene.camera
    centroid = np.array([0.0, 0.0, float(np.median(scene.positions[:, 2]))])
    cams = []
    phase = rng.uniform(0, 2 * np.pi)
    r = radius_frac * centroid[2]
    for k in range(count):
        a = phase + 2 * np.pi * k / max(count, 1)
        eye = np.array([r * np.cos(a), r * np.sin(a), 0.0])
        fwd = centroid - eye
        fwd /= np.linalg.norm(fwd)
        up0 = np.array([0.0, -1.0, 0.0])
        right = np.cross(up0, fwd)  # camera x
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)  # camera y
        R = np.stack([right, down, fwd], axis=0)  # world -> camera rotation (rows = camera axes)
        t = -R @ eye
        cams.append(CameraInfo(base.width, base.height, base.fx, base.fy, base.cx, base.cy,
                               R.astype(np.float32), t.astype(np.float32)))
    return cams
