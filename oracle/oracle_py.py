"""ctypes front-end of oracle/cugs_oracle.c (numpy in, numpy out).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module. It is the checker, never the product.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "libcugs_oracle.so"
_lib = None

_f = np.float32
_i = np.int32


def build() -> Path:
    src = _DIR / "cugs_oracle.c"
    if not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-f", str(_DIR / "Makefile")], cwd=str(_DIR.parent))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
        _lib.oracle_scan.restype = C.c_int64
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def view_arrays(camera):
    view = _c(camera.world_to_camera().reshape(-1), _f)
    cam = _c(camera.camera_center(), _f)
    return view, cam


def preprocess_fwd(scene_arrays, camera, deg, scale_mod=1.0):
    pos, rot, scl, opa, sh = [_c(a, _f) for a in scene_arrays]
    n = pos.shape[0]
    Cn = sh.shape[2]
    view, cam = view_arrays(camera)
    out = dict(means_2d=np.empty((n, 2), _f), depths=np.empty((n,), _f), cov_2d_inv=np.empty((n, 3), _f),
               radii=np.empty((n,), _i), tiles_touched=np.empty((n,), _i), rgb=np.empty((n, 3), _f),
               opacities_act=np.empty((n,), _f))
    lib().oracle_preprocess_fwd(C.c_int64(n), _p(view), C.c_float(camera.fx), C.c_float(camera.fy),
                                C.c_float(camera.cx), C.c_float(camera.cy), C.c_int(camera.width),
                                C.c_int(camera.height), C.c_float(scale_mod), _p(cam), C.c_int(deg), C.c_int(Cn),
                                _p(pos), _p(rot), _p(scl), _p(opa), _p(sh), _p(out["means_2d"]), _p(out["depths"]),
                                _p(out["cov_2d_inv"]), _p(out["radii"]), _p(out["tiles_touched"]), _p(out["rgb"]),
                                _p(out["opacities_act"]))
    return out


def scan(tiles):
    tiles = _c(tiles, _i)
    off = np.empty_like(tiles)
    total = lib().oracle_scan(C.c_int64(tiles.shape[0]), _p(tiles), _p(off))
    return off, int(total)


def fill_keys(means_2d, depths, radii, offsets, W, H, P):
    keys = np.empty((P,), np.uint64)
    vals = np.empty((P,), _i)
    lib().oracle_fill_keys(C.c_int64(radii.shape[0]), _p(_c(means_2d, _f)), _p(_c(depths, _f)), _p(_c(radii, _i)),
                           _p(_c(offsets, _i)), C.c_int(W), C.c_int(H), C.c_int64(P), _p(keys), _p(vals))
    return keys, vals


def sort_pairs(keys, vals):
    keys, vals = _c(keys, np.uint64), _c(vals, _i)
    ko, vo = np.empty_like(keys), np.empty_like(vals)
    lib().oracle_sort_pairs(C.c_int64(keys.shape[0]), _p(keys), _p(vals), _p(ko), _p(vo))
    return ko, vo


def tile_ranges(keys_sorted, num_tiles):
    r = np.empty((num_tiles, 2), _i)
    lib().oracle_tile_ranges(C.c_int64(keys_sorted.shape[0]), _p(_c(keys_sorted, np.uint64)), C.c_int(num_tiles), _p(r))
    return r


def blend_fwd(W, H, bg, ranges, gidx, means_2d, conic, rgb, opa):
    color = np.empty((H, W, 3), _f)
    T = np.empty((H, W), _f)
    nc = np.empty((H, W), _i)
    stats = np.zeros(3, np.int64)
    lib().oracle_blend_fwd(C.c_int(W), C.c_int(H), _p(_c(bg, _f)), _p(_c(ranges, _i)), _p(_c(gidx, _i)),
                           _p(_c(means_2d, _f)), _p(_c(conic, _f)), _p(_c(rgb, _f)), _p(_c(opa, _f)), _p(color),
                           _p(T), _p(nc), _p(stats))
    return color, T, nc, stats


def blend_bwd(W, H, bg, ranges, gidx, means_2d, conic, rgb, opa, dL_dcolor, final_T, n_contrib, n):
    d_rgb, d_opa = np.empty((n, 3), _f), np.empty((n,), _f)
    d_mean, d_conic = np.empty((n, 2), _f), np.empty((n, 3), _f)
    stats = np.zeros(2, np.int64)
    lib().oracle_blend_bwd(C.c_int(W), C.c_int(H), _p(_c(bg, _f)), _p(_c(ranges, _i)), _p(_c(gidx, _i)),
                           _p(_c(means_2d, _f)), _p(_c(conic, _f)), _p(_c(rgb, _f)), _p(_c(opa, _f)),
                           _p(_c(dL_dcolor, _f)), _p(_c(final_T, _f)), _p(_c(n_contrib, _i)), C.c_int64(n),
                           _p(d_rgb), _p(d_opa), _p(d_mean), _p(d_conic), _p(stats))
    return d_rgb, d_opa, d_mean, d_conic, stats


def preprocess_bwd(scene_arrays, camera, deg, radii, dL_dmean2d, dL_dconic, dL_drgb, dL_dopa_act, scale_mod=1.0):
    pos, rot, scl, opa, sh = [_c(a, _f) for a in scene_arrays]
    n, Cn = pos.shape[0], sh.shape[2]
    view, cam = view_arrays(camera)
    out = dict(dL_dpositions=np.empty((n, 3), _f), dL_drotations=np.empty((n, 4), _f),
               dL_dscales=np.empty((n, 3), _f), dL_dopacities=np.empty((n, 1), _f), dL_dsh_coeffs=np.empty_like(sh))
    lib().oracle_preprocess_bwd(C.c_int64(n), _p(view), C.c_float(camera.fx), C.c_float(camera.fy),
                                C.c_float(scale_mod), _p(cam), C.c_int(deg), C.c_int(Cn), _p(pos), _p(rot), _p(scl),
                                _p(opa), _p(sh), _p(_c(radii, _i)), _p(_c(dL_dmean2d, _f)), _p(_c(dL_dconic, _f)),
                                _p(_c(dL_drgb, _f)), _p(_c(dL_dopa_act, _f)), _p(out["dL_dpositions"]),
                                _p(out["dL_drotations"]), _p(out["dL_dscales"]), _p(out["dL_dopacities"]),
                                _p(out["dL_dsh_coeffs"]))
    return out


def sh_forward(deg, sh, dirs, clamp=False):
    sh, dirs = _c(sh, _f), _c(dirs, _f)
    out = np.empty((sh.shape[0], 3), _f)
    lib().oracle_sh_forward(C.c_int64(sh.shape[0]), C.c_int(deg), C.c_int(sh.shape[2]), _p(sh), _p(dirs), _p(out),
                            C.c_int(1 if clamp else 0))
    return out


def sh_backward(deg, sh, dirs, dL_drgb):
    sh, dirs, g = _c(sh, _f), _c(dirs, _f), _c(dL_drgb, _f)
    out = np.empty_like(sh)
    lib().oracle_sh_backward(C.c_int64(sh.shape[0]), C.c_int(deg), C.c_int(sh.shape[2]), _p(sh), _p(dirs), _p(g), _p(out))
    return out


def loss(rendered, target, lam=0.2, want_grad=True):
    r, t = _c(rendered, _f), _c(target, _f)
    H, W = r.shape[0], r.shape[1]
    g = np.empty_like(r) if want_grad else None
    sc = np.empty(3, _f)
    lib().oracle_loss(C.c_int(W), C.c_int(H), C.c_float(lam), _p(r), _p(t), _p(g) if want_grad else None, _p(sc))
    return sc, g


def adam(p, g, m, v, lr, b1, b2, eps, bc1, bc2):
    lib().oracle_adam(C.c_int64(p.size), _p(p), _p(_c(g, _f)), _p(m), _p(v), C.c_float(lr), C.c_float(b1),
                      C.c_float(b2), C.c_float(eps), C.c_float(bc1), C.c_float(bc2))


def accumulate_stats(dL_dmean2d, radii, grad_accum, grad_count, max_radii):
    lib().oracle_accumulate_stats(C.c_int64(radii.shape[0]), _p(_c(dL_dmean2d, _f)), _p(_c(radii, _i)),
                                  _p(grad_accum), _p(grad_count), _p(max_radii))


def render_forward(scene, deg=3, bg=(0.0, 0.0, 0.0), scale_mod=1.0):
    """Whole forward pipeline on the CPU; returns a dict of every stage output."""
    cam = scene.camera
    arrs = (scene.positions, scene.rotations, scene.scales, scene.opacities, scene.sh_coeffs)
    o = preprocess_fwd(arrs, cam, deg, scale_mod)
    off, P = scan(o["tiles_touched"])
    keys, vals = fill_keys(o["means_2d"], o["depths"], o["radii"], off, cam.width, cam.height, P)
    ks, vs = sort_pairs(keys, vals)
    nt = ((cam.width + 15) // 16) * ((cam.height + 15) // 16)
    rng = tile_ranges(ks, nt)
    color, T, nc, st = blend_fwd(cam.width, cam.height, bg, rng, vs, o["means_2d"], o["cov_2d_inv"], o["rgb"],
                                 o["opacities_act"])
    o.update(offsets=off, P=P, keys_unsorted=keys, values_unsorted=vals, keys_sorted=ks, gaussian_indices=vs,
             tile_ranges=rng, color=color, final_T=T, n_contrib=nc, fwd_stats=st)
    return o


def render_backward(scene, fwd, dL_dcolor, deg=3, bg=(0.0, 0.0, 0.0), scale_mod=1.0):
    cam = scene.camera
    arrs = (scene.positions, scene.rotations, scene.scales, scene.opacities, scene.sh_coeffs)
    d_rgb, d_opa, d_mean, d_conic, st = blend_bwd(cam.width, cam.height, bg, fwd["tile_ranges"],
                                                  fwd["gaussian_indices"], fwd["means_2d"], fwd["cov_2d_inv"],
                                                  fwd["rgb"], fwd["opacities_act"], dL_dcolor, fwd["final_T"],
                                                  fwd["n_contrib"], scene.n)
    out = preprocess_bwd(arrs, cam, deg, fwd["radii"], d_mean, d_conic, d_rgb, d_opa, scale_mod)
    out.update(dL_drgb=d_rgb, dL_dopacity_act=d_opa, dL_dmeans_2d=d_mean, dL_dcov_2d_inv=d_conic, bwd_stats=st)
    return out
