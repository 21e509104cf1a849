#!/usr/bin/env python
"""Parity report (GPU box): the sm_100a path versus the unmodified reference (oracle/_ref) on one
synthetic scene, stage by stage, with mismatch counts and examples. Test infrastructure.

    python tools/parity_report.py [N W H seed [adversarial]]
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
import cugs_ref as ref  # noqa: E402
import cuda_gaussian_splatting_b200 as cugs  # noqa: E402


def bits(t):
    return t.contiguous().view(torch.int32)


def report(name, mine, theirs, exact=True):
    mine, theirs = mine.contiguous(), theirs.contiguous()
    if exact:
        if mine.dtype.is_floating_point:
            neq = bits(mine) != bits(theirs)
        else:
            neq = mine != theirs
        cnt = int(neq.sum())
        msg = f"{name:24s} mismatching elements {cnt} / {mine.numel()}"
        if cnt and mine.dtype.is_floating_point:
            d = (mine.double() - theirs.double()).abs()[neq]
            rel = d / theirs.double().abs()[neq].clamp_min(1e-30)
            idx = neq.reshape(-1).nonzero()[:5].reshape(-1).tolist()
            msg += f"  max abs {float(d.max()):.3e} max rel {float(rel.max()):.3e} first idx {idx}"
            for i in idx[:3]:
                msg += f"\n      [{i}] mine {float(mine.reshape(-1)[i])!r} ({int(bits(mine).reshape(-1)[i]) & 0xffffffff:08x}) ref {float(theirs.reshape(-1)[i])!r} ({int(bits(theirs).reshape(-1)[i]) & 0xffffffff:08x})"
        elif cnt:
            idx = neq.reshape(-1).nonzero()[:8].reshape(-1).tolist()
            msg += f"  first idx {idx} mine {[int(mine.reshape(-1)[i]) for i in idx]} ref {[int(theirs.reshape(-1)[i]) for i in idx]}"
        print(msg)
        return cnt
    d = (mine.double() - theirs.double())
    nrm = float(theirs.double().norm())
    print(f"{name:24s} max abs {float(d.abs().max()):.3e}  norm-rel {float(d.norm()) / max(nrm, 1e-30):.3e}  (ref max {float(theirs.abs().max()):.3e})")
    return 0


def main():
    a = sys.argv[1:]
    n, W, H, seed = (int(a[0]), int(a[1]), int(a[2]), int(a[3])) if len(a) >= 4 else (100_000, 1280, 720, 1235)
    adv = len(a) >= 5
    scene = cugs.synth(n, W, H, seed=seed, adversarial=adv)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    m = cugs.GaussianModel(t(scene.positions), t(scene.sh_coeffs), t(scene.opacities), t(scene.rotations), t(scene.scales))
    cam = scene.camera.as_ref_list()
    deg, bg = 3, (0.1, 0.2, 0.3)
    print(f"== scene N={n} {W}x{H} seed={seed} adversarial={adv}")
    rp = ref.project_gaussians(m.positions, m.rotations, m.scales, m.opacities, m.sh_coeffs, cam, deg, 1.0)
    mp = cugs.project_gaussians(m.positions, m.rotations, m.scales, m.opacities, m.sh_coeffs, scene.camera, deg)
    for nm, a_, b_ in zip(["means_2d", "depths", "cov_2d_inv", "radii", "tiles_touched", "rgb", "opacities_act"],
                          [mp.means_2d, mp.depths, mp.cov_2d_inv, mp.radii, mp.tiles_touched, mp.rgb, mp.opacities_act], rp):
        report("preprocess." + nm, a_, b_)
    rk, rv, rr, rP = ref.sort_gaussians(rp[0], rp[1], rp[3], rp[4], W, H)
    ms = cugs.sort_gaussians(rp[0], rp[1], rp[3], rp[4], W, H)
    print(f"P ref {int(rP.item())} mine {ms.total_pairs}")
    if ms.total_pairs == int(rP.item()):
        report("sort.keys", ms.gaussian_keys_sorted, rk)
        report("sort.values", ms.gaussian_values_sorted, rv)
        report("sort.ranges", ms.tile_ranges, rr)
    st = cugs.RenderSettings(bg, deg, 1.0)
    ro = ref.render(m.positions, m.sh_coeffs, m.opacities, m.rotations, m.scales, cam, list(bg), deg, 1.0)
    mo = cugs.render(m, scene.camera, st)
    torch.cuda.synchronize()
    if mo.gaussian_indices.numel() == ro[9].numel():
        report("render.gaussian_indices", mo.gaussian_indices, ro[9])
    else:
        print("render P differs", mo.gaussian_indices.numel(), ro[9].numel())
    report("render.tile_ranges", mo.tile_ranges, ro[10])
    report("render.n_contrib", mo.n_contrib, ro[2])
    report("render.final_T", mo.final_T, ro[1])
    report("render.color(bits)", mo.color, ro[0])
    report("render.color", mo.color, ro[0], exact=False)
    # forward blend on IDENTICAL inputs (the reference's intermediates): isolates the blend kernel
    f = cugs.rasterize_forward(ro[3], ro[5], ro[7], ro[8], ro[10], ro[9], W, H, bg)
    report("blend(ref inputs).n_contrib", f.n_contrib, ro[2])
    report("blend(ref inputs).final_T", f.final_T, ro[1])
    report("blend(ref inputs).color", f.color, ro[0])
    g = torch.from_numpy(np.random.default_rng(4321).uniform(-1, 1, size=(H, W, 3)).astype(np.float32)).cuda()
    rb = ref.render_backward(g, ro, m.positions, m.sh_coeffs, m.opacities, m.rotations, m.scales, cam, list(bg), deg, 1.0)
    rb2 = ref.render_backward(g, ro, m.positions, m.sh_coeffs, m.opacities, m.rotations, m.scales, cam, list(bg), deg, 1.0)
    mb = cugs.render_backward(g, mo, m, scene.camera, st)
    names = ["dL_dpositions", "dL_drotations", "dL_dscales", "dL_dopacities", "dL_dsh_coeffs", "dL_dmeans_2d"]
    for nm, r1, r2 in zip(names, rb, rb2):
        report("backward." + nm, getattr(mb, nm), r1, exact=False)
        report("  (ref vs ref rerun)", r2, r1, exact=False)
        d = (getattr(mb, nm).double() - r1.double()).abs().reshape(-1)
        i = int(d.argmax())
        print(f"      worst element {i}: mine {float(getattr(mb, nm).reshape(-1)[i])!r} ref {float(r1.reshape(-1)[i])!r} ref-rerun {float(r2.reshape(-1)[i])!r}")


if __name__ == "__main__":
    main()
