"""GPU parity tests for the code paths the round-1 suite did not reach ("weak" #1):

* every tile-bit plan of the fused path's packed pair sort -- 14 bits at 2560x1440 (7+7), 15 bits at
  3840x2160 (8+7), 16 bits at 4096x2176 (8+8) -- through render() against the compiled reference;
* the packed sort itself (cugs_b200_sort_packed, tile_binning.cu) on raw elements for key widths 1..16,
  20 and 32 bits, ragged sizes, with and without the tile-range output, and with the element count read
  on the device (capacity-sized launches);
* render(sync=False): no host round trip for P, bitwise equal to the blocking path; overflow is flagged and safe;
* BASELINE config E (20 M Gaussians, 3840x2160, P = 125 M pairs) against the compiled reference.

Reference anchors: rasterizer/sorting.cu:82-109 (tile ranges), :190-211 (the 64-bit CUB sort whose order
must be reproduced bit for bit). Nothing here reads /root/reference at run time.
"""
import ctypes as C

import numpy as np
import pytest

import cuda_gaussian_splatting_b200 as cugs
from cuda_gaussian_splatting_b200 import _lib
from conftest import to_torch

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4
GRAD_REL = 1e-3


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("no CUDA device")
    return t


def _dL(torch, H, W, seed=4321):
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.uniform(-1, 1, size=(H, W, 3)).astype(np.float32)).cuda()


def _forward_backward_vs_reference(ref, torch, scene, deg=3, bg=(0.1, 0.2, 0.3)):
    m = to_torch(scene)
    cam = scene.camera.as_ref_list()
    r = ref.render(m.positions, m.sh_coeffs, m.opacities, m.rotations, m.scales, cam, list(bg), deg, 1.0)
    settings = cugs.RenderSettings(bg, deg, 1.0)
    out = cugs.render(m, scene.camera, settings)
    torch.cuda.synchronize()
    assert torch.equal(out.radii, r[6]), "radii"
    assert torch.equal(out.tile_ranges, r[10]), "tile ranges must be bit-exact"
    assert torch.equal(out.gaussian_indices, r[9]), "sort order must be bit-exact"
    assert torch.equal(out.n_contrib, r[2]), "n_contrib must be bit-exact"
    assert float((out.color - r[0]).abs().max()) <= IMG_TOL
    assert float((out.final_T - r[1]).abs().max()) <= 1e-6
    g = _dL(torch, scene.camera.height, scene.camera.width)
    rb = ref.render_backward(g, r, m.positions, m.sh_coeffs, m.opacities, m.rotations, m.scales, cam, list(bg), deg, 1.0)
    b = cugs.render_backward(g, out, m, scene.camera, settings)
    for nm, rt in zip(["dL_dpositions", "dL_drotations", "dL_dscales", "dL_dopacities", "dL_dsh_coeffs", "dL_dmeans_2d"], rb):
        mine_t, ref_t = getattr(b, nm).double(), rt.double()
        rel = float((mine_t - ref_t).norm() / ref_t.norm().clamp_min(1e-30))
        assert rel <= GRAD_REL, f"{nm}: norm-rel {rel:.3e}"
    return int(out.gaussian_indices.numel())


# (name, N, W, H, seed, tile bits, plan)
RESOLUTIONS = [
    ("1440p_14bit_7+7", 200_000, 2560, 1440, 2101, 14),
    ("4K_15bit_8+7", 300_000, 3840, 2160, 2102, 15),
    ("4096x2176_16bit_8+8", 150_000, 4096, 2176, 2103, 16),
    ("big_splats_4K", 20_000, 3840, 2160, 2104, 15),       # sigma 25 px: long per-tile lists at 15 tile bits
]


@pytest.mark.parametrize("name,n,W,H,seed,bits", RESOLUTIONS, ids=[r[0] for r in RESOLUTIONS])
def test_resolution_plans_vs_reference(ref, torch, name, n, W, H, seed, bits):
    lib, h = _lib.load_library(), _lib.handle(0)
    scene = cugs.synth(n, W, H, seed=seed, sigma_px=25.0 if name.startswith("big") else 2.0)
    assert max(0, (cugs.rasterizer.num_tiles(W, H) - 1).bit_length()) == bits
    _forward_backward_vs_reference(ref, torch, scene)
    passes, key_bits = C.c_int(0), C.c_int(0)
    lib.cugs_b200_last_sort_plan(h, C.byref(passes), C.byref(key_bits))
    assert key_bits.value == 32 + bits and passes.value == 4 + 2


def _packed_sort(torch, lib, h, hi, lo, key_bits, num_tiles, want_ranges, out32, n_cap=None, use_n_dev=False):
    n = hi.shape[0]
    cap = n_cap or n
    elts = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    a = torch.zeros((cap,), dtype=torch.int64, device="cuda")
    a[:n] = torch.from_numpy(elts.view(np.int64)).cuda()
    b = torch.empty_like(a)
    o32 = torch.full((cap,), -7, dtype=torch.int32, device="cuda") if out32 else None
    rng = torch.full((max(num_tiles, 1), 2), -1, dtype=torch.int32, device="cuda") if want_ranges else None
    tmp = torch.empty((lib.cugs_b200_sort_packed_temp_bytes(cap, key_bits, num_tiles if want_ranges else 0),),
                      dtype=torch.uint8, device="cuda")
    n_dev = torch.tensor([n], dtype=torch.int64, device="cuda") if use_n_dev else None
    st = lib.cugs_b200_sort_packed(h, torch.cuda.current_stream().cuda_stream, cap if use_n_dev else n, key_bits,
                                   a.data_ptr(), b.data_ptr(), o32.data_ptr() if out32 else None,
                                   num_tiles if want_ranges else 0, rng.data_ptr() if want_ranges else None,
                                   tmp.data_ptr(), tmp.numel(), n_dev.data_ptr() if use_n_dev else None)
    _lib.check(h, st, "cugs_b200_sort_packed")
    torch.cuda.synchronize()
    passes = lib.cugs_b200_sort_packed_passes(key_bits)
    res = (b if passes & 1 else a).cpu().numpy().view(np.uint64)[:n]
    return res, (o32.cpu().numpy() if out32 else None), (rng.cpu().numpy() if want_ranges else None), elts


@pytest.mark.parametrize("key_bits", list(range(1, 17)) + [20, 32])
def test_packed_sort_every_key_width(torch, key_bits):
    """cugs_b200_sort_packed against numpy's stable sort: all digit plans (1 pass of 1..8 bits, 2 passes of
    5+4 .. 8+8, 3 and 4 passes), element counts around the 4096-element tile, duplicates (stability)."""
    lib, h = _lib.load_library(), _lib.handle(0)
    rng = np.random.default_rng(100 + key_bits)
    plan_passes = lib.cugs_b200_sort_packed_passes(key_bits)
    assert plan_passes == max(1, -(-key_bits // 8))
    for n in (1, 33, 4095, 4096, 4097, 50_001, 700_000):
        hi_max = (1 << key_bits) if key_bits < 32 else (1 << 32)
        num_tiles = min(hi_max, 40_000) if key_bits <= 16 else 0
        top = num_tiles if num_tiles else hi_max
        hi = rng.integers(0, top, size=n, dtype=np.uint64).astype(np.uint32)
        if n > 1000:
            hi[: n // 3] = hi[0]                      # one very long run: stability across tiles
        lo = np.arange(n, dtype=np.uint32)
        order = np.argsort(hi, kind="stable")
        want_ranges = num_tiles > 0
        for out32 in (False, True):
            res, o32, ranges, elts = _packed_sort(torch, lib, h, hi, lo, key_bits, num_tiles, want_ranges, out32)
            if out32:
                assert np.array_equal(o32.view(np.uint32), lo[order]), (key_bits, n, "payload order")
            else:
                assert np.array_equal(res, elts[order]), (key_bits, n, "element order")
            if want_ranges:
                counts = np.bincount(hi, minlength=num_tiles)[:num_tiles]
                ends = np.cumsum(counts)
                starts = ends - counts
                exp = np.stack([np.where(counts > 0, starts, 0), np.where(counts > 0, ends, 0)], axis=1)
                assert np.array_equal(ranges, exp.astype(np.int32)), (key_bits, n, "ranges")


@pytest.mark.parametrize("key_bits", [6, 13, 15, 16])
def test_packed_sort_count_on_the_device(torch, key_bits):
    """Capacity-sized launches: the element count is read on the device (n_dev), buffers are larger."""
    lib, h = _lib.load_library(), _lib.handle(0)
    rng = np.random.default_rng(7 + key_bits)
    num_tiles = min(1 << key_bits, 33_000)
    for n, cap in ((0, 5000), (1, 4096), (4097, 9000), (100_003, 131_072), (100_003, 100_003)):
        hi = rng.integers(0, num_tiles, size=n, dtype=np.uint64).astype(np.uint32)
        lo = np.arange(n, dtype=np.uint32)
        order = np.argsort(hi, kind="stable")
        res, o32, ranges, elts = _packed_sort(torch, lib, h, hi, lo, key_bits, num_tiles, True, True, n_cap=cap,
                                              use_n_dev=True)
        assert np.array_equal(o32[:n].view(np.uint32), lo[order]), (key_bits, n, cap)
        assert (o32[n:] == -7).all(), "entries beyond the device-side count must not be written"
        counts = np.bincount(hi, minlength=num_tiles)[:num_tiles]
        ends = np.cumsum(counts)
        exp = np.stack([np.where(counts > 0, ends - counts, 0), np.where(counts > 0, ends, 0)], axis=1)
        assert np.array_equal(ranges, exp.astype(np.int32))


@pytest.mark.parametrize("key_bits,n,cap", [(13, (8 << 20) - 1, None), (13, (8 << 20) + 8193, None),
                                            (15, 5_000_003, 9_000_000), (24, (8 << 20) + 1, None)])
def test_packed_sort_across_the_large_tile_limit(torch, key_bits, n, cap):
    """From 8 Mi elements of launch capacity the passes use 8192-element tiles (16 items per thread,
    tile_binning.cu pk_items_for): one element below the limit, ragged tails above it, a capacity above the
    limit with a device-side count below it, and a 3-pass plan."""
    lib, h = _lib.load_library(), _lib.handle(0)
    rng = np.random.default_rng(n % 1000 + key_bits)
    num_tiles = min(1 << key_bits, 32_400) if key_bits <= 16 else 0
    hi = rng.integers(0, num_tiles or (1 << key_bits), size=n, dtype=np.uint64).astype(np.uint32)
    hi[: n // 4] = hi[0]
    lo = np.arange(n, dtype=np.uint32)
    order = np.argsort(hi, kind="stable")
    res, o32, ranges, elts = _packed_sort(torch, lib, h, hi, lo, key_bits, num_tiles, num_tiles > 0, True,
                                          n_cap=cap, use_n_dev=cap is not None)
    assert np.array_equal(o32[:n].view(np.uint32), lo[order])
    if cap is not None:
        assert (o32[n:] == -7).all()
    if num_tiles:
        counts = np.bincount(hi, minlength=num_tiles)[:num_tiles]
        ends = np.cumsum(counts)
        exp = np.stack([np.where(counts > 0, ends - counts, 0), np.where(counts > 0, ends, 0)], axis=1)
        assert np.array_equal(ranges, exp.astype(np.int32))


@pytest.mark.parametrize("name,n,W,H,seed", [("ragged", 3001, 333, 211, 12), ("A", 100_000, 1280, 720, 1235),
                                              ("adversarial", 20_000, 640, 360, 13)])
def test_render_without_host_sync_equals_blocking_path(torch, name, n, W, H, seed):
    """render(sync=False) == render(sync=True), bit for bit, forward and backward; status = {P, 0}."""
    scene = cugs.synth(n, W, H, seed=seed, adversarial=(name == "adversarial"))
    m = to_torch(scene)
    settings = cugs.RenderSettings((0.2, 0.1, 0.0), 3, 1.0)
    b1 = cugs.FrameBuffers(n, W, H, 16, "cuda")
    o1 = cugs.render(m, scene.camera, settings, b1)
    P = int(o1.gaussian_indices.numel())
    g = _dL(torch, H, W)
    g1 = cugs.render_backward(g, o1, m, scene.camera, settings, b1)
    ref_vals = {k: getattr(o1, k).clone() for k in ("color", "final_T", "n_contrib", "tile_ranges", "radii")}
    ref_idx = o1.gaussian_indices.clone()
    ref_grads = {k: getattr(g1, k).clone() for k in ("dL_dpositions", "dL_dsh_coeffs", "dL_dmeans_2d")}

    b2 = cugs.FrameBuffers(n, W, H, 16, "cuda")
    b2.ensure_capacity(P)                       # capacity = 1.25 P + 1024: larger than P
    assert b2.p_capacity > P
    b2.gaussian_indices.fill_(-1)
    o2 = cugs.render(m, scene.camera, settings, b2, sync=False)
    b2.fetch_status()
    torch.cuda.synchronize()
    assert b2.last_pairs() == (P, False)
    for k, v in ref_vals.items():
        assert torch.equal(getattr(o2, k), v), k
    assert torch.equal(o2.gaussian_indices[:P], ref_idx)
    assert bool((o2.gaussian_indices[P:] == -1).all()), "entries beyond P must not be written"
    g2 = cugs.render_backward(g, o2, m, scene.camera, settings, b2)
    for k, v in ref_grads.items():   # the backward's float atomics make sums order-dependent: tolerance, not bits
        a = getattr(g2, k).double()
        assert float((a - v.double()).norm() / v.double().norm().clamp_min(1e-30)) <= 1e-5, k

    # overflow: capacity below P -> flagged, nothing runs past a buffer, the next blocking frame is fine
    b3 = cugs.FrameBuffers(n, W, H, 16, "cuda")
    b3.ensure_capacity(max(P // 4, 1))
    if b3.p_capacity < P:
        guard = torch.full((4096,), 12345, dtype=torch.int32, device="cuda")  # allocated right after the buffers
        cugs.render(m, scene.camera, settings, b3, sync=False)
        b3.fetch_status()
        torch.cuda.synchronize()
        assert b3.last_pairs() == (P, True)
        assert bool((guard == 12345).all())
        o3 = cugs.render(m, scene.camera, settings, b3)   # blocking path grows the buffers
        assert torch.equal(o3.color, ref_vals["color"]) and torch.equal(o3.gaussian_indices, ref_idx)


def test_config_E_20M_4K_vs_reference(ref, torch):
    """BASELINE config E: 20 M Gaussians at 3840x2160 (P = 124.8 M pairs, 15 tile bits): integers bit-exact,
    image and gradients in tolerance, against the compiled reference on the same GPU."""
    free, total = torch.cuda.mem_get_info()
    if total < 120 * (1 << 30):
        pytest.skip("needs a 180 GB B200")
    scene = cugs.synth(20_000_000, 3840, 2160, seed=1239)
    P = _forward_backward_vs_reference(ref, torch, scene, bg=(0.0, 0.0, 0.0))
    assert P > 100_000_000
    torch.cuda.empty_cache()


def test_evaluation_counter_matches_the_forward_and_a_numpy_walk(torch):
    """cugs_b200_count_evaluations (the work units of the blend roofline): contributing counts equal
    sum(n_contrib) exactly, forward and backward; the rejected counts match a numpy walk of the reference
    traversal (forward.cu:121-157, backward.cu:117-145) on a small scene."""
    scene = cugs.synth(3000, 160, 112, seed=41, sigma_px=5.0)
    m = to_torch(scene)
    out = cugs.render(m, scene.camera, cugs.RenderSettings((0, 0, 0), 3, 1.0))
    c = cugs.count_evaluations(out, scene.camera)
    total = int(out.n_contrib.sum())
    assert c["fwd_contributing"] == total and c["bwd_contributing"] == total
    W, H = scene.camera.width, scene.camera.height
    ntx = (W + 15) // 16
    rg, idx = out.tile_ranges.cpu().numpy(), out.gaussian_indices.cpu().numpy()
    m2, con, op = out.means_2d.cpu().numpy(), out.cov_2d_inv.cpu().numpy(), out.opacities_act.cpu().numpy()
    ncon = out.n_contrib.cpu().numpy()
    f32 = np.float32
    fr = br = 0
    for t in range(rg.shape[0]):
        s, e = int(rg[t, 0]), int(rg[t, 1])
        if e <= s:
            continue
        ty, tx = divmod(t, ntx)
        py, px = np.meshgrid(np.arange(ty * 16, ty * 16 + 16), np.arange(tx * 16, tx * 16 + 16), indexing="ij")
        inside = (px < W) & (py < H)
        pxf, pyf = px.astype(f32) + f32(0.5), py.astype(f32) + f32(0.5)
        g = idx[s:e]
        dx = pxf[..., None] - m2[g, 0][None, None, :]
        dy = pyf[..., None] - m2[g, 1][None, None, :]
        a, b, cc = con[g, 0], con[g, 1], con[g, 2]
        power = f32(-0.5) * (dx * (a * dx + b * dy) + dy * (b * dx + cc * dy))
        alpha = np.minimum(op[g] * np.exp(power.astype(np.float64)).astype(f32), f32(0.99))
        passing = (power <= 0) & (alpha >= f32(1.0 / 255.0))
        # forward: walk until T < 1/255 (the crossing Gaussian is evaluated and composited)
        T = np.ones(px.shape, dtype=f32)
        alive = inside.copy()
        for j in range(e - s):
            fr += int((alive & ~passing[..., j]).sum())
            T = np.where(alive & passing[..., j], T * (f32(1.0) - alpha[..., j]), T)
            alive &= ~(passing[..., j] & (T < f32(1.0 / 255.0)))
        # backward: from the back until more than n_contrib passing Gaussians were met
        nc = ncon[np.clip(py, 0, H - 1), np.clip(px, 0, W - 1)]
        alive = inside.copy()
        found = np.zeros(px.shape, dtype=np.int64)
        for j in range(e - s - 1, -1, -1):
            br += int((alive & ~passing[..., j]).sum())
            found += (alive & passing[..., j])
            over = alive & passing[..., j] & (found > nc)
            br += int(over.sum())
            alive &= ~over
    assert abs(c["fwd_rejected"] - fr) <= 1e-3 * max(fr, 1), (c, fr)
    assert abs(c["bwd_rejected"] - br) <= 1e-3 * max(br, 1), (c, br)
