"""Generates tests/golden/ref_gpu_golden.npz on a B200: outputs of the UNMODIFIED reference CUDA
kernels (oracle/_ref/cugs_ref*.so, built from /root/reference by oracle/Makefile.ref) on small seeded
scenes that the tests regenerate with cuda_gaussian_splatting_b200.synth. Run on the GPU box:

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/ref_gpu_golden.npz'

and copy the file to tests/golden/. The CPU-only suite then checks the CPU oracle against it
(tests/test_oracle_vs_reference_golden.py)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
import cugs_ref as ref  # noqa: E402
import cuda_gaussian_splatting_b200 as cugs  # noqa: E402  (synth only: the scene generator)

SCENES = {"plain": dict(n=2000, w=160, h=120, seed=77, adversarial=False, deg=3),
          "adversarial": dict(n=1500, w=160, h=120, seed=78, adversarial=True, deg=2)}
out = {}
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for name, c in SCENES.items():
    s = cugs.synth(c["n"], c["w"], c["h"], seed=c["seed"], adversarial=c["adversarial"])
    pos, sh, opa, rot, scl = t(s.positions), t(s.sh_coeffs), t(s.opacities), t(s.rotations), t(s.scales)
    cam, bg = s.camera.as_ref_list(), [0.1, 0.2, 0.3]
    p = ref.project_gaussians(pos, rot, scl, opa, sh, cam, c["deg"], 1.0)
    for k, v in zip(["means_2d", "depths", "cov_2d_inv", "radii", "tiles_touched", "rgb", "opacities_act"], p):
        out[f"{name}.{k}"] = v.cpu().numpy()
    keys, vals, ranges, P = ref.sort_gaussians(p[0], p[1], p[3], p[4], c["w"], c["h"])
    out[f"{name}.keys_sorted"] = keys.cpu().numpy().view(np.uint64)
    out[f"{name}.gaussian_indices"] = vals.cpu().numpy()
    out[f"{name}.tile_ranges"] = ranges.cpu().numpy()
    r = ref.render(pos, sh, opa, rot, scl, cam, bg, c["deg"], 1.0)
    out[f"{name}.color"], out[f"{name}.final_T"], out[f"{name}.n_contrib"] = (x.cpu().numpy() for x in r[:3])
    g = torch.from_numpy(np.random.default_rng(c["seed"] + 1000).uniform(-1, 1, size=(c["h"], c["w"], 3)).astype(np.float32)).cuda()
    b = ref.render_backward(g, r, pos, sh, opa, rot, scl, cam, bg, c["deg"], 1.0)
    for k, v in zip(["dL_dpositions", "dL_drotations", "dL_dscales", "dL_dopacities", "dL_dsh_coeffs", "dL_dmeans_2d"], b):
        out[f"{name}.{k}"] = v.cpu().numpy()
# loss + autograd gradient, and three FusedAdam steps
rng = np.random.default_rng(5)
x = rng.uniform(size=(40, 56, 3)).astype(np.float32)
y = rng.uniform(size=(40, 56, 3)).astype(np.float32)
l, l1, ss, gr = ref.combined_loss_with_grad(t(x), t(y), 0.2)
out["loss.scalars"] = np.array([float(l), float(l1), float(ss)], np.float32)
out["loss.grad"] = gr.cpu().numpy()
out["loss.ssim_map"] = ref.ssim(t(x), t(y)).cpu().numpy()
s = cugs.synth(257, 64, 48, seed=79)
params = [t(s.positions), t(s.sh_coeffs), t(s.opacities), t(s.rotations), t(s.scales)]
opt = ref.FusedAdam(*params)
for step in range(3):
    grads = [torch.from_numpy(rng.normal(size=tuple(q.shape)).astype(np.float32)).cuda()
             for q in (params[0], params[3], params[4], params[2], params[1])]  # pos, rot, scl, opa, sh
    out[f"adam.grads{step}"] = np.concatenate([q.cpu().numpy().reshape(-1) for q in grads])
    opt.step(grads, step)
for k, v in zip(["positions", "sh_coeffs", "opacities", "rotations", "scales"], opt.params()):
    out[f"adam.{k}"] = v.detach().cpu().numpy()
dst = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "tests" / "golden" / "ref_gpu_golden.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, sum(v.nbytes for v in out.values()), "bytes uncompressed")
