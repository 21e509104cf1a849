"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): ragged sizes, culled and
off-screen Gaussians, empty tiles, a training step."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import cuda_gaussian_splatting_b200 as cugs

def model_of(scene):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return cugs.GaussianModel(t(scene.positions), t(scene.sh_coeffs), t(scene.opacities), t(scene.rotations), t(scene.scales))

for (n, w, h, adv, c, deg) in [(3001, 333, 211, False, 16, 3), (5000, 200, 120, True, 16, 3), (777, 64, 48, False, 4, 1), (33, 17, 9, False, 1, 0)]:
    scene = cugs.synth(n, w, h, seed=5, adversarial=adv, num_coeffs=c)
    m = model_of(scene)
    st = cugs.RenderSettings((0.1, 0.2, 0.3), deg, 1.0)
    out = cugs.render(m, scene.camera, st)
    tgt = torch.rand((h, w, 3), device="cuda")
    sc, g = cugs.combined_loss_with_grad(out.color, tgt, 0.2)
    b = cugs.render_backward(g, out, m, scene.camera, st)
    tr = cugs.SyntheticTrainer(m, [scene.camera], [tgt], cugs.TrainConfig(max_sh_degree=deg))
    tr.train_step(5000)
    p = cugs.project_gaussians(m.positions, m.rotations, m.scales, m.opacities, m.sh_coeffs, scene.camera, deg)
    s = cugs.sort_gaussians(p.means_2d, p.depths, p.radii, p.tiles_touched, w, h)
    f = cugs.rasterize_forward(p.means_2d, p.cov_2d_inv, p.rgb, p.opacities_act, s.tile_ranges, s.gaussian_values_sorted, w, h, (0, 0, 0))
    rb = cugs.rasterize_backward(g, p.means_2d, p.cov_2d_inv, p.rgb, p.opacities_act, s.tile_ranges, s.gaussian_values_sorted, f.final_T, f.n_contrib, w, h, (0, 0, 0), n)
    torch.cuda.synchronize()
    print("ok", n, w, h, int(out.gaussian_indices.numel()), float(sc[0]))
print("sanitize smoke done")
