"""View-parallel training with the model-resizing schedules switched on (N GPUs, torchrun): every
rank renders its own view, gradients and densification statistics are exchanged once per step, ADC
densification / MCMC relocation run on every rank independently -- and the replicas must stay
BIT-IDENTICAL (same all-reduced statistics, same Philox draws). Prints one line per mode."""
import hashlib
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import cuda_gaussian_splatting_b200 as cugs

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def digest(model):
    h = hashlib.sha256()
    for x in (model.positions, model.sh_coeffs, model.opacities, model.rotations, model.scales):
        h.update(x.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def run(mode):
    gt_scene = cugs.synth(6000, 320, 240, seed=41)
    gt = cugs.GaussianModel(t(gt_scene.positions), t(gt_scene.sh_coeffs), t(gt_scene.opacities), t(gt_scene.rotations),
                            t(gt_scene.scales))
    cams = cugs.ring_cameras(gt_scene, 2 * world)
    st = cugs.RenderSettings((0, 0, 0), 3, 1.0)
    mine = cugs.shard_views(len(cams), world, rank)
    targets = [cugs.render(gt, cams[v], st).color.clone() for v in mine]
    s = cugs.synth(2500, 320, 240, seed=42)
    model = cugs.GaussianModel(t(s.positions), t(s.sh_coeffs), t(s.opacities), t(s.rotations), t(s.scales))
    if mode == "adc":
        extent = float(torch.exp(model.scales).max(dim=1).values.median()) / 0.01
        cfg = cugs.TrainConfig(densification=cugs.DensificationConfig(densify_from=6, densify_every=6, densify_until=30,
                                                                       opacity_reset_every=14, grad_threshold=2e-5),
                               scene_extent=extent)
    else:
        model.opacities[::5] = -8.0
        cfg = cugs.TrainConfig(mcmc=cugs.MCMCConfig(relocate_from=4, relocate_every=4, relocate_until=28,
                                                    noise_lr_init=0.5, noise_lr_final=0.1),
                               mcmc_relocation=True, scene_extent=3.0)
    tr = cugs.SyntheticTrainer(model, [cams[v] for v in mine], targets, cfg, total_views_per_step=len(cams))
    events, sizes = 0, []
    for step in range(32):
        loss = float(tr.train_step(step)[0])
        assert np.isfinite(loss)
        events += tr.last_density_event is not None
        sizes.append(model.num_gaussians())
    d = digest(model)
    all_d = [None] * world
    dist.all_gather_object(all_d, (d, sizes[-1]))
    same = all(x == all_d[0] for x in all_d)
    if rank == 0:
        print(f"{mode}: world {world}, 32 steps, {events} resize events, N {sizes[0]} -> {sizes[-1]}, "
              f"replicas bit-identical: {same}  ({all_d[0][0][:16]}...)")
    assert same, all_d
    assert events > 0


run("adc")
run("mcmc")
dist.barrier()
dist.destroy_process_group()
