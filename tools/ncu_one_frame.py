"""One forward + backward frame of workload B (3 M Gaussians, SH 3, 1920x1080), twice: the second frame is the
one to capture, e.g.

    ncu --section SpeedOfLight --section InstructionStats --section WarpStateStats --section SchedulerStats \
        --section LaunchStats --section Occupancy --clock-control none -k regex:k_blend_bwd -s 1 -c 1 \
        --csv --page raw --log-file gpurun_out/ncu_blend_bwd.csv python tools/ncu_one_frame.py

Nothing here is a benchmark number: it only puts the hot kernels of one frame in front of the profiler without the
minutes of replay a whole bench.py run costs under ncu."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cuda_gaussian_splatting_b200 as cugs  # noqa: E402


def main():
    n, W, H = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (3_000_000, 1920, 1080)))
    dev = torch.device("cuda:0")
    scene = cugs.synth(n, W, H, seed=1236)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    model = cugs.GaussianModel(t(scene.positions), t(scene.sh_coeffs), t(scene.opacities), t(scene.rotations),
                               t(scene.scales))
    settings = cugs.RenderSettings((0.0, 0.0, 0.0), 3, 1.0)
    dL = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, size=(H, W, 3)).astype(np.float32)).to(dev)
    for _ in range(2):
        out = cugs.render(model, scene.camera, settings)
        cugs.render_backward(dL, out, model, scene.camera, settings)
    torch.cuda.synchronize()
    print("frames done, P =", int(out.gaussian_indices.numel()))


if __name__ == "__main__":
    main()
