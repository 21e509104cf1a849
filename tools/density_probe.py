"""Times ADC densification and MCMC relocation at N Gaussians against the reference's libtorch
controllers (oracle/_ref) on the same inputs, and checks the deterministic parts of the result.
Test infrastructure (GPU box):  python tools/density_probe.py [N]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
import cugs_ref as ref  # noqa: E402
import cuda_gaussian_splatting_b200 as cugs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
dev = torch.device("cuda")
s = cugs.synth(n, 1920, 1080, seed=77)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
base = [t(s.positions), t(s.sh_coeffs), t(s.opacities), t(s.rotations), t(s.scales)]
g = torch.Generator(device=dev).manual_seed(1)
count = torch.randint(0, 6, (n,), device=dev, generator=g).float()
accum = torch.rand((n,), device=dev, generator=g) * 8e-4 * count.clamp_min(1) * (count > 0)
radii = torch.randint(0, 41, (n,), device=dev, generator=g).float()
base[2][torch.rand((n,), device=dev, generator=g) < 0.05] = -7.0
extent = float(torch.exp(base[4]).max(dim=1).values.median()) / 0.01
cfg = cugs.DensificationConfig()


def timed(fn, reps=3):
    best, out = 1e30, None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, out


def mine():
    m = cugs.GaussianModel(*(x.clone() for x in base))
    c = cugs.DensificationController(cfg, extent, n, dev)
    c.grad_accum.copy_(accum); c.grad_count.copy_(count); c.max_radii_2d.copy_(radii)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = c.densify(m, 3100)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, m, r


def theirs():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = ref.densify(*base, accum, count, radii, extent, 3100,
                      [cfg.grad_threshold, cfg.opacity_threshold, cfg.percent_dense, cfg.max_screen_size, 0,
                       cfg.opacity_reset_every])
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, out


torch.cuda.reset_peak_memory_stats()
m0 = torch.cuda.memory_allocated()
tm = min(mine()[0] for _ in range(3))
_, model, res = mine()
peak_mine = torch.cuda.max_memory_allocated() - m0
torch.cuda.reset_peak_memory_stats()
m0 = torch.cuda.memory_allocated()
tr = min(theirs()[0] for _ in range(3))
_, rout = theirs()
peak_ref = torch.cuda.max_memory_allocated() - m0
rst = [int(x) for x in rout[5].tolist()]
head = res.num_after - 2 * res.num_split
same = all(torch.equal(a[:head].view(torch.int32), b[:head].view(torch.int32))
           for a, b in zip((model.positions, model.sh_coeffs, model.opacities, model.rotations, model.scales), rout[:5]))
print(f"densify N={n}: cloned {res.num_cloned} split {res.num_split} pruned {res.num_pruned} -> {res.num_after}; "
      f"reference counts {rst}; kept+cloned rows bit-identical: {same}")
print(f"  this library {tm:.2f} ms (incl. model clone-free host policy, one sync), peak extra memory {peak_mine / 2**20:.0f} MiB")
print(f"  reference    {tr:.2f} ms (incl. its input clones), peak extra memory {peak_ref / 2**20:.0f} MiB")

mc = cugs.MCMCConfig()
def mine_rel():
    m = cugs.GaussianModel(*(x.clone() for x in base))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = cugs.mcmc_relocate(m, 1000, mc, 5.0)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, st
def ref_rel():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = ref.mcmc_relocate(*base, 5.0, mc.dead_opacity_threshold, mc.relocate_cap)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, out
tm = min(mine_rel()[0] for _ in range(3)); st = mine_rel()[1]
try:
    tr = min(ref_rel()[0] for _ in range(3)); ro = ref_rel()[1]
    print(f"relocate N={n}: dead {st.num_dead} relocated {st.num_relocated}; reference {[int(x) for x in ro[5].tolist()]}")
    print(f"  this library {tm:.2f} ms, reference {tr:.2f} ms (incl. its input clones)")
except RuntimeError as e:  # torch::multinomial: "number of categories cannot exceed 2^24"
    print(f"relocate N={n}: dead {st.num_dead} relocated {st.num_relocated}; this library {tm:.2f} ms; "
          f"the reference controller FAILS at this size: {str(e).splitlines()[0]}")
