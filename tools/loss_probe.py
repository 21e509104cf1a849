"""Times the fused L1+SSIM loss (+gradient) at a given image size with CUDA events."""
import sys
import torch
sys.path.insert(0, ".")
import cuda_gaussian_splatting_b200 as cugs

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
g = torch.Generator(device="cuda").manual_seed(0)
xs = [torch.rand((H, W, 3), device="cuda", generator=g) for _ in range(8)]   # 8 x 25 MB pairs: L2 does not hold them all
ys = [torch.rand((H, W, 3), device="cuda", generator=g) for _ in range(8)]
for i in range(5):
    cugs.combined_loss_with_grad(xs[i % 8], ys[i % 8], 0.2)
torch.cuda.synchronize()
iters = 64
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    cugs.combined_loss_with_grad(xs[i % 8], ys[i % 8], 0.2)
e1.record()
torch.cuda.synchronize()
print(f"loss+grad {W}x{H}: {e0.elapsed_time(e1) / iters * 1000:.1f} us per call")
