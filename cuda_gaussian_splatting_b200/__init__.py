"""cuda_gaussian_splatting_b200 — B200-native (sm_100a) differentiable Gaussian-splatting rasterizer.

Hot path of Artemarius/cuda-gaussian-splatting (``cugs::render`` / ``cugs::render_backward`` plus
the fused L1+SSIM loss and fused Adam either side of it), rebuilt from scratch as hand-written
CUDA behind the C ABI of ``include/cugs_b200.h``. This package is the host-side mirror of the
reference's operator interface; all compute runs in ``libcugs_b200.so``. No CPU fallback.
"""
from ._lib import CugsError, LIB_PATH, load_library  # noqa: F401
from .rasterizer import (  # noqa: F401
    BackwardOutput, CameraInfo, ForwardOutput, FrameBuffers, GaussianModel, ProjectionBackwardOutput,
    ProjectionOutput, RasterizeBackwardOutput, RenderOutput, RenderSettings, SortingOutput,
    evaluate_sh_backward_cuda, evaluate_sh_cuda, project_backward, project_gaussians, rasterize_backward,
    rasterize_forward, render, render_backward, render_image, ImageBuffers, sort_gaussians, count_evaluations,
)
from .training import (  # noqa: F401
    AdamConfig, DensificationStats, FusedAdam, MCMCConfig, PositionLRConfig, mcmc_inject_noise, mcmc_noise_lr, SyntheticTrainer, TargetUploader, TrainConfig, active_sh_degree_for_step, combined_loss,
    combined_loss_with_grad, l1_loss, position_lr, ssim, ssim_loss, ssim_mean,
)
from .synth import Scene, default_camera, ring_cameras, synth  # noqa: F401
from .ply_io import read_gaussian_ply, write_gaussian_ply  # noqa: F401
from .density import (  # noqa: F401
    DensificationConfig, DensificationController, DensificationResult, MCMCStats, mcmc_relocate,
    mcmc_should_relocate,
)
from .native_trainer import NativeTrainer  # noqa: F401
from .parallel import (  # noqa: F401
    allreduce_step, arena_layout, fold_step_stats, grad_scale_for, shard_views, sparse_allreduce_step, MaskOverlap, P2PExchange,
)

__all__ = [n for n in dir() if not n.startswith("_")]
