"""ctypes binding of the C ABI declared in include/cugs_b200.h.

The library is hand-written CUDA for sm_100a. There is NO CPU fallback: if the shared object is
missing or the device is not a B200-class GPU, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libcugs_b200.so"


class CugsView(C.Structure):
    """struct cugs_view (include/cugs_b200.h) — CameraInfo + RenderSettings of the reference."""

    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("fx", C.c_float),
        ("fy", C.c_float),
        ("cx", C.c_float),
        ("cy", C.c_float),
        ("view", C.c_float * 16),
        ("cam_center", C.c_float * 3),
        ("bg", C.c_float * 3),
        ("active_sh_degree", C.c_int32),
        ("num_coeffs", C.c_int32),
        ("scale_modifier", C.c_float),
    ]


class CugsDensifyConfig(C.Structure):  # cugs_densify_config_t
    _fields_ = [
        ("grad_threshold", C.c_float),
        ("size_threshold", C.c_float),
        ("opacity_threshold", C.c_float),
        ("apply_size_pruning", C.c_int32),
        ("max_screen_size", C.c_float),
        ("ws_threshold", C.c_float),
    ]


class CugsTrainConfig(C.Structure):  # cugs_train_config_t
    _fields_ = [
        ("lambda_ssim", C.c_float), ("max_sh_degree", C.c_int32), ("background", C.c_float * 3),
        ("lr_position_init", C.c_float), ("lr_position_final", C.c_float), ("lr_position_max_steps", C.c_int32),
        ("lr_sh_coeffs", C.c_float), ("lr_opacities", C.c_float), ("lr_scales", C.c_float), ("lr_rotations", C.c_float),
        ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("accumulate_stats", C.c_int32), ("mcmc", C.c_int32),
        ("lambda_opacity", C.c_float), ("lambda_scale", C.c_float),
        ("noise_lr_init", C.c_float), ("noise_lr_final", C.c_float), ("noise_lr_max_steps", C.c_int32),
        ("noise_gate_k", C.c_float), ("noise_gate_t", C.c_float), ("noise_seed", C.c_uint64),
        ("frames_in_flight", C.c_int32), ("use_graph", C.c_int32),
    ]


class CugsTrainTensors(C.Structure):  # cugs_train_tensors_t
    _fields_ = [
        ("params", C.c_void_p * 5), ("adam_m", C.c_void_p * 5), ("adam_v", C.c_void_p * 5), ("grads", C.c_void_p * 5),
        ("dL_dmeans_2d", C.c_void_p), ("grad_accum", C.c_void_p), ("grad_count", C.c_void_p), ("max_radii", C.c_void_p),
        ("touch_mask", C.c_void_p),
    ]



_P = C.c_void_p
_I64 = C.c_int64
_SZ = C.c_size_t
_INT = C.c_int
_F = C.c_float
_VP = C.POINTER(CugsView)

# name -> (restype, argtypes); must list EVERY symbol include/cugs_b200.h declares
SIGNATURES = {
    "cugs_b200_create": (_INT, [_INT, C.POINTER(_P)]),
    "cugs_b200_destroy": (None, [_P]),
    "cugs_b200_last_error": (C.c_char_p, [_P]),
    "cugs_b200_abi_version": (_INT, []),
    "cugs_b200_sm_count": (_INT, [_P]),
    "cugs_b200_launch_count": (C.c_uint64, [_P]),
    "cugs_b200_device_info": (_INT, [_P, C.POINTER(_INT), C.POINTER(_INT), C.POINTER(_INT)]),
    "cugs_b200_preprocess_fwd": (_INT, [_P, _P, _I64, _VP] + [_P] * 14),
    "cugs_b200_sh_forward": (_INT, [_P, _P, _I64, _INT, _INT, _P, _P, _P]),
    "cugs_b200_sh_backward": (_INT, [_P, _P, _I64, _INT, _INT, _P, _P, _P, _P]),
    "cugs_b200_scan_temp_bytes": (_SZ, [_I64]),
    "cugs_b200_scan": (_INT, [_P, _P, _I64, _P, _P, _P, C.POINTER(_I64), _P, _SZ]),
    "cugs_b200_duplicate_with_keys": (_INT, [_P, _P, _I64, _INT, _INT, _P, _P, _P, _P, _P, _I64, _P, _P]),
    "cugs_b200_sort_temp_bytes": (_SZ, [_I64]),
    "cugs_b200_sort_pairs": (_INT, [_P, _P, _I64, _INT, _INT, _P, _P, _P, _P, _P, _SZ]),
    "cugs_b200_sort_packed_passes": (_INT, [_INT]),
    "cugs_b200_sort_packed_temp_bytes": (_SZ, [_I64, _INT, _INT]),
    "cugs_b200_sort_packed": (_INT, [_P, _P, _I64, _INT, _P, _P, _P, _INT, _P, _P, _SZ, _P]),
    "cugs_b200_tile_ranges": (_INT, [_P, _P, _I64, _P, _INT, _P]),
    "cugs_b200_blend_fwd": (_INT, [_P, _P, _VP] + [_P] * 10),
    "cugs_b200_blend_bwd": (_INT, [_P, _P, _I64, _VP] + [_P] * 15),
    "cugs_b200_count_evaluations": (_INT, [_P, _P, _VP] + [_P] * 7),
    "cugs_b200_preprocess_bwd": (_INT, [_P, _P, _I64, _VP] + [_P] * 19),
    "cugs_b200_render_workspace_bytes": (_SZ, [_I64, _I64]),
    "cugs_b200_render_forward": (_INT, [_P, _P, _I64, _I64, _VP] + [_P] * 16 + [_P, _SZ, _P, _SZ, _P]),
    "cugs_b200_render_plan": (_INT, [_P, _P, _I64, _VP] + [_P] * 11 + [_P, _SZ, C.POINTER(_I64)]),
    "cugs_b200_render_pair_scratch_bytes": (_SZ, [_I64]),
    "cugs_b200_render_finish": (_INT, [_P, _P, _I64, _I64, _VP] + [_P] * 11 + [_P, _SZ, _P, _SZ]),
    "cugs_b200_render_backward": (_INT, [_P, _P, _I64, _VP] + [_P] * 25 + [_INT, _P, _SZ]),
    "cugs_b200_compact_grad_floats": (_I64, [_I64, _INT]),
    "cugs_b200_gather_grad_rows": (_INT, [_P, _P, _I64, _INT, _P, _P, _I64, C.POINTER(_P), _P, _P, _P, _P]),
    "cugs_b200_scatter_grad_rows": (_INT, [_P, _P, _I64, _INT, _P, _P, _I64, _P, C.POINTER(_P), _P, _P]),
    "cugs_b200_build_touch_index": (_INT, [_P, _P, _I64, _P, _P, _P]),
    "cugs_b200_p2p_reduce_masks": (_INT, [_P, _P, _I64, _INT, _INT, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _P, _P, _P]),
    "cugs_b200_p2p_reduce_rows": (_INT, [_P, _P, _I64, _INT, _INT, _INT, _P, _P, C.POINTER(_P), C.POINTER(_P)]),
    "cugs_b200_last_sort_plan": (_INT, [_P, C.POINTER(_INT), C.POINTER(_INT)]),
    "cugs_b200_set_stage_timing": (_INT, [_P, _INT]),
    "cugs_b200_get_stage_ms": (_INT, [_P, C.POINTER(_F)]),
    "cugs_b200_loss_workspace_bytes": (_SZ, [_INT, _INT]),
    "cugs_b200_loss_l1_ssim": (_INT, [_P, _P, _INT, _INT, _F, _INT, _P, _P, _P, _P, _P, _SZ, _P]),
    "cugs_b200_adam_step": (_INT, [_P, _P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P),
                                    C.POINTER(_I64), C.POINTER(_F), _F, _F, _F, _F, _F, _F]),
    "cugs_b200_adam_step_mcmc": (_INT, [_P, _P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P),
                                         C.POINTER(_I64), C.POINTER(_F), _F, _F, _F, _F, _F, _F, _F, _F]),
    "cugs_b200_mcmc_inject_noise": (_INT, [_P, _P, _I64, _P, _P, _P, _F, _F, _F, C.c_uint64, C.c_uint32, _P]),
    "cugs_b200_trainer_workspace_bytes": (_SZ, [_I64, _INT, _INT, _INT, _I64, _INT]),
    "cugs_b200_trainer_create": (_INT, [_P, _I64, _INT, _INT, _INT, _I64, C.POINTER(CugsTrainConfig),
                                         C.POINTER(CugsTrainTensors), _P, _SZ, C.POINTER(_P)]),
    "cugs_b200_trainer_destroy": (None, [_P]),
    "cugs_b200_trainer_set_views": (_INT, [_P, _INT, _VP, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _INT]),
    "cugs_b200_trainer_step": (_INT, [_P, _P, _INT, _INT]),
    "cugs_b200_trainer_result": (_INT, [_P, _P, C.POINTER(_F), C.POINTER(_I64)]),
    "cugs_b200_trainer_set_adam_steps": (_INT, [_P, _I64]),
    "cugs_b200_trainer_adam_steps": (_I64, [_P]),
    "cugs_b200_accumulate_stats": (_INT, [_P, _P, _I64, _P, _P, _P, _P, _P]),
    "cugs_b200_densify_temp_bytes": (_SZ, [_I64]),
    "cugs_b200_densify_classify": (_INT, [_P, _P, _I64, _P, _P, _P, _P, _P, C.POINTER(CugsDensifyConfig), _P,
                                           C.POINTER(_I64), _P, _SZ]),
    "cugs_b200_densify_apply": (_INT, [_P, _P, _I64, _I64, _INT, _P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P),
                                        C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.c_uint64, _P, _P, _SZ]),
    "cugs_b200_mcmc_relocate_temp_bytes": (_SZ, [_I64]),
    "cugs_b200_mcmc_relocate": (_INT, [_P, _P, _I64, _INT, _P, _P, _P, _P, _P, _F, _I64, _F, C.c_uint64, C.c_uint32,
                                        _P, _P, C.POINTER(_I64), _P, _SZ]),
}

_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree library and bind every declared symbol (no GPU needed for this)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("CUGS_B200_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: the sm_100a CUDA library has not been built (run ./build.sh or "
            "__graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    header = _HERE.parent / "include" / "cugs_b200.h"
    if header.exists():  # a stale .so (older signatures) must fail loudly, not corrupt arguments
        import re
        want = int(re.search(r"#define CUGS_B200_ABI_VERSION (\d+)", header.read_text()).group(1))
        if lib.cugs_b200_abi_version() != want:
            raise RuntimeError(f"{path} has ABI version {lib.cugs_b200_abi_version()}, include/cugs_b200.h declares "
                               f"{want}: rebuild with ./build.sh")
    _lib = lib
    return lib


class CugsError(RuntimeError):
    """Raised for a non-zero status of the C ABI (mirrors the reference's c10::Error /
    std::runtime_error("CUDA error ...") behaviour, utils/cuda_utils.cuh:12-20)."""


_handles: dict[int, int] = {}


def handle(device_index: int) -> int:
    """One handle per device (created lazily)."""
    h = _handles.get(device_index)
    if h is None:
        lib = load_library()
        out = _P()
        st = lib.cugs_b200_create(int(device_index), C.byref(out))
        if st != 0:
            raise CugsError(f"cugs_b200_create(device={device_index}) failed with status {st} "
                            "(-5 = not an sm_100 device; there is no fallback path)")
        h = out.value
        _handles[device_index] = h
    return h


def check(h: int, status: int, what: str) -> None:
    if status != 0:
        msg = load_library().cugs_b200_last_error(h)
        raise CugsError(f"{what} failed (status {status}): {msg.decode() if msg else ''}")
