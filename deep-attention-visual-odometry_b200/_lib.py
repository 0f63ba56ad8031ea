"""ctypes binding of libdavo_b200.so (the C-ABI in include/davo_b200.h).

The library is built in-tree by ``make -C deep-attention-visual-odometry_b200/csrc`` (or
``__graft_entry__.build()``).  There is no fallback: if the library is missing, or a compute entry
point is reached without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DAVO_B200_LIB selects an A/B build of the same ABI (tools/); the product library is the default
LIB_PATH = os.environ.get("DAVO_B200_LIB") or os.path.join(_HERE, "libdavo_b200.so")
CSRC = os.path.join(_HERE, "csrc")

ABI_VERSION = 2
WORKSPACE_BYTES = 256
F32, F64 = 0, 1

MODEL_IDS = {"distort10": 0, "joint": 1, "angle_ba": 2, "sphere": 16, "sphere_offset": 17, "log_sphere": 18,
             "rosenbrock": 19, "cosine": 20, "x2_sine": 21, "distance": 22}
REASON_NAMES = {0: "threshold", 1: "step", 2: "cap", 3: "nan", 5: "dropped"}

# every symbol include/davo_b200.h declares (tests check the library exports all of them)
EXPORTED = ("davo_abi_version", "davo_strerror", "davo_launch_count", "davo_mlp_packed_bytes", "davo_solve_calibration",
            "davo_solve_training", "davo_solve_backward", "davo_eval_cost_grad", "davo_line_search", "davo_stage_matches", "davo_project",
            "davo_project_jacobian", "davo_least_squares", "davo_bfgs_update", "davo_bfgs_initial_scale",
            "davo_interpolate_alpha", "davo_interpolate_alpha_backward",
            "davo_mlp_pack_weights", "davo_mlp_forward",
            "davo_generate_distort10", "davo_generate_joint", "davo_generate_views_and_points")


class ProblemDesc(ctypes.Structure):
    """davo_problem_desc (include/davo_b200.h)."""

    _fields_ = [(k, ctypes.c_int32) for k in
                ("B", "N", "V", "n", "model", "dtype", "max_iters", "max_ls_iters", "strong", "has_weights",
                 "zoom_interpolation", "reserved0")] + \
               [(k, ctypes.c_double) for k in ("sufficient_decrease", "curvature", "error_threshold", "minimum_step")]


class TrainingDesc(ctypes.Structure):
    """davo_training_desc (include/davo_b200.h)."""

    _fields_ = [("capacity", ctypes.c_int32), ("return_second_last", ctypes.c_int32),
                ("drop_path_p", ctypes.c_double), ("seed", ctypes.c_uint64), ("hvp_rel_step", ctypes.c_double)]


class MlpDesc(ctypes.Structure):
    """davo_mlp_desc (include/davo_b200.h)."""

    _fields_ = [(k, ctypes.c_int32) for k in ("B", "in_features", "hidden", "out_features")]


class GeneratorDesc(ctypes.Structure):
    """davo_generator_desc (include/davo_b200.h)."""

    _fields_ = [(k, ctypes.c_int32) for k in ("B", "N", "V", "dtype", "ill_conditioned", "random_pose")] + \
               [("seed", ctypes.c_uint64), ("first_problem", ctypes.c_uint64)] + \
               [(k, ctypes.c_double) for k in ("fov", "noise", "pathological", "start_noise", "min_camera_distance")]


class DavoError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into libdavo_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j4"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise DavoError("building libdavo_b200.so failed")
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DavoError(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC}` or __graft_entry__.build(). "
                "This package has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        dp = ctypes.POINTER(ProblemDesc)
        L.davo_abi_version.restype = ctypes.c_int
        L.davo_strerror.restype = ctypes.c_char_p
        L.davo_strerror.argtypes = [ctypes.c_int]
        L.davo_launch_count.restype = i64
        L.davo_solve_calibration.argtypes = [dp] + [vp] * 12
        tp = ctypes.POINTER(TrainingDesc)
        L.davo_solve_training.argtypes = [dp, tp] + [vp] * 16
        L.davo_solve_backward.argtypes = [dp, tp] + [vp] * 15
        L.davo_eval_cost_grad.argtypes = [dp] + [vp] * 7
        L.davo_line_search.argtypes = [dp] + [vp] * 10
        L.davo_stage_matches.argtypes = [dp] + [vp] * 5
        L.davo_project.argtypes = [dp] + [vp] * 5
        L.davo_project_jacobian.argtypes = [dp] + [vp] * 6
        L.davo_least_squares.argtypes = [i32] * 4 + [vp] * 6
        L.davo_bfgs_update.argtypes = [i32] * 3 + [vp] * 4
        L.davo_bfgs_initial_scale.argtypes = [i32] * 3 + [vp] * 4
        L.davo_interpolate_alpha.argtypes = [i32, i64] + [vp] * 6
        L.davo_interpolate_alpha_backward.argtypes = [i32, i64] + [vp] * 10
        L.davo_mlp_packed_bytes.argtypes = [i32, i32]
        L.davo_mlp_packed_bytes.restype = i64
        L.davo_mlp_pack_weights.argtypes = [i32, i32] + [vp] * 3
        L.davo_mlp_forward.argtypes = [ctypes.POINTER(MlpDesc)] + [vp] * 13
        gp = ctypes.POINTER(GeneratorDesc)
        L.davo_generate_distort10.argtypes = [gp] + [vp] * 6
        L.davo_generate_joint.argtypes = [gp] + [vp] * 5
        L.davo_generate_views_and_points.argtypes = [gp] + [vp] * 9
        for name in EXPORTED[4:]:
            getattr(L, name).restype = ctypes.c_int
        if L.davo_abi_version() != ABI_VERSION:
            raise DavoError(f"libdavo_b200.so has ABI {L.davo_abi_version()}, expected {ABI_VERSION}")
        _lib = L
    return _lib


def check(status: int, what: str) -> None:
    """Map a davo_status to the exception the reference would raise for the same mistake."""
    if status == 0:
        return
    msg = f"{what}: {lib().davo_strerror(status).decode()} (status {status})"
    if status in (-2, -6):
        raise ValueError(msg)  # reference: ValueError on a bad parameter width
    if status == -3:
        raise NotImplementedError(msg)
    raise DavoError(msg)


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise DavoError("no CUDA device: the calibration solve runs only on the GPU (sm_100a); "
                        "there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.float64:
        return F64
    raise TypeError(f"unsupported dtype {dtype}: the solve kernels compute in float32 or float64")


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(lib().davo_launch_count())


def make_desc(B, N, V, n, model, dtype, *, iterations=1000, max_ls_iters=1000, strong=True, has_weights=False,
              sufficient_decrease=1e-4, curvature=0.9, error_threshold=1e-4, minimum_step=1e-8,
              zoom_interpolation=False) -> ProblemDesc:
    model_id = MODEL_IDS[model] if isinstance(model, str) else int(model)
    return ProblemDesc(int(B), int(N), int(V), int(n), model_id, dtype_code(dtype), int(iterations),
                       int(max_ls_iters), int(bool(strong)), int(bool(has_weights)), int(bool(zoom_interpolation)), 0,
                       float(sufficient_decrease),
                       float(curvature), float(error_threshold), float(minimum_step))
