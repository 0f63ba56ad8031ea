"""B200-native batched calibration solve (one hot path of jskinn/deep-attention-visual-odometry).

Host-side mirror of the reference's Python interface for that path; every numeric entry point
dispatches to hand-written sm_100a CUDA through the C-ABI in include/davo_b200.h.  There is no CPU
fallback: calling a compute function without the built library or without a CUDA device raises.
"""
from . import synthetic  # noqa: F401
from . import _lib  # noqa: F401
from .objectives import (AnalyticObjective, AngleDistanceObjective, CalibrationObjective, DistortionObjective,
                         JointPoseObjective)
from .solvers import BFGSSolver, PendingSolve, SolveInfo, interpolate_alpha, line_search_wolfe_conditions
from .camera_model import compute_distorted_camera_model, compute_distorted_camera_model_and_jacobian
from .base_types import CameraViewsAndPoints
from .networks import CalibrationNetwork, unpack_calibration_parameters
from .least_squares_utils import find_error, find_error_gradient, find_residuals

__all__ = [
    "interpolate_alpha",
    "PendingSolve",
    "AnalyticObjective", "AngleDistanceObjective", "BFGSSolver", "CalibrationNetwork", "CameraViewsAndPoints", "CalibrationObjective", "DistortionObjective", "JointPoseObjective",
    "SolveInfo", "compute_distorted_camera_model", "compute_distorted_camera_model_and_jacobian", "find_error",
    "find_error_gradient", "find_residuals", "line_search_wolfe_conditions", "synthetic", "unpack_calibration_parameters",
]
