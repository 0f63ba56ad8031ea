"""CPU: the C-ABI shared library loads, exports every symbol include/davo_b200.h declares, and its
argument validation (which runs before any CUDA call) returns the documented status codes."""
import ctypes
import os
import re

import pytest
import torch

import davo_b200
from davo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "davo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(davo_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/davo_b200.h but not exported"
    assert set(_lib.EXPORTED) == set(declared)


def test_abi_version_and_strerror():
    L = _lib.lib()
    assert L.davo_abi_version() == _lib.ABI_VERSION
    assert L.davo_strerror(0) == b"ok"
    for code in range(-6, 0):
        assert len(L.davo_strerror(code)) > 3
    assert b"unknown" in L.davo_strerror(-99)


def test_descriptor_layout_matches_header():
    # 10 int32 + 4 double, no padding surprises: the C struct is 40 + 32 bytes
    assert ctypes.sizeof(_lib.ProblemDesc) == 80
    assert _lib.ProblemDesc.sufficient_decrease.offset == 48


def test_validation_status_codes_without_gpu():
    L = _lib.lib()
    d = _lib.make_desc(4, 8, 1, 9, "distort10", torch.float32)  # n must be 10
    assert L.davo_solve_calibration(ctypes.byref(d), *([None] * 12)) == -2
    d = _lib.make_desc(4, 8, 2, 20, "joint", torch.float32)  # n must be 10 + 6 V = 22
    assert L.davo_eval_cost_grad(ctypes.byref(d), *([None] * 7)) == -2
    d = _lib.make_desc(4, 8, 4, 44, "angle_ba", torch.float32)  # n must be 3 + 3 N + 6 (V - 1) = 45
    assert L.davo_solve_calibration(ctypes.byref(d), *([None] * 12)) == -2
    d = _lib.make_desc(4, 8, 1, 27, "angle_ba", torch.float32)  # a single view has no relative pose
    assert L.davo_eval_cost_grad(ctypes.byref(d), *([None] * 7)) == -2
    d = _lib.make_desc(4, 8, 4, 45, "angle_ba", torch.float32)
    assert L.davo_solve_calibration(ctypes.byref(d), *([None] * 12)) == -1  # NULL pointers
    d = _lib.make_desc(4, 8, 1, 10, 99, torch.float32)  # unknown model
    assert L.davo_solve_calibration(ctypes.byref(d), *([None] * 12)) == -3
    d = _lib.make_desc(4, 0, 1, 17, "sphere", torch.float32)  # analytic n > 16 slots
    assert L.davo_solve_calibration(ctypes.byref(d), *([None] * 12)) == -3
    d = _lib.make_desc(4, 8, 1, 10, "distort10", torch.float32)
    assert L.davo_solve_calibration(ctypes.byref(d), *([None] * 12)) == -1  # NULL pointers
    d = _lib.make_desc(0, 8, 1, 10, "distort10", torch.float32)  # empty batch is a no-op
    assert L.davo_solve_calibration(ctypes.byref(d), *([None] * 12)) == 0
    assert L.davo_line_search(ctypes.byref(d), *([None] * 10)) == 0
    assert L.davo_stage_matches(ctypes.byref(d), *([None] * 5)) == 0
    assert L.davo_bfgs_update(0, 0, 3, None, None, None, None) == 0
    assert L.davo_bfgs_update(0, 2, 0, None, None, None, None) == -2
    assert L.davo_least_squares(0, 0, 4, 2, None, None, None, None, None, None) == 0
    assert L.davo_solve_calibration(None, *([None] * 12)) == -1


def test_python_surface_has_no_cpu_fallback():
    solver = davo_b200.BFGSSolver(error_threshold=1e-6).eval()
    with pytest.raises(TypeError):  # arbitrary Python callables cannot run inside the kernel
        solver(torch.zeros(3, 2), lambda x, m: x.square().sum(-1))
    if not torch.cuda.is_available():
        with pytest.raises(_lib.DavoError):
            davo_b200.AnalyticObjective("sphere", (3,), 2)
        with pytest.raises(_lib.DavoError):
            davo_b200.compute_distorted_camera_model(torch.zeros(1, 2, 3), torch.zeros(1, 16))
        with pytest.raises(_lib.DavoError):
            davo_b200.find_error(torch.zeros(1, 1, 2, 2))


def test_solver_constructor_matches_reference_defaults():
    """autograd_solvers/bfgs_solver.py:49-78."""
    s = davo_b200.BFGSSolver()
    assert (s.sufficient_decrease, s.curvature, s.error_threshold, s.iterations, s.minimum_step, s.drop_path_p,
            s.return_second_last) == (1e-4, 0.9, 1e-4, 1000, 1e-8, 0.1, False)
    assert s.training_iterations == 1000 and s.training_error_threshold == 1e-4
    s = davo_b200.BFGSSolver(error_threshold=1e-7, training_error_threshold=1e-3, training_iterations=5)
    assert s.training_error_threshold == 1e-3 and s.training_iterations == 5
    assert s.training  # nn.Module default, as in the reference


def test_calibration_network_mirrors_reference_module_layout():
    """networks/calibration_network.py:26-52: same sub-module names and shapes (a reference checkpoint loads),
    same solver thresholds; unpack_calibration_parameters slices like calibration_pinhole_camera_model.py:33-75."""
    net = davo_b200.CalibrationNetwork(4, 8)
    assert net.num_views == 4 and net.num_points == 8
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert shapes["initial_estimator.0.weight"] == (256, 64)
    assert shapes["initial_estimator.3.weight"] == (256, 256)
    assert shapes["initial_estimator.6.weight"] == (45, 256)
    assert net.solver.error_threshold == 1e-7 and net.solver.training_error_threshold == 1e-3
    x = torch.arange(2 * 45, dtype=torch.float32).reshape(2, 45)
    intr, pts, tr, rot = davo_b200.unpack_calibration_parameters(x, 4, 8)
    assert intr.shape == (2, 1, 1, 3) and pts.shape == (2, 1, 8, 3) and tr.shape == (2, 3, 1, 3) and rot.shape == (2, 3, 1, 3)
    assert pts[1, 0, 7, 2] == 45 + 26 and tr[0, 0, 0, 0] == 27 and rot[0, 2, 0, 2] == 44
    with pytest.raises(ValueError):
        davo_b200.unpack_calibration_parameters(x, 4, 7)
