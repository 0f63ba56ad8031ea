"""ctypes binding of the C oracle (oracle/calib_oracle.c) — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
Arrays are numpy (host); the dtype of `x0`/`x` picks the float32 or float64 instantiation.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdavo_oracle.so")

MODEL_IDS = {"distort10": 0, "joint": 1, "angle_ba": 2, "sphere": 16, "sphere_offset": 17, "log_sphere": 18,
             "rosenbrock": 19, "cosine": 20, "x2_sine": 21, "distance": 22}
REASONS = {0: "threshold", 1: "step", 2: "cap", 3: "nan"}


class ProblemDesc(ctypes.Structure):
    """davo_problem_desc, include/davo_b200.h."""

    _fields_ = [(k, ctypes.c_int32) for k in
                ("B", "N", "V", "n", "model", "dtype", "max_iters", "max_ls_iters", "strong", "has_weights",
                 "zoom_interpolation", "reserved0")] + \
               [(k, ctypes.c_double) for k in ("sufficient_decrease", "curvature", "error_threshold", "minimum_step")]


def make_desc(B, N, V, n, model, dtype, *, iterations=1000, max_ls_iters=1000, strong=True, has_weights=False,
              sufficient_decrease=1e-4, curvature=0.9, error_threshold=1e-4, minimum_step=1e-8,
              zoom_interpolation=False) -> ProblemDesc:
    model_id = MODEL_IDS[model] if isinstance(model, str) else int(model)
    return ProblemDesc(int(B), int(N), int(V), int(n), model_id, 1 if np.dtype(dtype) == np.float64 else 0,
                       int(iterations), int(max_ls_iters), int(bool(strong)), int(bool(has_weights)),
                       int(bool(zoom_interpolation)), 0,
                       float(sufficient_decrease), float(curvature), float(error_threshold), float(minimum_step))


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
            for f in ("calib_oracle.c", "calib_oracle_impl.h")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _suffix(dtype):
    return "f64" if np.dtype(dtype) == np.float64 else "f32"


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def max_threads() -> int:
    return int(lib().davo_oracle_max_threads())


def solve(model, x0, data0=None, data1=None, weights=None, *, N=0, V=1, threads=0, **solver_kwargs):
    """Run the scalar restatement of BFGSSolver.eval() on every row of x0 [B,n].
    Returns dict(x, cost, converged, iters, fevals, reason)."""
    dt = x0.dtype
    B, n = x0.shape
    d = make_desc(B, N, V, n, model, dt, has_weights=weights is not None, **solver_kwargs)
    data0, data1, weights, x0 = _c(data0, dt), _c(data1, dt), _c(weights, dt), _c(x0, dt)
    x = np.empty_like(x0)
    cost = np.empty(B, dt)
    conv = np.empty(B, np.uint8)
    iters = np.empty(B, np.int32)
    fevals = np.empty(B, np.int32)
    reason = np.empty(B, np.int32)
    st = getattr(lib(), "davo_oracle_solve_" + _suffix(dt))(
        ctypes.byref(d), _p(data0), _p(data1), _p(weights), _p(x0), _p(x), _p(cost), _p(conv), _p(iters),
        _p(fevals), _p(reason), ctypes.c_int(int(threads)))
    if st != 0:
        raise ValueError(f"davo_oracle_solve failed with status {st}")
    return dict(x=x, cost=cost, converged=conv.astype(bool), iters=iters, fevals=fevals, reason=reason)


def eval_cost_grad(model, x, data0=None, data1=None, weights=None, *, N=0, V=1, want_grad=True):
    dt = x.dtype
    B, n = x.shape
    d = make_desc(B, N, V, n, model, dt, has_weights=weights is not None)
    data0, data1, weights, x = _c(data0, dt), _c(data1, dt), _c(weights, dt), _c(x, dt)
    cost = np.empty(B, dt)
    grad = np.empty((B, n), dt) if want_grad else None
    st = getattr(lib(), "davo_oracle_eval_" + _suffix(dt))(
        ctypes.byref(d), _p(data0), _p(data1), _p(weights), _p(x), _p(cost), _p(grad))
    if st != 0:
        raise ValueError(f"davo_oracle_eval failed with status {st}")
    return cost, grad


def line_search(model, x, direction, base_cost, base_grad, data0=None, data1=None, weights=None, *, N=0, V=1,
                sufficient_decrease=1e-4, curvature=0.9, strong=False, max_ls_iters=1000, zoom_interpolation=False):
    dt = x.dtype
    B, n = x.shape
    d = make_desc(B, N, V, n, model, dt, has_weights=weights is not None, strong=strong,
                  sufficient_decrease=sufficient_decrease, curvature=curvature, max_ls_iters=max_ls_iters,
                  zoom_interpolation=zoom_interpolation)
    args = [_c(a, dt) for a in (data0, data1, weights, x, direction, base_cost, base_grad)]
    alpha = np.empty(B, dt)
    fevals = np.empty(B, np.int32)
    st = getattr(lib(), "davo_oracle_line_search_" + _suffix(dt))(
        ctypes.byref(d), *[_p(a) for a in args], _p(alpha), _p(fevals))
    if st != 0:
        raise ValueError(f"davo_oracle_line_search failed with status {st}")
    return alpha, fevals


def interpolate_alpha(alpha_1, alpha_2, value_1, value_2, grad_out=None):
    """utils/func_interpolate_alpha.py: candidate (and, with grad_out, the four input gradients)."""
    dt = alpha_1.dtype
    a1, a2, v1, v2 = (_c(np.ravel(a), dt) for a in (alpha_1, alpha_2, value_1, value_2))
    out = np.empty_like(a1)
    go = None if grad_out is None else _c(np.ravel(grad_out), dt)
    grads = [np.empty_like(a1) for _ in range(4)] if go is not None else [None] * 4
    fn = getattr(lib(), "davo_oracle_interpolate_alpha_" + _suffix(dt))
    fn(ctypes.c_longlong(a1.size), _p(a1), _p(a2), _p(v1), _p(v2), _p(out), _p(go), *[_p(g) for g in grads])
    out = out.reshape(alpha_1.shape)
    if go is None:
        return out
    return (out,) + tuple(g.reshape(alpha_1.shape) for g in grads)


def bfgs_update(H, s, y):
    dt = H.dtype
    k, n = s.shape
    H = np.array(H, dtype=dt, order="C", copy=True)
    s, y = _c(s, dt), _c(y, dt)
    getattr(lib(), "davo_oracle_bfgs_update_" + _suffix(dt))(ctypes.c_int(k), ctypes.c_int(n), _p(H), _p(s), _p(y))
    return H


def bfgs_initial_scale(s, y):
    dt = s.dtype
    k, n = s.shape
    s, y = _c(s, dt), _c(y, dt)
    out = np.empty(k, dt)
    getattr(lib(), "davo_oracle_bfgs_initial_scale_" + _suffix(dt))(
        ctypes.c_int(k), ctypes.c_int(n), _p(s), _p(y), _p(out))
    return out


def project(points_3d, params16, jacobian=False):
    dt = params16.dtype
    B, N, _ = points_3d.shape
    pts, th = _c(points_3d, dt), _c(params16, dt)
    u = np.empty((B, N), dt)
    v = np.empty((B, N), dt)
    J = np.empty((B, 2 * N, 16), dt) if jacobian else None
    getattr(lib(), "davo_oracle_project_" + _suffix(dt))(
        ctypes.c_int(B), ctypes.c_int(N), _p(pts), _p(th), _p(u), _p(v), _p(J))
    return (J, u, v) if jacobian else (u, v)


def stage(points_3d, obs, pose=None):
    dt = points_3d.dtype
    B, N, _ = points_3d.shape
    pts, ob, ps = _c(points_3d, dt), _c(obs, dt), _c(pose, dt)
    out = np.empty((B, N, 4), dt)
    getattr(lib(), "davo_oracle_stage_" + _suffix(dt))(ctypes.c_int(B), ctypes.c_int(N), _p(pts), _p(ob), _p(ps), _p(out))
    return out


def least_squares(residuals, jacobian=None, weights=None):
    dt = residuals.dtype
    B = residuals.shape[0]
    res = _c(residuals, dt).reshape(B, -1)
    Rn = res.shape[1]
    jac = None if jacobian is None else _c(jacobian, dt).reshape(B, Rn, -1)
    P = 0 if jac is None else jac.shape[2]
    # point_weights is BxFxNx1 in the reference and broadcasts over the trailing (u, v) axis
    w = None if weights is None else _c(np.broadcast_to(weights, residuals.shape), dt).reshape(B, Rn)
    err = np.empty(B, dt)
    grad = np.empty((B, P), dt) if jac is not None else None
    getattr(lib(), "davo_oracle_least_squares_" + _suffix(dt))(
        ctypes.c_int(B), ctypes.c_int(Rn), ctypes.c_int(P), _p(res), _p(jac), _p(w), _p(err), _p(grad))
    return err, grad


def solve_batch(batch, threads=0, **solver_kwargs):
    """Convenience: run the oracle on a davo_b200.synthetic.CalibrationBatch."""
    if batch.model == "distort10":
        staged = stage(batch.points_3d, batch.obs, batch.pose)
        return solve("distort10", batch.x0, staged, N=batch.N, V=1, threads=threads, **solver_kwargs)
    if batch.model == "joint":
        return solve("joint", batch.x0, batch.points_3d, batch.obs, N=batch.N, V=batch.views, threads=threads,
                     **solver_kwargs)
    if batch.model == "angle_ba":
        return solve("angle_ba", batch.x0, batch.obs, None, batch.weights, N=batch.N, V=batch.views, threads=threads,
                     **solver_kwargs)
    raise ValueError(batch.model)
