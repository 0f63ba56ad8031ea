"""Instrumented runs of the UNMODIFIED reference — TEST INFRASTRUCTURE, build container only.

Imports /root/reference (read-only) with the constants-only `spatial_maths` shim ahead of it on
sys.path and drives the reference's own `BFGSSolver(...).eval()`
(autograd_solvers/bfgs_solver.py:26-215), `line_search_wolfe_conditions`
(autograd_solvers/line_search/wolfe_conditions.py:23-239) and `compute_distorted_camera_model`
(camera_model/distorted_camera_model.py:106-111) on this repo's synthetic inputs.  It exists to
(1) pin the C oracle and (2) write the golden fixtures under tests/golden/ (oracle/make_golden.py).
/root/reference does not exist on the GPU box, so nothing that runs there may import this module.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("DAVO_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")

sys.dont_write_bytecode = True  # the reference tree is read-only


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "deep_attention_visual_odometry"))


def _import_reference():
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    from deep_attention_visual_odometry.autograd_solvers import bfgs_solver as bfgs_mod
    from deep_attention_visual_odometry.autograd_solvers.line_search import wolfe_conditions as wolfe_mod
    from deep_attention_visual_odometry.camera_model import distorted_camera_model as cam_mod
    from deep_attention_visual_odometry.solvers import least_squares_utils as lsq_mod
    return bfgs_mod, wolfe_mod, cam_mod, lsq_mod


# ---- objectives in the reference calling convention: (params[k,n], mask[(B..)]) -> err[k] ----------

class Distort10Objective:
    """compute_distorted_camera_model on [x | fixed pose], squared residual cost."""

    def __init__(self, points_3d, obs, pose):
        _, _, self.cam, self.lsq = _import_reference()
        self.points_3d, self.obs, self.pose = points_3d, obs, pose

    def __call__(self, params, mask):
        m = mask.reshape(-1)
        full = torch.cat([params, self.pose[m]], dim=-1)
        u, v = self.cam.compute_distorted_camera_model(self.points_3d[m], full)
        est = torch.stack([u, v], dim=-1).unsqueeze(1)  # B x F(=1) x N x 2
        res = self.lsq.find_residuals(est, self.obs[m].unsqueeze(1))
        return self.lsq.find_error(res)


class JointObjective:
    """Config 3: the same forward model evaluated per view on [B*V,N,3] with the intrinsics broadcast."""

    def __init__(self, points_3d, obs, views):
        _, _, self.cam, self.lsq = _import_reference()
        self.points_3d, self.obs, self.V = points_3d, obs, views

    def __call__(self, params, mask):
        m = mask.reshape(-1)
        pts = self.points_3d[m]
        k, N = pts.shape[0], pts.shape[1]
        intr = params[:, None, :10].expand(k, self.V, 10)
        full = torch.cat([intr, params[:, 10:].reshape(k, self.V, 6)], dim=-1).reshape(k * self.V, 16)
        pv = pts[:, None].expand(k, self.V, N, 3).reshape(k * self.V, N, 3)
        u, v = self.cam.compute_distorted_camera_model(pv, full)
        est = torch.stack([u, v], dim=-1).reshape(k, self.V, N, 2)
        return self.lsq.find_error(self.lsq.find_residuals(est, self.obs[m]))


class AngleBAObjective:
    """The entry script's objective, networks/calibration_network.py:58-67, composed from the reference's own
    unpack_calibration_parameters (camera_model/calibration_pinhole_camera_model.py:33-75),
    pixel_coordinates_to_homogeneous (geometry/homogeneous_projection.py:21-44), rotate_vector_axis_angle
    (geometry/axis_angle_rotation.py:25-54) and projective_plane_angle_distance
    (geometry/projective_plane_angle_distance.py:20-64).

    get_camera_relative_points (calibration_pinhole_camera_model.py:78-117) crashes on batched input at HEAD
    (the scale is reduced to shape (B,) and divided into (B,M-1,1,3)); `relative_points` below is
    that function with keepdim=True in the three means and nothing else changed, and
    `check_against_unbatched_reference` verifies it against the UNMODIFIED function problem by problem."""

    def __init__(self, obs, visibility, views, points):
        if not available():
            raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
        _import_reference()
        from deep_attention_visual_odometry import camera_model as cm, geometry as geo
        self.cm, self.geo = cm, geo
        self.obs, self.vis, self.M, self.N = obs, visibility, int(views), int(points)

    def relative_points(self, world_points, camera_translations, camera_rotations):
        num_points = world_points.size(-2)
        num_views = camera_translations.size(-3) + 1
        points_scale = world_points.abs().mean(dim=(-1, -2, -3), keepdim=True)
        camera_scale = camera_translations.abs().mean(dim=(-1, -2, -3), keepdim=True)
        overall_scale = (points_scale * num_points + camera_scale * num_views) / (num_points + num_views)
        camera_translations = camera_translations / overall_scale
        world_points = world_points / overall_scale
        rel = self.geo.rotate_vector_axis_angle(world_points, camera_rotations) + camera_translations
        return torch.concatenate([world_points, rel], dim=-3)

    def error(self, params, obs, vis, relative_points=None):
        cp = self.cm.unpack_calibration_parameters(params, self.M, self.N)
        homogeneous = self.geo.pixel_coordinates_to_homogeneous(obs, cp.intrinsics)
        rel = (relative_points or self.relative_points)(world_points=cp.world_points,
                                                        camera_translations=cp.camera_translations,
                                                        camera_rotations=cp.camera_rotations)
        distance = self.geo.projective_plane_angle_distance(homogeneous, rel)
        return (distance * vis).sum(dim=(-1, -2))

    def __call__(self, params, mask):
        m = mask.reshape(-1)
        return self.error(params, self.obs[m], self.vis[m])

    def check_against_unbatched_reference(self, params):
        """max |error(batched, keepdim fix) - error(one problem at a time, unmodified reference function)|."""
        with torch.no_grad():
            batched = self.error(params, self.obs, self.vis)
            single = torch.stack([self.error(params[b], self.obs[b], self.vis[b], self.cm.get_camera_relative_points)
                                  for b in range(params.shape[0])])
        return float((batched - single).abs().max())


def _norm(x):
    return torch.linalg.vector_norm(x, dim=-1, keepdim=True)


ANALYTIC = {
    # tests/autograd_solvers/reference_functions.py:20-62, tests/autograd_solvers/test_bfgs_solver.py:33-46
    "sphere": lambda x, _=None: x.square().sum(dim=-1),
    "sphere_offset": lambda x, _=None: x.square().sum(dim=-1) + 10.0,
    "log_sphere": lambda x, _=None: (x.square().sum(dim=-1) + 1.0).log(),
    "rosenbrock": lambda p, _=None: (1.0 - p[..., 0]).square() + 100.0 * (p[..., 1] - p[..., 0].square()).square(),
    "cosine": lambda x, _=None: (1.0 - (x / _norm(x))[..., 0]) + (1.0 - _norm(x)).square().squeeze(-1),
    "x2_sine": lambda x, _=None: torch.linalg.vector_norm(x, dim=-1).square()
    * (torch.linalg.vector_norm(x, dim=-1).sin() + 2.0),
}


def make_objective(batch_or_name, dtype=torch.float64):
    if isinstance(batch_or_name, str):
        return ANALYTIC[batch_or_name]
    b = batch_or_name
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)
    if b.model == "distort10":
        return Distort10Objective(t(b.points_3d), t(b.obs), t(b.pose))
    if b.model == "angle_ba":
        return AngleBAObjective(t(b.obs), t(b.weights), b.views, b.N)
    return JointObjective(t(b.points_3d), t(b.obs), b.views)


# ---- instrumented solve ---------------------------------------------------------------------------

def reference_solve(objective, x0: torch.Tensor, *, error_threshold=1e-4, iterations=1000, minimum_step=1e-8,
                    sufficient_decrease=1e-4, curvature=0.9, threads: int | None = None):
    """BFGSSolver(...).eval()(x0, objective) with per-problem bookkeeping.

    Returns dict(x, cost, converged, iters, fevals, reason) as numpy arrays; `cost` is the objective
    re-evaluated at the returned parameters (what networks/calibration_network.py:71 does)."""
    bfgs_mod, _, _, _ = _import_reference()
    if threads:
        torch.set_num_threads(threads)
    B = x0.shape[0]
    fevals = np.zeros(B, np.int64)
    steps = np.zeros(B, np.int64)
    last_outer = np.full(B, np.nan)
    step_retired = np.zeros(B, bool)
    state = {"in_ls": False, "first": None}

    def counted(params, mask):
        err = objective(params, mask)
        m = mask.reshape(-1).numpy()
        fevals[m] += 1
        if state["in_ls"]:
            if state["first"] is None:  # the first probe's mask is exactly the set taking this step
                state["first"] = m.copy()
                steps[m] += 1
        else:
            last_outer[m] = err.detach().double().numpy()
        return err

    real_ls = bfgs_mod.line_search_wolfe_conditions

    def wrapped_ls(**kw):
        state["in_ls"], state["first"] = True, None
        try:
            alpha = real_ls(**kw)
        finally:
            state["in_ls"] = False
        # bfgs_solver.py:191,203-205: the step-size retirement test, recomputed with the same arithmetic
        nrm = torch.linalg.vector_norm(alpha.unsqueeze(-1) * kw["search_direction"], dim=-1)
        small = ~torch.greater(nrm, minimum_step)
        if state["first"] is not None:
            idx = np.nonzero(state["first"])[0]
            step_retired[idx[small.numpy()]] = True
        return alpha

    solver = bfgs_mod.BFGSSolver(sufficient_decrease=sufficient_decrease, curvature=curvature,
                                 error_threshold=error_threshold, iterations=iterations,
                                 minimum_step=minimum_step).eval()
    bfgs_mod.line_search_wolfe_conditions = wrapped_ls
    try:
        x = solver(x0.clone(), counted)
    finally:
        bfgs_mod.line_search_wolfe_conditions = real_ls
    with torch.no_grad():
        cost = objective(x, torch.ones(B, dtype=torch.bool)).double().numpy()
    thr = float(torch.tensor(error_threshold, dtype=x0.dtype))
    reason = np.full(B, 2, np.int32)  # CAP
    reason[step_retired] = 1
    reason[~step_retired & np.isnan(last_outer)] = 3
    reason[~step_retired & (last_outer <= thr)] = 0
    return dict(x=x.numpy(), cost=cost.astype(x.numpy().dtype), converged=cost <= thr,
                iters=steps.astype(np.int32), fevals=fevals.astype(np.int32), reason=reason)


def reference_line_search(objective, x, d, *, sufficient_decrease=1e-4, curvature=0.9, strong=False):
    """line_search_wolfe_conditions on (x, d) with base error/gradient from autograd; returns alpha, probes."""
    _, wolfe_mod, _, _ = _import_reference()
    B = x.shape[0]
    probes = np.zeros(B, np.int64)

    def counted(params, mask):
        probes[mask.reshape(-1).numpy()] += 1
        return objective(params, mask)

    xg = x.clone().requires_grad_(True)
    f0 = objective(xg, torch.ones(B, dtype=torch.bool))
    g = torch.autograd.grad(f0.sum(), xg)[0]
    alpha = wolfe_mod.line_search_wolfe_conditions(
        parameters=x, search_direction=d, base_error=f0.detach(), base_gradient=g,
        error_function=counted, sufficient_decrease=sufficient_decrease, curvature=curvature, strong=strong)
    return alpha.numpy(), probes.astype(np.int32), f0.detach().numpy(), g.numpy()


def reference_cost_grad(objective, x):
    xg = x.clone().requires_grad_(True)
    f = objective(xg, torch.ones(x.shape[0], dtype=torch.bool))
    g = torch.autograd.grad(f.sum(), xg)[0]
    return f.detach().numpy(), g.numpy()


def reference_project(points_3d, params16):
    _, _, cam, _ = _import_reference()
    u, v = cam.compute_distorted_camera_model(points_3d, params16)
    return u.numpy(), v.numpy()


def reference_project_autograd_jacobian(points_3d, params16):
    """d stack(u', v') / d params by autograd of the reference forward: J[B,2N,16]
    (the layout of compute_distorted_camera_model_and_jacobian, distorted_camera_model.py:364-384)."""
    _, _, cam, _ = _import_reference()
    B, N, _ = points_3d.shape
    J = np.empty((B, 2 * N, 16), dtype=params16.numpy().dtype)
    for b in range(B):
        def f(th):
            u, v = cam.compute_distorted_camera_model(points_3d[b:b + 1], th[None])
            return torch.cat([u[0], v[0]])
        J[b] = torch.autograd.functional.jacobian(f, params16[b]).numpy()
    return J


def reference_bfgs_update(H, s, y):
    bfgs_mod, _, _, _ = _import_reference()
    return bfgs_mod.BFGSSolver.update_inverse_hessian(H, s, y).numpy()


def reference_initial_scale(s, y):
    bfgs_mod, _, _, _ = _import_reference()
    return bfgs_mod.BFGSSolver.scale_initial_inverse_hessian(s, y).squeeze(-1).numpy()
