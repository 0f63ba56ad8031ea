"""Objective descriptors: what the reference passes to BFGSSolver as a Python closure.

The reference solver takes ``error_function(params[k,n], mask[(B..)]) -> err[k]`` and differentiates
it with autograd (autograd_solvers/bfgs_solver.py:131-135).  A Python closure cannot run inside a
CUDA kernel, so the B200 solver takes a *descriptor*: an object that owns the device buffers of the
problem set and names the objective the kernels have compiled in.  A descriptor is also callable with
the reference's convention (cost and its autograd gradient both come from the CUDA evaluator
``davo_eval_cost_grad``), so the reference's own solver can drive it on the GPU.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _to_device(t, dtype, device, non_blocking=True):
    if t is None:
        return None
    t = torch.as_tensor(t)
    return t.to(device=device, dtype=dtype, non_blocking=non_blocking).contiguous()


class _EvalFunction(torch.autograd.Function):
    """cost = objective(params) with d cost / d params from the analytic J^T r kernel."""

    @staticmethod
    def forward(ctx, params, objective, index):
        cost, grad = objective.evaluate(params, index=index, want_grad=True)
        ctx.save_for_backward(grad.to(device=params.device, dtype=params.dtype))
        return cost.to(device=params.device, dtype=params.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad_out.unsqueeze(-1) * grad, None, None


class CalibrationObjective:
    """Base class: B independent problems of one compiled-in model."""

    model: str = ""
    views: int = 1
    #: a tensor of the problem data that requires grad (the solve then back-propagates into it), or None
    differentiable_data = None

    def __init__(self, batch_shape, n, N, dtype, device):
        self.batch_shape = tuple(batch_shape)
        self.B = 1
        for s in self.batch_shape:
            self.B *= int(s)
        self.n, self.N, self.dtype, self.device = int(n), int(N), dtype, device
        self.data0 = self.data1 = self.weights = None

    # ---- what the solver needs -------------------------------------------------------------------
    def has_weights(self) -> bool:
        return self.weights is not None

    def desc(self, B=None, **kw) -> _lib.ProblemDesc:
        return _lib.make_desc(self.B if B is None else B, self.N, self.views, self.n, self.model, self.dtype,
                              has_weights=self.has_weights(), **kw)

    def _select(self, index):
        """Rows of the problem set for a subset of problems (the reference's `mask`)."""
        if index is None:
            return self.data0, self.data1, self.weights
        pick = lambda t: None if t is None else t[index].contiguous()  # buffers are stored flat: [B, ...]
        return pick(self.data0), pick(self.data1), pick(self.weights)

    def evaluate(self, params: torch.Tensor, index=None, want_grad: bool = True):
        """cost[k] (and grad[k,n]) for params[k,n] on the problems `index` (all when None)."""
        _lib.require_cuda()
        p = params.detach().to(device=self.device, dtype=self.dtype).reshape(-1, self.n).contiguous()
        d0, d1, w = self._select(index)
        k = p.shape[0]
        cost = torch.empty(k, dtype=self.dtype, device=self.device)
        grad = torch.empty(k, self.n, dtype=self.dtype, device=self.device) if want_grad else None
        desc = self.desc(B=k)
        st = _lib.lib().davo_eval_cost_grad(ctypes.byref(desc), _lib.ptr(d0), _lib.ptr(d1), _lib.ptr(w), _lib.ptr(p),
                                            _lib.ptr(cost), _lib.ptr(grad), _lib.stream_ptr())
        _lib.check(st, "davo_eval_cost_grad")
        return cost, grad

    def as_float64(self) -> "CalibrationObjective":
        """This problem set with float64 buffers (the backward pass of the differentiable solve runs in float64);
        `self` when it already is, otherwise a cached twin that shares nothing mutable."""
        if self.dtype == torch.float64:
            return self
        twin = getattr(self, "_twin64", None)
        if twin is None:
            import copy
            d0, d1, w = self.data0, self.data1, self.weights  # stages a lazily staged problem set first
            twin = copy.copy(self)
            twin.dtype = torch.float64
            if hasattr(twin, "_raw"):
                twin._raw = None
            up = lambda t: None if t is None else t.double()
            twin.data0, twin.data1, twin.weights = up(d0), up(d1), up(w)
            self._twin64 = twin
        return twin

    # ---- the reference's calling convention ------------------------------------------------------
    def __call__(self, params: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        index = None
        if mask is not None and mask.numel() == self.B and mask.dtype == torch.bool:
            index = mask.reshape(-1).to(self.device).nonzero(as_tuple=True)[0]
        lead = params.shape[:-1]
        out = _EvalFunction.apply(params.reshape(-1, self.n), self, index)
        return out.reshape(lead)


class DistortionObjective(CalibrationObjective):
    """Fit (cx,cy,k1,k2,k3,p1,p2,fx,s,fy) of the 16-parameter camera model with the pose fixed.

    error = sum_matches w * |compute_distorted_camera_model(points_3d, [x | pose]) - observed|^2
    (camera_model/distorted_camera_model.py:24-111, solvers/least_squares_utils.py:4-28).  The
    constructor runs ``davo_stage_matches`` once: {x'/z', y'/z', u*, v*} per match, 16 B in float32.

    points_3d [(B..),N,3], observed_2d [(B..),N,2], pose [(B..),6]=(rx,ry,rz,tx,ty,tz) or None
    (identity), weights [(B..),N] or None.  Host tensors are copied to the current CUDA device.
    """

    model = "distort10"

    def __init__(self, points_3d, observed_2d, pose=None, weights=None, dtype=None, device=None):
        device = _lib.require_cuda() if device is None else torch.device(device)
        points_3d = torch.as_tensor(points_3d)
        observed_2d = torch.as_tensor(observed_2d)
        dtype = dtype or points_3d.dtype
        batch_shape, N = points_3d.shape[:-2], points_3d.shape[-2]
        super().__init__(batch_shape, 10, N, dtype, device)
        if tuple(observed_2d.shape) != tuple(batch_shape) + (N, 2):
            raise ValueError(f"observed_2d must be {tuple(batch_shape) + (N, 2)}, got {tuple(observed_2d.shape)}")
        if points_3d.shape[-1] != 3:
            raise ValueError("points_3d must end in a dimension of 3")
        flat = lambda t, tail: None if t is None else torch.as_tensor(t).reshape((self.B,) + tail)
        # the differentiable solve also returns d loss / d observed_2d (davo_solve_backward, grad_data)
        self.differentiable_data = observed_2d if observed_2d.requires_grad else None
        detach = lambda t: t.detach() if isinstance(t, torch.Tensor) else t
        points_3d, observed_2d, pose, weights = detach(points_3d), detach(observed_2d), detach(pose), detach(weights)
        self._raw = (flat(points_3d, (N, 3)), flat(observed_2d, (N, 2)), flat(pose, (6,)), flat(weights, (N,)))
        self._data0 = None
        self.weights = None
        # Host inputs are staged lazily so that BFGSSolver.forward can overlap the host-to-device copies of
        # one chunk of problems with the staging and solve of the previous chunk; device inputs stage now.
        if all(t is None or t.device.type == "cuda" for t in self._raw):
            self.materialize()

    @property
    def data0(self):
        if self._data0 is None and getattr(self, "_raw", None) is not None:
            self.materialize()
        return self._data0

    @data0.setter
    def data0(self, value):
        self._data0 = value

    @property
    def is_staged(self) -> bool:
        return self._data0 is not None

    def has_weights(self) -> bool:
        # a lazily staged (host-input) problem set knows it is weighted before its weights are on the device
        raw = getattr(self, "_raw", None)
        return self.weights is not None or (raw is not None and raw[3] is not None)

    def upload_poses(self, d_pose) -> None:
        """Host -> device copy of every problem's pose (24 to 48 bytes per problem: one copy for the whole batch, a
        copy per chunk would be eighteen small transfers between the large ones)."""
        if self._raw[2] is not None:
            d_pose.copy_(self._raw[2], non_blocking=True)

    def upload_rows(self, lo: int, hi: int, d_pts: torch.Tensor, d_obs: torch.Tensor, d_pose, weights_out=None,
                    poses: bool = True) -> None:
        """Host -> device copy of the raw inputs of problems [lo, hi) into rows lo:hi of preallocated device
        buffers, on the current stream (the streamed solve runs this on its copy stream)."""
        pts, obs, pose, w = self._raw
        d_pts[lo:hi].copy_(pts[lo:hi], non_blocking=True)
        d_obs[lo:hi].copy_(obs[lo:hi], non_blocking=True)
        if pose is not None and poses:
            d_pose[lo:hi].copy_(pose[lo:hi], non_blocking=True)
        if w is not None and weights_out is not None:
            weights_out[lo:hi].copy_(w[lo:hi], non_blocking=True)

    def stage_device_rows(self, lo: int, hi: int, d_pts, d_obs, d_pose, staged: torch.Tensor) -> None:
        """davo_stage_matches on device-resident rows lo:hi, writing staged[lo:hi]; current stream."""
        desc = self.desc(B=hi - lo)
        st = _lib.lib().davo_stage_matches(ctypes.byref(desc), _lib.ptr(d_pts[lo:hi]), _lib.ptr(d_obs[lo:hi]),
                                           _lib.ptr(None if d_pose is None else d_pose[lo:hi]),
                                           _lib.ptr(staged[lo:hi]), _lib.stream_ptr())
        _lib.check(st, "davo_stage_matches")

    def stage_rows(self, lo: int, hi: int, staged: torch.Tensor, weights_out=None) -> None:
        """Copy problems [lo, hi) to the device (if they are on the host) and run davo_stage_matches on them,
        writing staged[lo:hi].  Stream ordered on the current stream; no host synchronisation."""
        pts, obs, pose, w = self._raw
        up = lambda t: None if t is None else t[lo:hi].to(device=self.device, dtype=self.dtype, non_blocking=True).contiguous()
        d_pts, d_obs, d_pose = up(pts), up(obs), up(pose)
        if w is not None and weights_out is not None:
            weights_out[lo:hi].copy_(w[lo:hi], non_blocking=True)
        desc = self.desc(B=hi - lo)
        st = _lib.lib().davo_stage_matches(ctypes.byref(desc), _lib.ptr(d_pts), _lib.ptr(d_obs), _lib.ptr(d_pose),
                                           _lib.ptr(staged[lo:hi]), _lib.stream_ptr())
        _lib.check(st, "davo_stage_matches")

    def materialize(self) -> None:
        """Stage every problem now (one davo_stage_matches launch)."""
        if self._data0 is not None:
            return
        with torch.cuda.device(self.device):
            staged = torch.empty(self.B, self.N, 4, dtype=self.dtype, device=self.device)
            w = self._raw[3]
            wdev = None if w is None else torch.empty(self.B, self.N, dtype=self.dtype, device=self.device)
            self.stage_rows(0, self.B, staged, wdev)
            self._data0, self.weights = staged, wdev

    @classmethod
    def from_staged(cls, staged: torch.Tensor, weights=None):
        """Wrap an already staged [(B..),N,4] device buffer (no copy)."""
        self = cls.__new__(cls)
        self._raw = None
        self._data0 = None
        CalibrationObjective.__init__(self, staged.shape[:-2], 10, staged.shape[-2], staged.dtype, staged.device)
        self.data0 = staged.contiguous().reshape(self.B, self.N, 4)
        self.weights = None if weights is None else weights.to(staged.device, staged.dtype).reshape(self.B, self.N)
        return self


class JointPoseObjective(CalibrationObjective):
    """Fit the 10 intrinsics plus (rx,ry,rz,tx,ty,tz) of each of V views: n = 10 + 6V.

    points_3d [(B..),N,3] are world points shared by the views; observed_2d [(B..),V,N,2].
    """

    model = "joint"

    def __init__(self, points_3d, observed_2d, weights=None, dtype=None, device=None):
        device = _lib.require_cuda() if device is None else torch.device(device)
        points_3d = torch.as_tensor(points_3d)
        observed_2d = torch.as_tensor(observed_2d)
        dtype = dtype or points_3d.dtype
        batch_shape, N = points_3d.shape[:-2], points_3d.shape[-2]
        V = observed_2d.shape[-3]
        if tuple(observed_2d.shape) != tuple(batch_shape) + (V, N, 2):
            raise ValueError(f"observed_2d must be {tuple(batch_shape) + (V, N, 2)}, got {tuple(observed_2d.shape)}")
        super().__init__(batch_shape, 10 + 6 * V, N, dtype, device)
        self.views = V
        # the differentiable solve also returns d loss / d observed_2d (davo_solve_backward, grad_data)
        self.differentiable_data = observed_2d if observed_2d.requires_grad else None
        points_3d, observed_2d = points_3d.detach(), observed_2d.detach()
        self.data0 = _to_device(points_3d, dtype, device).reshape(self.B, N, 3)
        self.data1 = _to_device(observed_2d, dtype, device).reshape(self.B, V, N, 2)
        self.weights = None if weights is None else _to_device(weights, dtype, device).reshape(self.B, V, N)


class AngleDistanceObjective(CalibrationObjective):
    """The entry script's objective (networks/calibration_network.py:58-67): bundle adjustment of
    (f, cx, cy | N world points | M-1 translations | M-1 axis-angle rotations), n = 3 + 3N + 6(M-1), with the
    visibility-weighted angle between each pixel's ray and its camera-relative point as the error.

    true_projected_points [(B..),M,N,2], visibility_mask [(B..),M,N] (bool or float; None = all visible):
    the two tensors CalibrationNetwork.forward receives.
    """

    model = "angle_ba"

    def __init__(self, true_projected_points, visibility_mask=None, dtype=None, device=None):
        device = _lib.require_cuda() if device is None else torch.device(device)
        obs = torch.as_tensor(true_projected_points)
        if obs.dim() < 3 or obs.shape[-1] != 2:
            raise ValueError("true_projected_points must be (B..) x M x N x 2")
        dtype = dtype or obs.dtype
        batch_shape, M, N = obs.shape[:-3], obs.shape[-3], obs.shape[-2]
        if M < 2:
            raise ValueError("the bundle-adjustment objective needs at least two views")
        super().__init__(batch_shape, 3 + 3 * N + 6 * (M - 1), N, dtype, device)
        self.views = M
        # the differentiable solve also returns d loss / d true_projected_points (davo_solve_backward, grad_data)
        self.differentiable_data = obs if obs.requires_grad else None
        obs = obs.detach()
        self.data0 = _to_device(obs, dtype, device).reshape(self.B, M, N, 2)
        if visibility_mask is not None:
            vis = torch.as_tensor(visibility_mask)
            if tuple(vis.shape) != tuple(batch_shape) + (M, N):
                raise ValueError(f"visibility_mask must be {tuple(batch_shape) + (M, N)}, got {tuple(vis.shape)}")
            self.weights = _to_device(vis, dtype, device).reshape(self.B, M, N)


class AnalyticObjective(CalibrationObjective):
    """The reference's analytic test objectives (tests/autograd_solvers/reference_functions.py:20-62,
    test_bfgs_solver.py:33-46, line_search/test_wolffe_conditions.py:214-305), compiled into the solver
    so that the reference's solver tests can be mirrored on the GPU path."""

    def __init__(self, name: str, batch_shape, n: int, dtype=torch.float32, target=None, device=None):
        device = _lib.require_cuda() if device is None else torch.device(device)
        if name not in _lib.MODEL_IDS or _lib.MODEL_IDS[name] < 16:
            raise ValueError(f"unknown analytic objective {name!r}")
        super().__init__(batch_shape, n, 0, dtype, device)
        self.model = name
        if name == "distance":
            if target is None:
                raise ValueError("the 'distance' objective needs a target")
            self.data0 = _to_device(target, dtype, device).expand(self.batch_shape + (n,)).contiguous().reshape(self.B, n)
