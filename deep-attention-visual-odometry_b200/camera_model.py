"""compute_distorted_camera_model[_and_jacobian] with the reference's signatures on the GPU.

Reference: deep_attention_visual_odometry/camera_model/distorted_camera_model.py:106-111, 114-385.
The Jacobian is d(u', v')/d theta of the forward model for all 16 parameters; the reference's
hand-written columns fx, s, fy, tx, ty (when p1, p2 != 0) and rx, ry, rz disagree with autograd of its
own forward model and are NOT reproduced.
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import torch

from . import _lib


def _prepare(points_3d: torch.Tensor, parameters: torch.Tensor):
    device = _lib.require_cuda() if points_3d.device.type != "cuda" else points_3d.device
    if points_3d.dim() != 3 or points_3d.shape[-1] != 3:
        raise ValueError(f"points_3d must be B x N x 3, got {tuple(points_3d.shape)}")
    if parameters.dim() != 2 or parameters.shape[-1] != 16 or parameters.shape[0] != points_3d.shape[0]:
        raise ValueError(f"parameters must be B x 16, got {tuple(parameters.shape)}")
    dt = parameters.dtype
    pts = points_3d.detach().to(device=device, dtype=dt).contiguous()
    th = parameters.detach().to(device=device).contiguous()
    return device, dt, pts, th


class _ProjectWithGrad(torch.autograd.Function):
    """(u', v') = forward model, differentiable with respect to the 16 parameters: the backward pass contracts the
    incoming gradients with the Jacobian the forward pass computed (davo_project_jacobian)."""

    @staticmethod
    def forward(ctx, points_3d, parameters):
        J, u, v = compute_distorted_camera_model_and_jacobian(points_3d, parameters)
        ctx.save_for_backward(J)
        return u, v

    @staticmethod
    def backward(ctx, grad_u, grad_v):
        (J,) = ctx.saved_tensors
        N = J.shape[1] // 2
        zero = lambda g: torch.zeros(J.shape[0], N, dtype=J.dtype, device=J.device) if g is None else g.to(J.dtype)
        g = torch.cat([zero(grad_u), zero(grad_v)], dim=1)           # [B, 2N]
        return None, torch.bmm(g.unsqueeze(1), J).squeeze(1)         # d/d points_3d is not provided


def compute_distorted_camera_model(points_3d: torch.Tensor, parameters: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """points_3d [B,N,3], parameters [B,16] -> (u'[B,N], v'[B,N]).  Differentiable with respect to `parameters`
    (the reference's TorchScript function is differentiated by autograd; here the analytic Jacobian is used);
    `points_3d` is treated as data."""
    if points_3d.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("compute_distorted_camera_model is differentiable with respect to `parameters` "
                                  "only; points_3d is treated as data (detach it)")
    if parameters.requires_grad and torch.is_grad_enabled():
        return _ProjectWithGrad.apply(points_3d, parameters)
    device, dt, pts, th = _prepare(points_3d, parameters)
    B, N = pts.shape[0], pts.shape[1]
    with torch.cuda.device(device):
        u = torch.empty(B, N, dtype=dt, device=device)
        v = torch.empty(B, N, dtype=dt, device=device)
        desc = _lib.make_desc(B, N, 1, 16, "distort10", dt)
        st = _lib.lib().davo_project(ctypes.byref(desc), _lib.ptr(pts), _lib.ptr(th), _lib.ptr(u), _lib.ptr(v),
                                     _lib.stream_ptr())
    _lib.check(st, "davo_project")
    return u.to(parameters.device), v.to(parameters.device)


def compute_distorted_camera_model_and_jacobian(points_3d: torch.Tensor, parameters: torch.Tensor
                                                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (J[B,2N,16], u'[B,N], v'[B,N]); rows 0..N-1 of J are du'/dtheta, rows N..2N-1 dv'/dtheta."""
    device, dt, pts, th = _prepare(points_3d, parameters)
    B, N = pts.shape[0], pts.shape[1]
    with torch.cuda.device(device):
        u = torch.empty(B, N, dtype=dt, device=device)
        v = torch.empty(B, N, dtype=dt, device=device)
        J = torch.empty(B, 2 * N, 16, dtype=dt, device=device)
        desc = _lib.make_desc(B, N, 1, 16, "distort10", dt)
        st = _lib.lib().davo_project_jacobian(ctypes.byref(desc), _lib.ptr(pts), _lib.ptr(th), _lib.ptr(J),
                                              _lib.ptr(u), _lib.ptr(v), _lib.stream_ptr())
    _lib.check(st, "davo_project_jacobian")
    out = parameters.device
    return J.to(out), u.to(out), v.to(out)
