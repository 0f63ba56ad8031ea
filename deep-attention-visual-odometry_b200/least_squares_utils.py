"""find_residuals / find_error / find_error_gradient with the reference's signatures on the GPU.

Reference: deep_attention_visual_odometry/solvers/least_squares_utils.py:4-48.
"""
from __future__ import annotations

import torch

from . import _lib


def find_residuals(estimated_points: torch.Tensor, true_points: torch.Tensor) -> torch.Tensor:
    """least_squares_utils.py:4-13 — a plain subtraction; stays a tensor op on whatever device holds it."""
    return estimated_points - true_points


def _flat(residuals, point_weights, device):
    B = residuals.shape[0]
    dt = residuals.dtype
    res = residuals.detach().to(device).contiguous().reshape(B, -1)
    w = None
    if point_weights is not None:  # B x F x N x 1 in the reference: broadcast over the trailing (u, v) axis
        w = point_weights.detach().to(device=device, dtype=dt).expand(residuals.shape).contiguous().reshape(B, -1)
    return B, dt, res, w


def find_error(residuals: torch.Tensor, point_weights: torch.Tensor | None = None) -> torch.Tensor:
    """residuals B x F x N x 2 (any trailing shape), weights broadcastable -> error[B] = sum w r^2."""
    device = _lib.require_cuda() if residuals.device.type != "cuda" else residuals.device
    B, dt, res, w = _flat(residuals, point_weights, device)
    with torch.cuda.device(device):
        err = torch.empty(B, dtype=dt, device=device)
        st = _lib.lib().davo_least_squares(_lib.dtype_code(dt), B, res.shape[1], 0, _lib.ptr(res), None, _lib.ptr(w),
                                           _lib.ptr(err), None, _lib.stream_ptr())
    _lib.check(st, "davo_least_squares")
    return err.to(residuals.device)


def find_error_gradient(residuals: torch.Tensor, jacobian: torch.Tensor,
                        point_weights: torch.Tensor | None = None) -> torch.Tensor:
    """residuals B x F x N x 2, jacobian B x F x N x 2 x P -> gradient[B,P] = sum 2 w r J."""
    device = _lib.require_cuda() if residuals.device.type != "cuda" else residuals.device
    B, dt, res, w = _flat(residuals, point_weights, device)
    P = jacobian.shape[-1]
    if tuple(jacobian.shape[:-1]) != tuple(residuals.shape):
        raise ValueError(f"jacobian {tuple(jacobian.shape)} does not match residuals {tuple(residuals.shape)}")
    jac = jacobian.detach().to(device=device, dtype=dt).contiguous().reshape(B, -1, P)
    with torch.cuda.device(device):
        grad = torch.empty(B, P, dtype=dt, device=device)
        st = _lib.lib().davo_least_squares(_lib.dtype_code(dt), B, res.shape[1], P, _lib.ptr(res), _lib.ptr(jac),
                                           _lib.ptr(w), None, _lib.ptr(grad), _lib.stream_ptr())
    _lib.check(st, "davo_least_squares")
    return grad.to(residuals.device)
