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


def _sum_to(t: torch.Tensor, shape) -> torch.Tensor:
    """Reduce a gradient of the broadcast shape back to `shape` (what autograd does for an expanded operand)."""
    shape = tuple(shape)
    while t.dim() > len(shape):
        t = t.sum(0)
    dims = [i for i, (a, b) in enumerate(zip(t.shape, shape)) if b == 1 and a != 1]
    return t.sum(dims, keepdim=True) if dims else t


class _ErrorWithGrad(torch.autograd.Function):
    """find_error as a differentiable op (the reference's is plain torch): the forward value comes from the kernel,
    the backward pass is d/dr = 2 w r, d/dw = r^2."""

    @staticmethod
    def forward(ctx, residuals, point_weights):
        ctx.save_for_backward(residuals, point_weights)
        return _find_error(residuals, point_weights)

    @staticmethod
    def backward(ctx, grad_out):
        r, w = ctx.saved_tensors
        g = grad_out.reshape((-1,) + (1,) * (r.dim() - 1))
        gr = gw = None
        if ctx.needs_input_grad[0]:
            gr = 2.0 * r * g if w is None else 2.0 * w * r * g
        if w is not None and ctx.needs_input_grad[1]:
            gw = _sum_to(r.square() * g, w.shape)
        return gr, gw


class _ErrorGradientWithGrad(torch.autograd.Function):
    """find_error_gradient as a differentiable op: out[b,p] = sum 2 w r J."""

    @staticmethod
    def forward(ctx, residuals, jacobian, point_weights):
        ctx.save_for_backward(residuals, jacobian, point_weights)
        return _find_error_gradient(residuals, jacobian, point_weights)

    @staticmethod
    def backward(ctx, grad_out):
        r, J, w = ctx.saved_tensors
        g = grad_out.reshape((r.shape[0],) + (1,) * (r.dim() - 1) + (J.shape[-1],))
        Jg = (J * g).sum(-1)                                   # [B, ...] directional derivative of each residual
        wr = r if w is None else w * r
        gr = gJ = gw = None
        if ctx.needs_input_grad[0]:
            gr = 2.0 * Jg if w is None else 2.0 * w * Jg
        if ctx.needs_input_grad[1]:
            gJ = 2.0 * wr.unsqueeze(-1) * g
        if w is not None and ctx.needs_input_grad[2]:
            gw = _sum_to(2.0 * r * Jg, w.shape)
        return gr, gJ, gw


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def find_error(residuals: torch.Tensor, point_weights: torch.Tensor | None = None) -> torch.Tensor:
    """residuals B x F x N x 2 (any trailing shape), weights broadcastable -> error[B] = sum w r^2.
    Differentiable with respect to the residuals and the weights, like the reference's torch ops."""
    if _needs_grad(residuals, point_weights):
        return _ErrorWithGrad.apply(residuals, point_weights)
    return _find_error(residuals, point_weights)


def _find_error(residuals: torch.Tensor, point_weights: torch.Tensor | None = None) -> torch.Tensor:
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
    """residuals B x F x N x 2, jacobian B x F x N x 2 x P -> gradient[B,P] = sum 2 w r J.
    Differentiable with respect to residuals, jacobian and weights, like the reference's torch ops."""
    if _needs_grad(residuals, jacobian, point_weights):
        return _ErrorGradientWithGrad.apply(residuals, jacobian, point_weights)
    return _find_error_gradient(residuals, jacobian, point_weights)


def _find_error_gradient(residuals: torch.Tensor, jacobian: torch.Tensor,
                         point_weights: torch.Tensor | None = None) -> torch.Tensor:
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
