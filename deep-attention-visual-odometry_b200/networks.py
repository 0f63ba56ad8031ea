"""CalibrationNetwork with the reference's interface (networks/calibration_network.py:26-73): an MLP proposes
the calibration parameters from the projected points and the batched BFGS solve refines them against the
bundle-adjustment objective.

The refinement is the hot path and runs in the sm_100a solve kernel (AngleDistanceObjective + BFGSSolver through
the C-ABI).  The initial-guess MLP is outside this repo's scope table and is kept as the
same stock torch modules the reference builds, with the same parameter names, so a reference checkpoint loads
with ``load_state_dict``.  In training mode (and whenever the guess requires grad, as in the reference:
create_graph = parameters.requires_grad, autograd_solvers/bfgs_solver.py:85) the solve is differentiable: gradients
reach the MLP's weights through BFGSSolver's backward kernel (csrc/solver_train.cuh).
"""
from __future__ import annotations

import torch
import torch.nn as nn

import ctypes

from . import _lib
from .objectives import AngleDistanceObjective
from .solvers import BFGSSolver


class FusedInitialEstimator:
    """The initial-guess MLP (networks/calibration_network.py:35-43: Linear-GELU-BatchNorm1d x 2 + Linear) in
    inference form as one tcgen05 kernel (csrc/mlp_kernels.cu, davo_mlp_forward), fed from the torch modules'
    parameters: weights are re-laid (davo_mlp_pack_weights) and the BatchNorm statistics folded whenever a parameter
    or buffer has changed since the last call."""

    def __init__(self, sequential: nn.Sequential):
        self.lin = [sequential[0], sequential[3], sequential[6]]
        self.bn = [sequential[2], sequential[5]]
        self._key = None
        self._packed = None

    def supported(self, x: torch.Tensor) -> bool:
        l1, l2, l3 = self.lin
        H = l1.out_features
        return (x.is_cuda and x.dtype == torch.float32 and l1.weight.dtype == torch.float32 and l1.weight.is_cuda
                and l1.in_features % 8 == 0 and H % 16 == 0 and H <= 256 and l2.in_features == H
                and l2.out_features == H and l3.in_features == H and l3.out_features <= 256
                and all(l.bias is not None for l in self.lin)
                and all(b.track_running_stats and b.running_mean is not None for b in self.bn))

    def _state_key(self):
        ts = [t for l in self.lin for t in (l.weight, l.bias)]
        ts += [t for b in self.bn for t in (b.weight, b.bias, b.running_mean, b.running_var) if t is not None]
        return tuple((t.data_ptr(), t._version) for t in ts)

    def _prepare(self, device):
        key = self._state_key()
        if key == self._key:
            return self._packed
        lib = _lib.lib()
        packed = []
        with torch.cuda.device(device), torch.no_grad():
            for l in self.lin:
                N, K = l.out_features, l.in_features
                buf = torch.empty(int(lib.davo_mlp_packed_bytes(N, K)), dtype=torch.uint8, device=device)
                w = l.weight.detach().contiguous()
                _lib.check(lib.davo_mlp_pack_weights(N, K, _lib.ptr(w), _lib.ptr(buf), _lib.stream_ptr()),
                           "davo_mlp_pack_weights")
                packed.append(buf)
            fold = []
            for b in self.bn:  # BatchNorm1d in eval mode: (x - mean) / sqrt(var + eps) * weight + bias
                scale = torch.rsqrt(b.running_var.detach() + b.eps)
                if b.weight is not None:
                    scale = scale * b.weight.detach()
                shift = -b.running_mean.detach() * scale
                if b.bias is not None:
                    shift = shift + b.bias.detach()
                fold.append((scale.contiguous(), shift.contiguous()))
            biases = [l.bias.detach().contiguous() for l in self.lin]
        self._packed = (packed, fold, biases)
        self._key = key
        return self._packed

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        l1, _, l3 = self.lin
        device = x.device
        packed, fold, biases = self._prepare(device)
        x = x.detach().contiguous()
        out = torch.empty(x.shape[0], l3.out_features, dtype=torch.float32, device=device)
        desc = _lib.MlpDesc(x.shape[0], l1.in_features, l1.out_features, l3.out_features)
        with torch.cuda.device(device):
            st = _lib.lib().davo_mlp_forward(
                ctypes.byref(desc), _lib.ptr(x), _lib.ptr(packed[0]), _lib.ptr(biases[0]), _lib.ptr(fold[0][0]),
                _lib.ptr(fold[0][1]), _lib.ptr(packed[1]), _lib.ptr(biases[1]), _lib.ptr(fold[1][0]),
                _lib.ptr(fold[1][1]), _lib.ptr(packed[2]), _lib.ptr(biases[2]), _lib.ptr(out), _lib.stream_ptr())
        _lib.check(st, "davo_mlp_forward")
        return out


def unpack_calibration_parameters(parameters: torch.Tensor, num_views: int, num_points: int):
    """camera_model/calibration_pinhole_camera_model.py:33-75: views of the parameter vector
    (intrinsics (B..)x1x1x3, world points (B..)x1xNx3, translations and rotations (B..)x(M-1)x1x3)."""
    expected = 3 + 3 * num_points + 6 * (num_views - 1)
    if parameters.size(-1) != expected:
        raise ValueError(f"The final dimension of the input tensor must be 3 + 3 * num_points + 6 * (num_views - 1) "
                         f"= {expected}, got {parameters.size(-1)}")
    lead = parameters.shape[:-1]
    pe = 3 + 3 * num_points
    te = pe + 3 * (num_views - 1)
    return (parameters[..., 0:3].reshape(lead + (1, 1, 3)),
            parameters[..., 3:pe].reshape(lead + (1, num_points, 3)),
            parameters[..., pe:te].reshape(lead + (num_views - 1, 1, 3)),
            parameters[..., te:].reshape(lead + (num_views - 1, 1, 3)))


class CalibrationNetwork(nn.Module):
    def __init__(self, num_views: int, num_points: int, hidden_size: int = -1):
        super().__init__()
        self._num_views = int(num_views)
        self._num_points = int(num_points)
        num_inputs = num_views * num_points * 2
        num_parameters = 3 + 3 * num_points + 6 * (num_views - 1)
        if hidden_size <= 0:
            hidden_size = 4 * num_inputs
        self.initial_estimator = nn.Sequential(  # networks/calibration_network.py:35-43
            nn.Linear(num_inputs, hidden_size), nn.GELU(), nn.BatchNorm1d(hidden_size),
            nn.Linear(hidden_size, hidden_size), nn.GELU(), nn.BatchNorm1d(hidden_size),
            nn.Linear(hidden_size, num_parameters))
        self.solver = BFGSSolver(error_threshold=1e-7, training_error_threshold=1e-3)  # :44

    @property
    def num_views(self) -> int:
        return self._num_views

    @property
    def num_points(self) -> int:
        return self._num_points

    def estimate(self, inputs: torch.Tensor) -> torch.Tensor:
        """The initial guess (networks/calibration_network.py:55-56): the fused kernel in inference, else the modules."""
        fused = self.__dict__.get("_fused")
        if fused is None:
            fused = self.__dict__["_fused"] = FusedInitialEstimator(self.initial_estimator)
        if not self.training and not torch.is_grad_enabled() and fused.supported(inputs):
            return fused(inputs)
        return self.initial_estimator(inputs)

    def forward(self, true_projected_points: torch.Tensor, visibility_mask: torch.Tensor, return_error: bool = False,
                return_info: bool = False):
        """true_projected_points Bx M x N x 2, visibility_mask B x M x N -> refined parameters B x n
        (and the final error per problem with return_error=True, networks/calibration_network.py:70-72)."""
        inputs = true_projected_points.reshape(-1, 2 * self.num_views * self.num_points)
        initial_guess = self.estimate(inputs)
        objective = AngleDistanceObjective(true_projected_points, visibility_mask, dtype=initial_guess.dtype)
        info = self.solver(initial_guess, objective, return_info=True)
        if return_info:
            return info
        if return_error:
            return info.parameters, info.cost
        return info.parameters
