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

from .objectives import AngleDistanceObjective
from .solvers import BFGSSolver


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

    def forward(self, true_projected_points: torch.Tensor, visibility_mask: torch.Tensor, return_error: bool = False,
                return_info: bool = False):
        """true_projected_points Bx M x N x 2, visibility_mask B x M x N -> refined parameters B x n
        (and the final error per problem with return_error=True, networks/calibration_network.py:70-72)."""
        inputs = true_projected_points.reshape(-1, 2 * self.num_views * self.num_points)
        initial_guess = self.initial_estimator(inputs)
        objective = AngleDistanceObjective(true_projected_points, visibility_mask, dtype=initial_guess.dtype)
        info = self.solver(initial_guess, objective, return_info=True)
        if return_info:
            return info
        if return_error:
            return info.parameters, info.cost
        return info.parameters
