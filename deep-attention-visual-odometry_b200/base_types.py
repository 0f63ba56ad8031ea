"""Batch types of the reference's data loaders (base_types/camera_views_and_points.py:21-33)."""
from __future__ import annotations

from typing import NamedTuple

from torch import Tensor


class CameraViewsAndPoints(NamedTuple):
    """A set of world points, camera parameters, and those points as viewed by those cameras: M views of N
    points, with a leading batch dimension.  Same field names and shapes as the reference's NamedTuple."""

    projected_points: Tensor      # (Bx)MxNx2
    visibility_mask: Tensor       # (Bx)MxN
    camera_intrinsics: Tensor     # (Bx)3
    camera_orientations: Tensor   # (Bx)(M - 1)x3
    camera_translations: Tensor   # (Bx)(M - 1)x3
    world_points: Tensor          # (Bx)Nx3
