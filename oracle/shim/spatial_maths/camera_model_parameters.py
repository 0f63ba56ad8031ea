"""Index layout of the 16-parameter camera model (see package docstring for the evidence)."""
import torch

CX = 0
CY = 1
K1 = 2
K2 = 3
K3 = 4
P1 = 5
P2 = 6
FX = 7
S = 8
FY = 9
RX = 10
RY = 11
RZ = 12
TX = 13
TY = 14
TZ = 15


def make_camera_parameters(cx, cy, k1, k2, k3, p1, p2, fx, s, fy, tx, ty, tz, rx, ry, rz):
    """Keyword order follows tests/camera_model/test_distorted_camera_model.py:144-161 (t* before r*);
    the returned vector is in INDEX order (r* before t*)."""
    return torch.stack(
        [torch.as_tensor(v) for v in (cx, cy, k1, k2, k3, p1, p2, fx, s, fy, rx, ry, rz, tx, ty, tz)], dim=-1
    )
