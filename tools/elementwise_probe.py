"""Achieved HBM bandwidth of the stand-alone elementwise entry points (GPU box)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e-3
B, N = 32768, 256
pts = torch.randn(B, N, 3, device="cuda").abs() + 1.0
th = torch.zeros(B, 16, device="cuda"); th[:, 7] = 1.2; th[:, 9] = 1.1; th[:, 10:13] = 0.1 * torch.randn(B, 3, device="cuda")
t = timed(lambda: davo_b200.compute_distorted_camera_model(pts, th))
print(f"project      : {t*1e6:.0f} us -> {B*N*(12+8)/t/1e9:.0f} GB/s (12 B read + 8 B written per match)")
Bj = 8192
t = timed(lambda: davo_b200.compute_distorted_camera_model_and_jacobian(pts[:Bj], th[:Bj]))
print(f"project + J  : {t*1e6:.0f} us -> {Bj*N*(12+8+128)/t/1e9:.0f} GB/s (12 B read + 136 B written per match)")
res = torch.randn(B, 1, N, 2, device="cuda")
t = timed(lambda: davo_b200.find_error(res))
print(f"find_error   : {t*1e6:.0f} us -> {B*N*8/t/1e9:.0f} GB/s")
jac = torch.randn(2048, 1, N, 2, 10, device="cuda")
t = timed(lambda: davo_b200.find_error_gradient(res[:2048], jac))
print(f"find_gradient: {t*1e6:.0f} us -> {2048*N*2*44/t/1e9:.0f} GB/s")
H = torch.randn(B * 4, 10, 10, device="cuda"); s = torch.randn(B * 4, 10, device="cuda"); y = torch.randn(B * 4, 10, device="cuda")
t = timed(lambda: davo_b200.BFGSSolver.update_inverse_hessian(H, s, y))
print(f"bfgs update  : {t*1e6:.0f} us -> {B*4*(200*4+80)/t/1e9:.0f} GB/s (incl. the wrapper's clone)")
