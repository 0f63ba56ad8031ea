"""Time the host-input (streamed) path for several chunk sizes.  GPU box: python tools/e2e_probe.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import SOLVER_KW, make_batch
B = 65536
batch = make_batch("cfg2", B, 0xB200)
h_pts, h_obs, h_x0 = (torch.from_numpy(a).pin_memory() for a in (batch.points_3d, batch.obs, batch.x0))
solver = davo_b200.BFGSSolver(**SOLVER_KW["cfg2"]).eval()
d = torch.empty_like(h_pts, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(h_pts, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"H2D pinned {h_pts.numel()*4/1e6:.0f} MB: {dt*1e3:.2f} ms = {h_pts.numel()*4/dt/1e9:.1f} GB/s")
for chunk in (65536, 32768, 16384, 8192, 4096):
    solver.stream_chunk = chunk
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        obj = davo_b200.DistortionObjective(h_pts, h_obs)
        info = solver(h_x0, obj, return_info=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"chunk {chunk}: {dt*1e3:.2f} ms -> {B/dt/1e6:.2f} M solves/s")
