"""Time the host-input (streamed) path for several chunk / tail sizes.  GPU box: python tools/e2e_probe.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import SOLVER_KW, make_batch
B = 65536
batch = make_batch("cfg2", B, 0xB200)
h_pts, h_obs, h_x0 = (torch.from_numpy(a).pin_memory() for a in (batch.points_3d, batch.obs, batch.x0))
solver = davo_b200.BFGSSolver(**SOLVER_KW["cfg2"]).eval()
for chunk, tail in ((4096, 1024), (8192, 1024), (16384, 1024), (8192, 2048), (4096, 1024)):
    solver.stream_chunk = chunk
    davo_b200.BFGSSolver.stream_tail = tail
    ts = []
    for rep in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        obj = davo_b200.DistortionObjective(h_pts, h_obs)
        info = solver(h_x0, obj, return_info=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts[2:]))
    print(f"chunk {chunk} tail {tail}: {dt*1e3:.2f} ms -> {B/dt/1e6:.2f} M solves/s")
