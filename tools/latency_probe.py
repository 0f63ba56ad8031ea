"""Latency of one small-batch solve call (BASELINE configs[0]: batch 64) on the GPU beside the C oracle on the host.
GPU box: python tools/latency_probe.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from oracle import c_oracle
for B, N in ((64, 64), (64, 256), (256, 256), (1024, 256), (4096, 256)):
    b = davo_b200.synthetic.make_distort10(B, N, seed=5, dtype=np.float32)
    obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d).cuda(), torch.from_numpy(b.obs).cuda())
    x0 = torch.from_numpy(b.x0).cuda()
    solver = davo_b200.BFGSSolver(error_threshold=1e-7, iterations=1000).eval()
    ts = []
    for rep in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        info = solver(x0, obj, return_info=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    gpu = float(np.median(ts[3:]))
    t0 = time.perf_counter()
    ref = c_oracle.solve_batch(b, iterations=1000, error_threshold=1e-7)
    cpu = time.perf_counter() - t0
    print(f"B={B} N={N}: GPU {gpu*1e3:.3f} ms per call (max fevals {int(info.evaluations.max())}, mean {float(info.evaluations.float().mean()):.0f}), "
          f"C oracle {cpu*1e3:.1f} ms ({os.cpu_count()} cpus)")
