"""One bundle-adjustment solve at 32K problems, float32, threshold 1e-4 (for ncu).  GPU box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
b = davo_b200.synthetic.make_angle_ba(32768, 8, 4, seed=0xB200, dtype=np.float32)
obj = davo_b200.AngleDistanceObjective(torch.from_numpy(b.obs), torch.from_numpy(b.weights))
solver = davo_b200.BFGSSolver(error_threshold=1e-4, iterations=1000).eval()
x0 = torch.from_numpy(b.x0).cuda()
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); buf = solver.solve_into(x0, obj); e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1):.2f} ms, iters {float(buf.iterations.float().mean()):.1f}, converged {float(buf.converged.float().mean()):.3f}")
