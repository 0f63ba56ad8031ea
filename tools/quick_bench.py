"""Time the solve kernel alone (CUDA events, staged inputs resident) and print a behaviour checksum.
Usage (GPU box): python tools/quick_bench.py [cfg2|cfg3|cfg4] [reps] [B]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import SOLVER_KW, algorithmic_flops, default_B, make_batch

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
B = int(sys.argv[3]) if len(sys.argv) > 3 else default_B(cfg)
batch = make_batch(cfg, B, 0xB200)
if batch.model == "distort10":
    obj = davo_b200.DistortionObjective(torch.from_numpy(batch.points_3d), torch.from_numpy(batch.obs))
else:
    obj = davo_b200.JointPoseObjective(torch.from_numpy(batch.points_3d), torch.from_numpy(batch.obs))
solver = davo_b200.BFGSSolver(**SOLVER_KW[cfg]).eval()
x0 = torch.from_numpy(batch.x0).cuda()
for _ in range(3):
    buf = solver.solve_into(x0, obj)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); buf = solver.solve_into(x0, obj); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
it = buf.iterations.cpu().numpy(); fe = buf.evaluations.cpu().numpy()
ms = float(np.median(ts))
fl = algorithmic_flops(batch, fe, it)
print(f"{cfg} B={B}: solve kernel median {ms:.3f} ms (min {min(ts):.3f}) -> {B / ms * 1e3 / 1e6:.3f} M solves/s, "
      f"{fl / ms / 1e9:.2f} TFLOP/s algorithmic; iters sum {it.sum()} fevals sum {fe.sum()} "
      f"converged {float(buf.converged.float().mean()):.4f} cost sum {float(buf.cost.double().sum()):.6e}")
