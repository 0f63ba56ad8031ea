"""Work-distribution analysis of one solve launch: how far is the makespan from total work / resident
warps?  Usage (GPU box): python tools/analyze_tail.py cfg4 [B]"""
import heapq
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import SOLVER_KW, make_batch, default_B

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else default_B(cfg)
batch = make_batch(cfg, B, 0xB200)
if batch.model == "distort10":
    obj = davo_b200.DistortionObjective(torch.from_numpy(batch.points_3d), torch.from_numpy(batch.obs))
else:
    obj = davo_b200.JointPoseObjective(torch.from_numpy(batch.points_3d), torch.from_numpy(batch.obs))
solver = davo_b200.BFGSSolver(**SOLVER_KW[cfg]).eval()
x0 = torch.from_numpy(batch.x0).cuda()
for _ in range(2):
    buf = solver.solve_into(x0, obj)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); buf = solver.solve_into(x0, obj); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
it = buf.iterations.cpu().numpy().astype(np.int64)
fe = buf.evaluations.cpu().numpy().astype(np.int64)
ev = fe - it  # evaluations actually executed (upper bound: reuse skips one per step)
print(f"{cfg}: B={B} kernel {ms:.2f} ms; iters mean {it.mean():.1f} max {it.max()}; executed evals mean {ev.mean():.1f} "
      f"p50 {np.median(ev):.0f} p99 {np.quantile(ev, .99):.0f} max {ev.max()}; reasons {np.bincount(buf.reason.cpu().numpy(), minlength=4)}")
work = ev + 0.6 * it  # eval-equivalents: a BFGS update costs ~0.6 of an evaluation (ncu instruction counts)
for warps in (148 * 20, 148 * 16, 148 * 8):
    # greedy list scheduling in queue order, constant speed per warp
    heap = [0.0] * warps
    heapq.heapify(heap)
    for w in work:
        t = heapq.heappop(heap)
        heapq.heappush(heap, t + w)
    makespan = max(heap)
    print(f"  {warps} warps: ideal {work.sum() / warps:.0f}, greedy makespan {makespan:.0f} eval-units "
          f"(x{makespan / (work.sum() / warps):.2f}); longest problem {work.max():.0f}")
