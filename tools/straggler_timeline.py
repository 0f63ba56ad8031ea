"""One config-4 solve with a -DDAVO_TIMELINE=1 build: the kernels print hand-off / early-CTA timestamps (ns).
GPU box: DAVO_B200_LIB=.../libdavo_b200_tl.so python tools/straggler_timeline.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import SOLVER_KW, make_batch
b = make_batch("cfg4", 65536, 0xB200)
obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d).cuda(), torch.from_numpy(b.obs).cuda())
x0 = torch.from_numpy(b.x0).cuda()
solver = davo_b200.BFGSSolver(**SOLVER_KW["cfg4"]).eval()
for rep in range(2):
    torch.cuda.synchronize()
    print(f"==== solve {rep}", flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); buf = solver.solve_into(x0, obj); e1.record(); torch.cuda.synchronize()
    print(f"==== solve {rep}: {e0.elapsed_time(e1):.2f} ms", flush=True)
