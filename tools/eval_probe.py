"""davo_eval_cost_grad / davo_line_search on 64K x 256 staged problems (GPU box; run under ncu --metrics ...)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import make_batch
b = make_batch("cfg2", 65536, 0xB200)
obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d).cuda(), torch.from_numpy(b.obs).cuda())
x = torch.from_numpy(b.x0).cuda()
for _ in range(3):
    cost, grad = obj.evaluate(x)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cost, grad = obj.evaluate(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"evaluate 64K x 256: {np.median(ts)*1e3:.0f} us -> {65536*256*16/np.median(ts)/1e6:.0f} GB/s of staged matches")
