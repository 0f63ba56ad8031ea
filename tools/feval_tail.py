"""How many problems of configs 2 and 4 pass a given number of reference-equivalent evaluations (the hand-off cap's
trade-off).  GPU box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import SOLVER_KW, make_batch
for cfg in ("cfg2", "cfg4"):
    b = make_batch(cfg, 65536, 0xB200)
    obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d).cuda(), torch.from_numpy(b.obs).cuda())
    info = davo_b200.BFGSSolver(**SOLVER_KW[cfg]).eval()(torch.from_numpy(b.x0).cuda(), obj, return_info=True)
    fe = info.evaluations.cpu().numpy()
    print(cfg, "max", fe.max(), {t: int((fe > t).sum()) for t in (1024, 1536, 2048, 2560, 3072, 3584, 4096, 8192, 16384)})
