"""GPU results for chosen problem indices of a config (float32 and float64).  python tools/probe_problems.py cfg4 60403 48468"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import SOLVER_KW, make_batch
cfg = sys.argv[1]; idx = [int(a) for a in sys.argv[2:]]
batch = make_batch(cfg, 65536, 0xB200)
for dt in (np.float32, np.float64):
    b = batch.astype(dt)
    sel = lambda a: torch.from_numpy(np.ascontiguousarray(a[idx]))
    obj = davo_b200.DistortionObjective(sel(b.points_3d), sel(b.obs))
    info = davo_b200.BFGSSolver(**SOLVER_KW[cfg]).eval()(sel(b.x0), obj, return_info=True)
    print(os.environ.get("DAVO_B200_LIB", "default")[-24:], dt.__name__, "iters", info.iterations.tolist(), "fevals", info.evaluations.tolist(),
          "reason", info.reason.tolist(), "cost", [f"{c:.3e}" for c in info.cost.tolist()])
