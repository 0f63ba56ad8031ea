"""Compare the GPU and the oracle on the problems of a config that cost the GPU the most evaluations."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import SOLVER_KW, make_batch
from oracle import c_oracle
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
batch = make_batch(cfg, 65536, 0xB200)
obj = davo_b200.DistortionObjective(torch.from_numpy(batch.points_3d), torch.from_numpy(batch.obs))
solver = davo_b200.BFGSSolver(**SOLVER_KW[cfg]).eval()
info = solver(torch.from_numpy(batch.x0), obj, return_info=True)
ev = (info.evaluations - info.iterations).numpy()
idx = np.argsort(-ev)[:6]
sub = batch.slice(0, 1)
for i in idx:
    one = batch.slice(int(i), int(i) + 1)
    o = c_oracle.solve_batch(one, threads=1, **SOLVER_KW[cfg])
    o64 = c_oracle.solve_batch(one.astype(np.float64), threads=1, **SOLVER_KW[cfg])
    print(f"problem {i}: GPU iters {int(info.iterations[i])} fevals {int(info.evaluations[i])} reason {int(info.reason[i])} "
          f"cost {float(info.cost[i]):.3e} | oracle f32 iters {o['iters'][0]} fevals {o['fevals'][0]} reason {o['reason'][0]} "
          f"cost {o['cost'][0]:.3e} | oracle f64 iters {o64['iters'][0]} fevals {o64['fevals'][0]} reason {o64['reason'][0]} cost {o64['cost'][0]:.3e}")
    print("   x0", batch.x0[i], "\n   x_gpu", info.parameters[i].numpy())
