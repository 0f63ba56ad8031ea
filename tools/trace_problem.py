"""Debug: per-step trace of one problem on the GPU (needs a -DDAVO_TRACE=1 build selected with DAVO_B200_LIB)
next to the oracle's totals.  python tools/trace_problem.py cfg4 [index]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from davo_b200 import _lib
from bench import SOLVER_KW, make_batch
from oracle import c_oracle
cfg = sys.argv[1]
batch = make_batch(cfg, 65536, 0xB200)
obj = davo_b200.DistortionObjective(torch.from_numpy(batch.points_3d), torch.from_numpy(batch.obs))
solver = davo_b200.BFGSSolver(**SOLVER_KW[cfg]).eval()
x0 = torch.from_numpy(batch.x0).cuda()
if len(sys.argv) > 2:
    idx = int(sys.argv[2])
else:
    buf = solver.solve_into(x0, obj)
    idx = int((buf.evaluations - buf.iterations).argmax())
trace = torch.zeros(1000, 8, device="cuda")
_lib.lib().davo_debug_trace(ctypes.c_void_p(trace.data_ptr()), ctypes.c_int(idx), ctypes.c_int(1000))
buf = solver.solve_into(x0, obj)
torch.cuda.synchronize()
t = trace.cpu().numpy()
one = batch.slice(idx, idx + 1)
o = c_oracle.solve_batch(one, threads=1, **SOLVER_KW[cfg])
print(f"problem {idx}: GPU iters {int(buf.iterations[idx])} fevals {int(buf.evaluations[idx])} reason {int(buf.reason[idx])} cost {float(buf.cost[idx]):.4e}"
      f" | oracle iters {o['iters'][0]} fevals {o['fevals'][0]} reason {o['reason'][0]} cost {o['cost'][0]:.4e}")
print("x0", batch.x0[idx]); print("truth", batch.truth[idx]); print("x_gpu", buf.x[idx].cpu().numpy()); print("x_oracle", o["x"][0])
print("   k        f0           g0        alpha  probes     |s|        f_new   reuse")
n = int(buf.iterations[idx])
rows = list(range(min(n, 40))) + list(range(max(40, n - 8), n))
for k in rows:
    r = t[k]
    print(f"{int(r[0]):4d} {r[1]:12.5e} {r[2]:12.4e} {r[3]:11.4e} {int(r[4]):4d} {r[5]:11.3e} {r[6]:12.5e} {int(r[7])}")
