"""Solve config 4 (64K ill-conditioned problems) on the device and print a digest of every output plus the solve
time: builds that must agree bit for bit (e.g. speculative vs sequential straggler line search) print the same digest.
GPU box: DAVO_B200_LIB=... python tools/cfg4_digest.py"""
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import SOLVER_KW, make_batch
b = make_batch("cfg4", 65536, 0xB200)
obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d).cuda(), torch.from_numpy(b.obs).cuda())
solver = davo_b200.BFGSSolver(**SOLVER_KW["cfg4"]).eval()
x0 = torch.from_numpy(b.x0).cuda()
ts = []
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); buf = solver.solve_into(x0, obj); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
h = hashlib.sha256()
for t in (buf.x, buf.cost, buf.iterations, buf.evaluations, buf.reason, buf.converged):
    h.update(t.cpu().numpy().tobytes())
ev = buf.evaluations.cpu().numpy()
print(f"solve {min(ts[1:]):.2f} ms  digest {h.hexdigest()[:16]}  max evals {ev.max()}  handed-off-size problems {(ev > 4096).sum()}")
idx = [60403, 48468] + list(np.argsort(-ev)[:8])
it = buf.iterations.cpu().numpy(); rs = buf.reason.cpu().numpy(); cs = buf.cost.cpu().numpy()
for j in idx:
    print(int(j), "iters", int(it[j]), "evals", int(ev[j]), "reason", int(rs[j]), "cost", float(cs[j]))
