"""Time ONE config-4 problem on the straggler route: a 16K-problem batch of well-conditioned problems (config 2) with
row 0 replaced by the chosen config-4 problem; the launch time minus the plain batch's time is that problem's
re-solve.  GPU box: DAVO_B200_LIB=... python tools/straggler_probe.py 32904 51471"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import SOLVER_KW, make_batch
hard = make_batch("cfg4", 65536, 0xB200)
easy = make_batch("cfg2", 16384, 0xB200)
solver = davo_b200.BFGSSolver(**SOLVER_KW["cfg4"]).eval()


def run(pts, obs, x0):
    obj = davo_b200.DistortionObjective(torch.from_numpy(pts).cuda(), torch.from_numpy(obs).cuda())
    x0 = torch.from_numpy(x0).cuda()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); buf = solver.solve_into(x0, obj); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:]), buf


base, _ = run(easy.points_3d, easy.obs, easy.x0)
print(f"plain batch {base:.2f} ms")
for j in [int(a) for a in sys.argv[1:]]:
    pts, obs, x0 = easy.points_3d.copy(), easy.obs.copy(), easy.x0.copy()
    pts[0], obs[0], x0[0] = hard.points_3d[j], hard.obs[j], hard.x0[j]
    t, buf = run(pts, obs, x0)
    print(f"problem {j}: {t:.2f} ms (+{t - base:.2f}), evals {int(buf.evaluations[0])}, iters {int(buf.iterations[0])}")
