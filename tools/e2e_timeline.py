"""Where one host-input step goes: copies, per-chunk compute, the tail after the last copy, the host's queueing time.
GPU box: python tools/e2e_timeline.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import SOLVER_KW, make_batch
B = 65536
batch = make_batch("cfg2", B, 0xB200)
h_pts, h_obs, h_x0 = (torch.from_numpy(a).pin_memory() for a in (batch.points_3d, batch.obs, batch.x0))
solver = davo_b200.BFGSSolver(**SOLVER_KW["cfg2"]).eval()
for rep in range(8):
    torch.cuda.synchronize()
    solver._trace = tr = {}
    t0 = time.perf_counter()
    obj = davo_b200.DistortionObjective(h_pts, h_obs)
    t_obj = time.perf_counter()
    pend = solver.submit(h_x0, obj, return_info=True)
    t_sub = time.perf_counter()
    info = pend.result()
    t_res = time.perf_counter()
    torch.cuda.synchronize()
    if rep < 3:
        continue
    st = tr["start"]
    landed = [st.elapsed_time(e) for e in tr["landed"]]
    done = [st.elapsed_time(e) for e in tr["done"]]
    sizes = [hi - lo for lo, hi in tr["spans"]]
    print(f"rep {rep}: total {1e3*(t_res-t0):.2f} ms | host: objective {1e3*(t_obj-t0):.2f}, submit {1e3*(t_sub-t_obj):.2f}, "
          f"wait {1e3*(t_res-t_sub):.2f} | device (from the start event): last copy landed {landed[-1]:.2f}, last chunk done "
          f"{done[-1]:.2f} (max {max(done):.2f})")
    if rep == 7:
        print("chunk sizes", sizes)
        print("landed", [round(v, 2) for v in landed])
        print("done  ", [round(v, 2) for v in done])
        print("lag   ", [round(d - l, 2) for d, l in zip(done, landed)])
