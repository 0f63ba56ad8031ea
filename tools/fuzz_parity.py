"""Randomised parity sweep of the solve paths against the C oracle (GPU box): random batch sizes (both sides of the
two-per-warp threshold), match counts (ragged, tiny, large), precisions, weights, iteration caps.
    python tools/fuzz_parity.py [seconds] [seed]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import davo_b200
from oracle import c_oracle
from parity import compare_solves, summary
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t0, runs, worst = time.time(), 0, {}
while time.time() - t0 < budget:
    kind = rng.choice(["d10", "d10big", "joint", "ba"])
    dt = np.float64 if rng.random() < 0.6 else np.float32
    thr = 1e-12 if dt == np.float64 else 1e-5
    iters = int(rng.choice([0, 1, 2, 7, 40, 200]))
    zoom = bool(rng.random() < 0.25)   # the opt-in secant zoom (generic solver on every model)
    if kind == "d10":
        B, N = int(rng.integers(1, 600)), int(rng.choice([1, 2, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 255, 256, 257, 300]))
        b = davo_b200.synthetic.make_distort10(B, N, seed=int(rng.integers(1 << 30)), dtype=dt, random_pose=bool(rng.random() < 0.5))
        w = rng.uniform(0, 2, size=(B, N)).astype(dt) if rng.random() < 0.4 else None
        obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d).cuda(), torch.from_numpy(b.obs).cuda(),
                                            torch.from_numpy(b.pose).cuda(), weights=None if w is None else torch.from_numpy(w).cuda())
        ref = c_oracle.solve("distort10", b.x0, c_oracle.stage(b.points_3d, b.obs, b.pose), None, w, N=N, iterations=iters, error_threshold=thr, zoom_interpolation=zoom)
    elif kind == "d10big":
        B, N = int(rng.integers(14300, 30000)), int(rng.choice([8, 31, 32, 48, 50, 64, 96]))
        b = davo_b200.synthetic.make_distort10(B, N, seed=int(rng.integers(1 << 30)), dtype=dt)
        obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d).cuda(), torch.from_numpy(b.obs).cuda())
        iters = min(iters, 40)
        ref = c_oracle.solve_batch(b, iterations=iters, error_threshold=thr, zoom_interpolation=zoom)
    elif kind == "joint":
        B, N, V = int(rng.integers(1, 200)), int(rng.choice([5, 32, 33, 64, 100, 256])), int(rng.integers(1, 7))
        b = davo_b200.synthetic.make_joint(B, N, V, seed=int(rng.integers(1 << 30)), dtype=dt)
        obj = davo_b200.JointPoseObjective(torch.from_numpy(b.points_3d).cuda(), torch.from_numpy(b.obs).cuda())
        iters = min(iters, 40)
        ref = c_oracle.solve_batch(b, iterations=iters, error_threshold=thr, zoom_interpolation=zoom)
    else:
        V, N = int(rng.integers(2, 7)), int(rng.integers(2, 20))
        if 3 + 3 * N + 6 * (V - 1) > 128: continue
        B = int(rng.integers(1, 300))
        b = davo_b200.synthetic.make_angle_ba(B, N, V, seed=int(rng.integers(1 << 30)), dtype=dt)
        obj = davo_b200.AngleDistanceObjective(torch.from_numpy(b.obs).cuda(), torch.from_numpy(b.weights).cuda())
        iters = min(iters, 25)
        thr = 1e-7
        ref = c_oracle.solve_batch(b, iterations=iters, error_threshold=thr, zoom_interpolation=zoom)
    solver = davo_b200.BFGSSolver(error_threshold=thr, iterations=iters).eval()
    solver.zoom_interpolation = zoom
    info = solver(torch.from_numpy(b.x0).cuda(), obj, return_info=True)
    got = dict(x=info.parameters.cpu().numpy(), cost=info.cost.cpu().numpy(), iters=info.iterations.cpu().numpy(),
               fevals=info.evaluations.cpu().numpy(), reason=info.reason.cpu().numpy())
    m = compare_solves(got, ref, thr)
    runs += 1
    key = f"{kind}/{np.dtype(dt).name}" + ("/secant" if zoom else "")
    cur = worst.setdefault(key, dict(runs=0, steps=1.0, reason=1.0, dtheta=0.0))
    cur["runs"] += 1; cur["steps"] = min(cur["steps"], m["steps_equal"]); cur["reason"] = min(cur["reason"], m["reason_equal"])
    cur["dtheta"] = max(cur["dtheta"], m["dtheta_median"])
    bad = (not np.isfinite(got["x"]).all() and np.isfinite(ref["x"]).all()) or \
          (dt == np.float64 and (m["steps_equal"] < 0.9 or m["dtheta_median"] > 1e-6)) or \
          (dt == np.float32 and (m["steps_equal"] < 0.5 or m["dtheta_median"] > 1e-2))
    if bad:
        print("SUSPECT", kind, "secant" if zoom else "bisect", np.dtype(dt).name, "B", b.B, "N", b.N, "V", b.views, "iters", iters, summary(m), flush=True)
print(f"{runs} runs in {time.time() - t0:.0f} s")
for k, v in sorted(worst.items()):
    print(k, v)
