"""Time the training-mode solve (forward with recording + backward) on the entry script's objective.  GPU box."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
for B, dt in ((64, np.float64), (4096, np.float64), (4096, np.float32)):
    b = davo_b200.synthetic.make_angle_ba(B, 8, 4, seed=3, dtype=dt)
    obj = davo_b200.AngleDistanceObjective(torch.from_numpy(b.obs).cuda(), torch.from_numpy(b.weights).cuda())
    for kw in (dict(drop_path_p=0.1, training_error_threshold=1e-3), dict(drop_path_p=0.0, training_error_threshold=1e-3, training_iterations=50)):
        solver = davo_b200.BFGSSolver(**kw).train()
        x0 = torch.from_numpy(b.x0).cuda().requires_grad_(True)
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            info = solver(x0, obj, return_info=True)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            info.parameters.square().sum().backward()
            torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"B={B} {np.dtype(dt).name} {kw}: forward {1e3*(t1-t0):.2f} ms, backward {1e3*(t2-t1):.2f} ms, mean steps "
              f"{float(info.iterations.float().mean()):.1f}, grad finite {bool(torch.isfinite(x0.grad).all())}")
