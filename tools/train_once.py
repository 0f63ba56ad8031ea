"""One training-mode solve (recording forward) and its backward pass on the entry script's objective, 4096 problems,
float64, 50-iteration cap (for ncu).  GPU box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
b = davo_b200.synthetic.make_angle_ba(4096, 8, 4, seed=3, dtype=np.float64)
obs = torch.from_numpy(b.obs).cuda().requires_grad_(True)
obj = davo_b200.AngleDistanceObjective(obs, torch.from_numpy(b.weights).cuda())
solver = davo_b200.BFGSSolver(drop_path_p=0.0, training_error_threshold=1e-3, training_iterations=50).train()
x0 = torch.from_numpy(b.x0).cuda().requires_grad_(True)
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
info = solver(x0, obj, return_info=True)
e[1].record()
info.parameters.square().sum().backward()
e[2].record()
torch.cuda.synchronize()
print(f"forward {e[0].elapsed_time(e[1]):.2f} ms, backward {e[1].elapsed_time(e[2]):.2f} ms, mean steps "
      f"{float(info.iterations.float().mean()):.1f}, |d/dx0| {float(x0.grad.abs().max()):.3e}, |d/dobs| {float(obs.grad.abs().max()):.3e}")
