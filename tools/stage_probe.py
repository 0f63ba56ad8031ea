"""Achieved HBM bandwidth of davo_stage_matches with and without a per-problem pose.  GPU box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
B, N = 65536, 256
pts = torch.randn(B, N, 3, device="cuda").abs() + 1.0
obs = torch.randn(B, N, 2, device="cuda")
pose = 0.1 * torch.randn(B, 6, device="cuda")
for name, p in (("identity", None), ("pose", pose)):
    for dt in (torch.float32, torch.float64):
        a, o, q = pts.to(dt), obs.to(dt), None if p is None else p.to(dt)
        for _ in range(3): obj = davo_b200.DistortionObjective(a, o, q)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); obj = davo_b200.DistortionObjective(a, o, q); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        esz = 4 if dt == torch.float32 else 8
        gb = B * N * 9 * esz / 1e9
        print(f"{name} {dt}: {np.median(ts)*1e3:.1f} us -> {gb / (np.median(ts) * 1e-3):.0f} GB/s")
