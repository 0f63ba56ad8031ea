"""Time the fused initial-guess kernel against the torch modules (cuBLAS + elementwise kernels) at 64K rows.  GPU box."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
torch.manual_seed(0)
net = davo_b200.CalibrationNetwork(4, 8).cuda().eval()
for B in (64, 4096, 65536, 1 << 20):
    x = torch.randn(B, 64, device="cuda")
    with torch.no_grad():
        for name, fn in (("fused tcgen05", net.estimate), ("torch modules", net.initial_estimator)):
            for _ in range(3):
                fn(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn(x)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"B={B:8d} {name:14s} {ms*1e3:9.1f} us  {B * 2 * (64*256 + 256*256 + 256*45) / ms / 1e9:8.2f} TFLOP/s (fp32-equivalent)")
