"""One fused initial-guess launch at 64K rows (for ncu).  GPU box."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
torch.manual_seed(0)
net = davo_b200.CalibrationNetwork(4, 8).cuda().eval()
x = torch.randn(65536, 64, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = net.estimate(x)
torch.cuda.synchronize()
print(float(y.abs().mean()))
