"""Population statistics of config-4 solves on the GPU for the library selected with DAVO_B200_LIB (A/B builds of
csrc/, see its Makefile), dumped to gpurun_out/ for comparison with the oracle.  python tools/variant_stats.py tag [B]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
from bench import SOLVER_KW, make_batch
tag = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
batch = make_batch("cfg4", 65536, 0xB200).slice(0, B)
for dt in (np.float32, np.float64):
    b = batch.astype(dt)
    obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d), torch.from_numpy(b.obs))
    info = davo_b200.BFGSSolver(**SOLVER_KW["cfg4"]).eval()(torch.from_numpy(b.x0), obj, return_info=True)
    it, rs, fe = info.iterations.numpy(), info.reason.numpy(), info.evaluations.numpy()
    print(f"{tag:10s} {dt.__name__}: mean iters {it.mean():.2f} capped {(rs == 2).mean():.4f} fevals {fe.mean():.1f} "
          f"reasons {np.bincount(rs, minlength=4)}")
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez(f"gpurun_out/cfg4_{tag}_{dt.__name__}.npz", iters=it, reason=rs, fevals=fe, x=info.parameters.numpy(),
             cost=info.cost.numpy())
