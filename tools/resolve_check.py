"""Debug: with a library built with -DDAVO_EVAL_CAP=6 -DDAVO_EVAL_CAP_EXACT=1 nearly every problem goes through the
straggler (CTA) launch; print a digest and a few rows.  GPU box."""
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import davo_b200
b = davo_b200.synthetic.make_distort10(16384, 256, seed=0xB200, dtype=np.float32)
obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d).cuda(), torch.from_numpy(b.obs).cuda())
info = davo_b200.BFGSSolver(error_threshold=1e-7).eval()(torch.from_numpy(b.x0).cuda(), obj, return_info=True)
h = hashlib.sha256()
for t in (info.parameters, info.cost, info.iterations, info.evaluations, info.reason):
    h.update(t.cpu().numpy().tobytes())
print("digest", h.hexdigest()[:16], "iters mean", float(info.iterations.float().mean()), "evals mean",
      float(info.evaluations.float().mean()), "reasons", np.bincount(info.reason.cpu().numpy(), minlength=4))
print(info.iterations[:12].cpu().numpy(), info.evaluations[:12].cpu().numpy())
