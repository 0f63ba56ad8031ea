"""Randomised sweep of the training-mode / differentiable solve against its numpy restatement (GPU box): random camera
objectives and shapes, iteration caps, thresholds, return_second_last; float64.  Compares the returned parameters and
d sum(w * x_out) / d x0 (oracle/train_oracle.py is itself pinned to the reference's autograd by tests/golden/training.npz).
    python tools/fuzz_training.py [seconds] [seed]"""
import os, sys, time
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
sys.path.insert(0, os.path.join(root, "tests"))
import davo_b200
from oracle import c_oracle, train_oracle
from test_training_oracle import relative_gradient_error

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t0, runs, worst, suspects = time.time(), 0, {}, 0
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
while time.time() - t0 < budget:
    kind = str(rng.choice(["d10", "joint", "ba"]))
    B = int(rng.integers(1, 6))
    iters = int(rng.choice([1, 2, 3, 5, 8, 12]))
    thr = float(rng.choice([1e-12, 1e-6, 1e-3]))
    second = bool(rng.random() < 0.3)
    seed = int(rng.integers(1 << 30))
    if kind == "d10":
        N = int(rng.choice([5, 16, 20, 33, 64]))
        b = davo_b200.synthetic.make_distort10(B, N, seed=seed, dtype=np.float64, random_pose=bool(rng.random() < 0.5))
        obj = davo_b200.DistortionObjective(T(b.points_3d), T(b.obs), T(b.pose))
        prob = train_oracle.Problem("distort10", c_oracle.stage(b.points_3d, b.obs, b.pose), N=N)
    elif kind == "joint":
        N, V = int(rng.choice([8, 16, 33])), int(rng.integers(1, 4))
        b = davo_b200.synthetic.make_joint(B, N, V, seed=seed, dtype=np.float64)
        obj = davo_b200.JointPoseObjective(T(b.points_3d), T(b.obs))
        prob = train_oracle.Problem("joint", b.points_3d, b.obs, N=N, V=V)
    else:
        V, N = int(rng.integers(2, 5)), int(rng.integers(3, 9))
        b = davo_b200.synthetic.make_angle_ba(B, N, V, seed=seed, dtype=np.float64)
        obj = davo_b200.AngleDistanceObjective(T(b.obs), T(b.weights))
        prob = train_oracle.Problem("angle_ba", b.obs, None, b.weights, N=N, V=V)
    w = rng.normal(size=b.x0.shape)
    want_x, want_g = train_oracle.solve_with_grad(prob, b.x0, w, error_threshold=thr, iterations=iters, second_last=second)
    solver = davo_b200.BFGSSolver(drop_path_p=0.0, training_iterations=iters, training_error_threshold=thr,
                                  return_second_last=second).train()
    x0 = T(b.x0).requires_grad_(True)
    x = solver(x0, obj)
    (grad,) = torch.autograd.grad((x * T(w)).sum(), x0)
    dx = float(np.abs(x.detach().cpu().numpy() - want_x).max() / (1.0 + np.abs(want_x).max()))
    dg = relative_gradient_error(grad.cpu().numpy(), want_g)
    runs += 1
    cur = worst.setdefault(kind, dict(runs=0, dx=0.0, dg=0.0))
    cur["runs"] += 1; cur["dx"] = max(cur["dx"], dx); cur["dg"] = max(cur["dg"], dg)
    # a chain whose gradient has exploded (|g| >> 1e4) is compared loosely, as the reference is against itself
    loose = np.abs(want_g).max() > 1e4
    if dx > 1e-6 or (dg > (1e-2 if loose else 1e-5)) or not np.isfinite(dg):
        suspects += 1
        print("SUSPECT", kind, "B", B, "N", b.N, "V", b.views, "iters", iters, "thr", thr, "second", second, "seed", seed,
              f"dx {dx:.2e} dg {dg:.2e} |g| {np.abs(want_g).max():.2e}", flush=True)
print(f"{runs} runs in {time.time() - t0:.0f} s, {suspects} suspects")
for k, v in sorted(worst.items()):
    print(k, v)
