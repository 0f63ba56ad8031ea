"""Generate tests/golden/*.npz from the UNMODIFIED reference — TEST INFRASTRUCTURE, build container only.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden [--only NAME ...]

Every fixture is an output of the reference's own code (ref_harness.py says which functions) on
inputs produced by davo_b200.synthetic from a seed (large cases: only the seed, the generator
arguments and a sha256 of the inputs are stored) or stored verbatim (small cases).  The GPU box has
no /root/reference; tests there compare against these files.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import davo_b200  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

syn = davo_b200.synthetic
GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (generator, generator kwargs, torch dtype, solver kwargs)
SOLVE_CASES = {
    # gate G64: float64 against the float64 reference
    "solve_cfg2_f64": ("make_distort10", dict(B=4096, N=256, seed=0xB200), "float64",
                       dict(error_threshold=1e-12, iterations=1000)),
    # gate G32: float32 at a threshold every problem reaches before the fp32 noise floor
    "solve_cfg2_f32": ("make_distort10", dict(B=4096, N=256, seed=0xB200), "float32",
                       dict(error_threshold=1e-5, iterations=1000)),
    # the driver's eval threshold (networks/calibration_network.py:44): chaotic in fp32, kept for the record
    "solve_cfg2_f32_thr1e-7": ("make_distort10", dict(B=4096, N=256, seed=0xB200), "float32",
                               dict(error_threshold=1e-7, iterations=1000)),
    # noisy observations: optimum cost ~ 2 N sigma^2, retire on step size / cap
    "solve_cfg2_noisy_f64": ("make_distort10", dict(B=256, N=256, seed=0xB201, noise=1e-3), "float64",
                             dict(error_threshold=1e-12, iterations=300)),
    "solve_cfg2_pose_f64": ("make_distort10", dict(B=128, N=96, seed=0xB202, random_pose=True), "float64",
                            dict(error_threshold=1e-12, iterations=1000)),
    "solve_cfg3_f64": ("make_joint", dict(B=4096, N=256, V=4, seed=0xB200), "float64",
                       dict(error_threshold=1e-10, iterations=1000)),
    "solve_cfg3_f32": ("make_joint", dict(B=4096, N=256, V=4, seed=0xB200), "float32",
                       dict(error_threshold=1e-5, iterations=1000)),
    "solve_cfg3_small_f64": ("make_joint", dict(B=64, N=40, V=2, seed=0xB203), "float64",
                             dict(error_threshold=1e-10, iterations=1000)),
    "solve_cfg4_f64": ("make_distort10", dict(B=384, N=256, seed=0xB204, ill_conditioned=True, pathological=0.02),
                       "float64", dict(error_threshold=1e-10, iterations=1000)),
    "solve_cfg4_f32": ("make_distort10", dict(B=384, N=256, seed=0xB204, ill_conditioned=True, pathological=0.02),
                       "float32", dict(error_threshold=1e-5, iterations=1000)),
    # the entry script's bundle-adjustment objective (M=4 views, N=8 points, n=45).  The
    # error is a sum of angles (not squares): BFGS crawls along its kinks for ~900 iterations and single
    # trajectories are chaotic even in float64, so the tight gate is a run capped at 30 accepted steps and the
    # full-length runs are compared at population level against the reference's own band.
    "solve_ba_f64_30steps": ("make_angle_ba", dict(B=64, N=8, V=4, seed=0xB205), "float64",
                             dict(error_threshold=1e-7, iterations=30)),
    "solve_ba_f32_30steps": ("make_angle_ba", dict(B=64, N=8, V=4, seed=0xB205), "float32",
                             dict(error_threshold=1e-7, iterations=30)),
    "solve_ba_f64": ("make_angle_ba", dict(B=256, N=8, V=4, seed=0xB205), "float64",
                     dict(error_threshold=1e-7, iterations=1000)),
    "solve_ba_f32": ("make_angle_ba", dict(B=256, N=8, V=4, seed=0xB205), "float32",
                     dict(error_threshold=1e-4, iterations=1000)),
    "solve_ba_small_f64": ("make_angle_ba", dict(B=32, N=5, V=2, seed=0xB206), "float64",
                           dict(error_threshold=1e-7, iterations=40)),
    # more than 64 parameters (3 views x 20 points: n = 75; 4 views x 30 points: n = 111): four components per lane
    "solve_ba_n75_f64": ("make_angle_ba", dict(B=32, N=20, V=3, seed=0xB208), "float64",
                         dict(error_threshold=1e-7, iterations=30)),
    "solve_ba_n111_f64": ("make_angle_ba", dict(B=16, N=30, V=4, seed=0xB209), "float64",
                          dict(error_threshold=1e-7, iterations=25)),
}


def _save(name, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)", flush=True)


def gen_solve(name):
    gen, gkw, dt, skw = SOLVE_CASES[name]
    np_dt = np.float64 if dt == "float64" else np.float32
    batch = getattr(syn, gen)(dtype=np_dt, **gkw)
    obj = rh.make_objective(batch, getattr(torch, dt))
    t = time.time()
    r = rh.reference_solve(obj, torch.as_tensor(batch.x0), **skw)
    dt_s = time.time() - t
    print(f"  {name}: reference {dt_s:.1f}s, {batch.B / dt_s:.1f} solves/s, iters mean {r['iters'].mean():.1f}, "
          f"fevals mean {r['fevals'].mean():.1f}, reasons {np.bincount(r['reason'], minlength=4)}", flush=True)
    # The reference's own reproducibility band: the same problems with the order of
    # the matches permuted, i.e. nothing but a different floating-point summation order.
    perm = np.random.default_rng(0x5EED).permutation(batch.N)
    if batch.model == "angle_ba":
        # the world points are parameters here: permute them in the observations AND in the parameter vector,
        # and undo the permutation on the result
        cols = np.arange(batch.n)
        cols[3:3 + 3 * batch.N] = 3 + (3 * perm[:, None] + np.arange(3)[None, :]).reshape(-1)
        pb = syn.CalibrationBatch(batch.model, None, batch.obs[:, :, perm], None, batch.x0[:, cols],
                                  batch.truth[:, cols], batch.views, batch.weights[:, :, perm])
        rp = rh.reference_solve(rh.make_objective(pb, getattr(torch, dt)), torch.as_tensor(pb.x0), **skw)
        back = np.empty_like(cols)
        back[cols] = np.arange(batch.n)
        rp["x"] = rp["x"][:, back]
    else:
        pb = syn.CalibrationBatch(batch.model, batch.points_3d[:, perm],
                                  batch.obs[:, perm] if batch.model == "distort10" else batch.obs[:, :, perm],
                                  batch.pose, batch.x0, batch.truth, batch.views)
        rp = rh.reference_solve(rh.make_objective(pb, getattr(torch, dt)), torch.as_tensor(batch.x0), **skw)
    print(f"  {name}: self-consistency under a permutation of the matches: identical steps "
          f"{(rp['iters'] == r['iters']).mean():.4f}, identical reason {(rp['reason'] == r['reason']).mean():.4f}",
          flush=True)
    meta = dict(generator=gen, generator_kwargs=gkw, dtype=dt, solver_kwargs=skw, digest=batch.digest(),
                reference_seconds=dt_s, reference_threads=torch.get_num_threads(), torch=torch.__version__)
    _save(name, meta=json.dumps(meta), x=r["x"], cost=r["cost"], iters=r["iters"], fevals=r["fevals"],
          reason=r["reason"], converged=r["converged"], perm_x=rp["x"], perm_cost=rp["cost"],
          perm_iters=rp["iters"], perm_fevals=rp["fevals"], perm_reason=rp["reason"])


def gen_camera_model():
    rng = np.random.default_rng(11)
    B, N = 6, 12
    pts = syn._points(rng, B, N, 0.7)
    th = np.concatenate([syn._intrinsics(rng, B, k_scale=(0.1, 0.02, 0.004), p_scale=0.02),
                         0.3 * rng.standard_normal((B, 3)), 0.3 * rng.standard_normal((B, 3))], axis=1)
    th[:, syn.S] = 0.05 * rng.standard_normal(B)
    pts[0, 0] = [0.3, -0.2, 0.0]  # with t_z = 0 and no rotation this hits the z' == 0 guard (:57)
    th[0, 10:] = 0.0
    u, v = rh.reference_project(torch.tensor(pts), torch.tensor(th))
    J = rh.reference_project_autograd_jacobian(torch.tensor(pts[1:]), torch.tensor(th[1:]))
    u32, v32 = rh.reference_project(torch.tensor(pts, dtype=torch.float32), torch.tensor(th, dtype=torch.float32))
    # cost + gradient of both calibration objectives by autograd of the reference forward
    b = syn.make_distort10(16, 48, seed=21, dtype=np.float64, random_pose=True)
    x = b.x0 + 0.01 * rng.standard_normal(b.x0.shape)
    f, g = rh.reference_cost_grad(rh.make_objective(b), torch.tensor(x))
    j = syn.make_joint(8, 24, 3, seed=22, dtype=np.float64)
    xj = j.x0 + 0.01 * rng.standard_normal(j.x0.shape)
    fj, gj = rh.reference_cost_grad(rh.make_objective(j), torch.tensor(xj))
    _save("camera_model", points_3d=pts, params16=th, u=u, v=v, u32=u32, v32=v32, J_autograd=J,
          d10_points=b.points_3d, d10_obs=b.obs, d10_pose=b.pose, d10_x=x, d10_cost=f, d10_grad=g,
          joint_points=j.points_3d, joint_obs=j.obs, joint_x=xj, joint_cost=fj, joint_grad=gj)


def gen_angle_ba():
    """Cost, autograd gradient and line search of the entry script's objective (networks/calibration_network.py:
    58-67) from the reference's own functions, at several (views, points) shapes; rows 1-4 of each case exercise
    the Taylor branches of sin(x)/x and (1-cos x)/x^2, the negative-f branch of elu and a zero rotation."""
    rng = np.random.default_rng(15)
    out = {}
    for tag, (V, N) in (("a", (4, 8)), ("b", (2, 5)), ("c", (3, 11)), ("d", (6, 7)), ("e", (3, 20)), ("f", (4, 30))):
        b = syn.make_angle_ba(12, N, V, seed=40 + V, dtype=np.float64)
        x = b.x0.copy()
        x[1, -3 * (V - 1):] *= 0.1
        x[2, -3 * (V - 1):] *= 1e-3
        x[3, 0] = -0.7
        x[4, -3 * (V - 1):] = 0.0
        obj = rh.make_objective(b)
        fix = obj.check_against_unbatched_reference(torch.tensor(x))
        assert fix <= 1e-12, fix  # the keepdim fix changes nothing for a single problem
        f, g = rh.reference_cost_grad(obj, torch.tensor(x))
        b32 = b.astype(np.float32)
        f32, g32 = rh.reference_cost_grad(rh.make_objective(b32, torch.float32), torch.tensor(x.astype(np.float32)))
        d = -g * rng.choice([1e-3, 1e-2, 0.1], size=(12, 1))
        a, p, _, _ = rh.reference_line_search(obj, torch.tensor(x), torch.tensor(d), strong=True)
        out.update({f"{tag}_obs": b.obs, f"{tag}_vis": b.weights, f"{tag}_x": x, f"{tag}_cost": f, f"{tag}_grad": g,
                    f"{tag}_cost32": f32, f"{tag}_grad32": g32, f"{tag}_d": d, f"{tag}_alpha": a, f"{tag}_probes": p,
                    f"{tag}_unbatched_diff": np.float64(fix)})
    _save("angle_ba", **out)


def gen_bfgs_update():
    rng = np.random.default_rng(12)
    out = {}
    for n in (3, 10, 34):
        k = 8
        A = rng.standard_normal((k, n, n))
        H = A @ A.transpose(0, 2, 1) / n + np.eye(n)
        s = rng.standard_normal((k, n))
        y = np.einsum("ki,kij->kj", s, H) + 0.3 * rng.standard_normal((k, n))  # mostly positive curvature
        y[0] = -s[0]  # negative curvature: update skipped
        y[1] = 0.0    # zero curvature: update skipped
        out[f"H{n}"], out[f"s{n}"], out[f"y{n}"] = H, s, y
        out[f"Hout{n}"] = rh.reference_bfgs_update(torch.tensor(H), torch.tensor(s), torch.tensor(y))
        out[f"scale{n}"] = rh.reference_initial_scale(torch.tensor(s), torch.tensor(y))
        out[f"Hout{n}_f32"] = rh.reference_bfgs_update(*[torch.tensor(a, dtype=torch.float32) for a in (H, s, y)])
    # the literal known-answer vectors of tests/autograd_solvers/test_bfgs_solver.py:307-332
    s = np.array([[-1.26262069, -0.78272035, 0.98543104]])
    y = np.array([[0.15339519, -0.28944666, 0.54194925]])
    H = np.array([[[2.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 3.0]]])
    out["kat_s"], out["kat_y"], out["kat_H"] = s, y, H
    out["kat_Hout"] = rh.reference_bfgs_update(torch.tensor(H), torch.tensor(s), torch.tensor(y))
    _save("bfgs_update", **out)


def gen_line_search():
    rng = np.random.default_rng(13)
    out = {}
    # (a) the reference tests' |x - target| cases, tests/autograd_solvers/line_search/test_wolffe_conditions.py:214-305
    targets = np.array([[1.0, 1.0], [10.0, 10.0], [0.25, 0.25], [-9.7, 2.2]])
    dirs = np.array([[10.0, 0.0], [0.2, 0.1], [1.0, 1.0], [1.0, 0.0]])
    for dt in ("float32", "float64"):
        tdt = getattr(torch, dt)
        tg = torch.tensor(targets, dtype=tdt)

        def dist(x, mask, tg=tg):
            return torch.linalg.vector_norm(x - tg[mask.reshape(-1)], dim=-1)

        for strong in (False, True):
            a, p, f0, g = rh.reference_line_search(dist, torch.zeros(4, 2, dtype=tdt), torch.tensor(dirs, dtype=tdt),
                                                   strong=strong)
            out[f"dist_alpha_{dt}_{int(strong)}"], out[f"dist_probes_{dt}_{int(strong)}"] = a, p
    out["dist_targets"], out["dist_dirs"] = targets, dirs
    # (b) analytic objectives from random points along random (mostly descent) directions, c1=.1, c2=.6 as in :152-211
    for name, n in (("sphere", 4), ("log_sphere", 3), ("rosenbrock", 2), ("cosine", 4), ("x2_sine", 2)):
        x = rng.normal(0.0, 2.0, size=(24, n))
        f0, g = rh.reference_cost_grad(rh.ANALYTIC[name], torch.tensor(x))
        d = -g * rng.uniform(0.01, 3.0, size=(24, 1)) + 0.05 * rng.standard_normal((24, n))
        out[f"{name}_x"], out[f"{name}_d"] = x, d
        for strong in (False, True):
            a, p, _, _ = rh.reference_line_search(rh.ANALYTIC[name], torch.tensor(x), torch.tensor(d),
                                                  sufficient_decrease=0.1, curvature=0.6, strong=strong)
            out[f"{name}_alpha_{int(strong)}"], out[f"{name}_probes_{int(strong)}"] = a, p
    # (c) the calibration objective, default c1/c2, steepest descent and a scaled-up direction
    b = syn.make_distort10(32, 64, seed=31, dtype=np.float64)
    obj = rh.make_objective(b)
    x = torch.tensor(b.x0)
    f0, g = rh.reference_cost_grad(obj, x)
    d = -g * rng.choice([1e-3, 1e-2, 1.0], size=(32, 1))
    a, p, _, _ = rh.reference_line_search(obj, x, torch.tensor(d), strong=True)
    out.update(d10_points=b.points_3d, d10_obs=b.obs, d10_pose=b.pose, d10_x=b.x0, d10_d=d, d10_alpha=a, d10_probes=p)
    _save("line_search", **out)


def gen_analytic_solves():
    rng = np.random.default_rng(14)
    out = {}
    cases = {
        # starts follow tests/autograd_solvers/test_bfgs_solver.py:49-97,203-230
        "sphere": np.concatenate([[[1.1, 2.3, 0.0, 0.0]], rng.normal(0, 1, (15, 4))]),
        "sphere_offset": np.concatenate([[[1.1, 2.3]], rng.normal(0, 1, (7, 2))]),
        "log_sphere": np.concatenate([[[-0.3, 0.8]], [[-1700.3, 24942.8]], rng.normal(0, 3, (6, 2))]),
        "rosenbrock": np.concatenate([[[-1.2, 1.0]], rng.uniform(-2, 2, (15, 2))]),
        "cosine": np.concatenate([[[0.03, -18.8, 23.8, 19.0]], rng.normal(0, 5, (7, 4))]),
        "x2_sine": np.array([[17.8885, 35.7771], [-7.641, -7.641], [10.288, -10.288], [9.232, 9.232], [-18.025, 6.0083]]),
    }
    for name, x0 in cases.items():
        for dt in ("float64", "float32"):
            r = rh.reference_solve(rh.ANALYTIC[name], torch.tensor(x0, dtype=getattr(torch, dt)),
                                   error_threshold=1e-6, iterations=1000)
            for k in ("x", "cost", "iters", "fevals", "reason"):
                out[f"{name}_{dt}_{k}"] = r[k]
        out[f"{name}_x0"] = x0
    _save("analytic_solves", **out)


def gen_training():
    """The differentiable / training-mode solve: the UNMODIFIED reference BFGSSolver in
    train() mode, `parameters.requires_grad` (so every gradient evaluation is made with create_graph=True,
    bfgs_solver.py:85,133-135), drop_path_p = 0 (its random retirement cannot be reproduced bit for bit) and both
    settings of return_second_last.  Stored: the returned parameters and d(sum(w * x_out))/d x0 for a random w —
    for the camera objectives also the gradient with respect to the observations."""
    bfgs_mod, _, _, _ = rh._import_reference()
    rng = np.random.default_rng(16)
    out = {}
    cases = {}
    for name, x0 in (("rosenbrock", rng.uniform(-1.5, 1.5, (6, 2))), ("log_sphere", rng.normal(0, 2, (6, 3))),
                     ("sphere", rng.normal(0, 1, (5, 4))), ("cosine", rng.normal(0, 3, (6, 4)))):
        cases[name] = (rh.ANALYTIC[name], x0, None)
    b = syn.make_distort10(6, 20, seed=51, dtype=np.float64, random_pose=True)
    out["d10_points"], out["d10_obs"], out["d10_pose"] = b.points_3d, b.obs, b.pose
    cases["d10"] = (rh.make_objective(b), b.x0, "obs")
    j = syn.make_joint(4, 16, 2, seed=52, dtype=np.float64)
    out["joint_points"], out["joint_obs"] = j.points_3d, j.obs
    cases["joint"] = (rh.make_objective(j), j.x0, "obs")
    a = syn.make_angle_ba(6, 6, 3, seed=53, dtype=np.float64)
    out["ba_obs"], out["ba_vis"] = a.obs, a.weights
    cases["ba"] = (rh.make_objective(a), a.x0, "obs")
    settings = {"k3": dict(training_iterations=3, training_error_threshold=1e-12),
                "k8": dict(training_iterations=8, training_error_threshold=1e-12),
                "k25thr": dict(training_iterations=25, training_error_threshold=1e-3),
                "k8second": dict(training_iterations=8, training_error_threshold=1e-12, return_second_last=True)}
    for cname, (obj, x0, data_attr) in cases.items():
        out[f"{cname}_x0"] = x0
        w = rng.standard_normal(x0.shape)
        out[f"{cname}_w"] = w
        for sname, skw in settings.items():
            solver = bfgs_mod.BFGSSolver(drop_path_p=0.0, **skw).train()
            xg = torch.tensor(x0, requires_grad=True)
            data = None
            if data_attr is not None:
                data = getattr(obj, data_attr).clone().requires_grad_(True)
                setattr(obj, data_attr, data)
            xo = solver(xg, obj)
            loss = (xo * torch.tensor(w)).sum()
            grads = torch.autograd.grad(loss, [xg] + ([data] if data is not None else []), allow_unused=True)
            out[f"{cname}_{sname}_x"] = xo.detach().numpy()
            out[f"{cname}_{sname}_grad_x0"] = grads[0].numpy()
            if data is not None:
                out[f"{cname}_{sname}_grad_{data_attr}"] = grads[1].numpy()
                setattr(obj, data_attr, data.detach())
            print(f"  training {cname}/{sname}: |x - x0| {np.abs(xo.detach().numpy() - x0).max():.3e}, "
                  f"|grad| {np.abs(grads[0].numpy()).max():.3e}", flush=True)
    out["meta"] = json.dumps(dict(settings=settings, torch=torch.__version__))
    _save("training", **out)


def gen_interpolate_alpha():
    """utils/func_interpolate_alpha.py run as is: the reference tests' known answers (tests/utils/
    test_interpolate_alpha.py:29-62) plus random vectors (interior zeros, forced bisections, equal values), forward and
    the custom backward, float32 and float64."""
    sys.path.insert(0, "/root/reference")
    from deep_attention_visual_odometry.utils import interpolate_alpha
    rng = np.random.default_rng(17)
    out = {"kat_in": np.array([[-1.0, 2.0, 0.8, -0.4], [-1.0, 2.0, 0.2, 0.4], [-3.0, -1.0, 6.0, 1.0],
                               [16.0, 2.0, 4.0, 4.0]]),
           "kat_out": np.array([1.0, 0.5, -2.0, 9.0])}
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        k = 4096
        a1, a2, v1 = rng.standard_normal(k), rng.standard_normal(k), rng.standard_normal(k)
        t = (0.05 + 0.9 * rng.random(k)) * (a2 - a1) + a1          # interior zero for the first half
        v2 = (a2 - t) * v1 / (a1 - t)
        v2[k // 2:] = rng.standard_normal(k - k // 2)              # anything for the rest
        v2[-64:] = v1[-64:]                                         # equal values
        a2[-128:-64] = a1[-128:-64] + 1e-3 * rng.standard_normal(64)  # narrow intervals
        ins = [torch.tensor(a.astype(dt), requires_grad=True) for a in (a1, a2, v1, v2)]
        res = interpolate_alpha(*ins)
        go = torch.tensor(rng.standard_normal(k).astype(dt))
        grads = torch.autograd.grad((res * go).sum(), ins)
        out[f"{name}_in"] = np.stack([t_.detach().numpy() for t_ in ins])
        out[f"{name}_out"] = res.detach().numpy()
        out[f"{name}_grad_out"] = go.numpy()
        out[f"{name}_grads"] = np.stack([g.numpy() for g in grads])
    _save("interpolate_alpha", **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    small = {"camera_model": gen_camera_model, "bfgs_update": gen_bfgs_update, "line_search": gen_line_search,
             "analytic_solves": gen_analytic_solves, "angle_ba": gen_angle_ba, "training": gen_training,
             "interpolate_alpha": gen_interpolate_alpha}
    names = args.only or (list(small) + list(SOLVE_CASES))
    for name in names:
        print(name, flush=True)
        if name in small:
            small[name]()
        else:
            gen_solve(name)


if __name__ == "__main__":
    main()
