"""GPU: the CUDA path, called through the Python surface -> C-ABI, against the C oracle on the same
seeded inputs and against the golden fixtures written from the reference."""
import numpy as np
import pytest
import torch

import davo_b200
from conftest import golden_batch, load_golden
from oracle import c_oracle
from parity import assert_within_band, compare_solves, record_parity, reference_band, summary

pytestmark = pytest.mark.gpu

TDT = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


def gpu_solve(batch, weights=None, **kw):
    """Solve a synthetic batch on the GPU; returns the same dict layout as the oracle."""
    dt = TDT[batch.x0.dtype]
    if batch.model == "distort10":
        obj = davo_b200.DistortionObjective(torch.from_numpy(batch.points_3d), torch.from_numpy(batch.obs),
                                            None if batch.pose is None else torch.from_numpy(batch.pose),
                                            weights=weights, dtype=dt)
    elif batch.model == "angle_ba":
        obj = davo_b200.AngleDistanceObjective(torch.from_numpy(batch.obs), torch.from_numpy(batch.weights), dtype=dt)
    else:
        obj = davo_b200.JointPoseObjective(torch.from_numpy(batch.points_3d), torch.from_numpy(batch.obs),
                                           weights=weights, dtype=dt)
    solver = davo_b200.BFGSSolver(**kw).eval()
    info = solver(torch.from_numpy(batch.x0), obj, return_info=True)
    return dict(x=info.parameters.numpy(), cost=info.cost.numpy(), converged=info.converged.numpy(),
                iters=info.iterations.numpy(), fevals=info.evaluations.numpy(), reason=info.reason.numpy())


# ---- objective evaluation ---------------------------------------------------------------------------

@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_stage_matches_oracle(dt):
    b = davo_b200.synthetic.make_distort10(37, 50, seed=3, dtype=dt, random_pose=True)
    obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d), torch.from_numpy(b.obs), torch.from_numpy(b.pose))
    ref = c_oracle.stage(b.points_3d, b.obs, b.pose)
    got = obj.data0.cpu().numpy()
    assert got.shape == ref.shape
    assert np.allclose(got, ref, rtol=3e-6 if dt == np.float32 else 1e-13, atol=1e-6 if dt == np.float32 else 1e-14)
    assert np.array_equal(got[..., 2:], ref[..., 2:])  # observations are copied bit for bit


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("B,N", [(1, 1), (3, 5), (37, 256), (1500, 33)])
def test_stage_identity_pose_is_bit_equal_to_zero_pose(dt, B, N):
    """pose=None takes the flat streaming kernel; it must equal the general kernel run with a zero pose, and the oracle."""
    b = davo_b200.synthetic.make_distort10(B, N, seed=B + N, dtype=dt)
    b.points_3d[0, 0, 2] = 0.0  # the z' == 0 guard (distorted_camera_model.py:57)
    pts, obs = torch.from_numpy(b.points_3d), torch.from_numpy(b.obs)
    flat = davo_b200.DistortionObjective(pts.cuda(), obs.cuda())
    general = davo_b200.DistortionObjective(pts.cuda(), obs.cuda(), torch.zeros(B, 6, dtype=pts.dtype).cuda())
    assert torch.equal(flat.data0, general.data0)
    assert np.array_equal(flat.data0.cpu().numpy(), c_oracle.stage(b.points_3d, b.obs, None))


@pytest.mark.parametrize("dt,rtol", [(np.float32, 2e-4), (np.float64, 1e-11)])
@pytest.mark.parametrize("N", [1, 31, 32, 33, 256, 700])
def test_cost_gradient_matches_oracle(dt, rtol, N):
    b = davo_b200.synthetic.make_distort10(19, N, seed=N, dtype=dt)
    rng = np.random.default_rng(N)
    x = (b.x0 + 0.02 * rng.standard_normal(b.x0.shape)).astype(dt)
    obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d), torch.from_numpy(b.obs), torch.from_numpy(b.pose))
    cost, grad = obj.evaluate(torch.from_numpy(x))
    fo, go = c_oracle.eval_cost_grad("distort10", x, c_oracle.stage(b.points_3d, b.obs, b.pose), N=N)
    assert np.allclose(cost.cpu().numpy(), fo, rtol=rtol)
    scale = np.abs(go).max(axis=1, keepdims=True)
    assert np.all(np.abs(grad.cpu().numpy() - go) <= rtol * scale + 1e-30)


def test_cost_gradient_matches_reference_autograd_golden():
    g = load_golden("camera_model")
    obj = davo_b200.DistortionObjective(torch.from_numpy(g["d10_points"]), torch.from_numpy(g["d10_obs"]),
                                        torch.from_numpy(g["d10_pose"]))
    cost, grad = obj.evaluate(torch.from_numpy(g["d10_x"]))
    assert np.allclose(cost.cpu().numpy(), g["d10_cost"], rtol=1e-11)
    assert np.allclose(grad.cpu().numpy(), g["d10_grad"], rtol=1e-9, atol=1e-12)


def test_weighted_cost_gradient_matches_oracle():
    b = davo_b200.synthetic.make_distort10(11, 70, seed=5, dtype=np.float64)
    rng = np.random.default_rng(5)
    w = rng.uniform(0.0, 2.0, size=(11, 70))
    w[:, ::7] = 0.0  # a visibility mask (networks/calibration_network.py:66)
    x = b.x0 + 0.02 * rng.standard_normal(b.x0.shape)
    obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d), torch.from_numpy(b.obs), weights=torch.from_numpy(w))
    cost, grad = obj.evaluate(torch.from_numpy(x))
    fo, go = c_oracle.eval_cost_grad("distort10", x, c_oracle.stage(b.points_3d, b.obs, b.pose), weights=w, N=70)
    assert np.allclose(cost.cpu().numpy(), fo, rtol=1e-11)
    assert np.allclose(grad.cpu().numpy(), go, rtol=1e-9, atol=1e-12)


def test_descriptor_is_callable_with_reference_convention():
    """(params[k,n], mask[B]) -> err[k] with autograd, so the reference's own solver could drive it."""
    b = davo_b200.synthetic.make_distort10(9, 40, seed=8, dtype=np.float64)
    obj = davo_b200.DistortionObjective(torch.from_numpy(b.points_3d), torch.from_numpy(b.obs))
    mask = torch.tensor([True, False, True, True, False, False, True, False, True])
    x = torch.from_numpy(b.x0)[mask].clone().requires_grad_(True)
    err = obj(x, mask)
    (g,) = torch.autograd.grad(err.sum(), x)
    fo, go = c_oracle.eval_cost_grad("distort10", b.x0, c_oracle.stage(b.points_3d, b.obs, b.pose), N=40)
    assert np.allclose(err.detach().numpy(), fo[mask.numpy()], rtol=1e-11)
    assert np.allclose(g.numpy(), go[mask.numpy()], rtol=1e-9, atol=1e-12)


# ---- camera model, least squares, BFGS helpers ---------------------------------------------------------

def test_forward_model_and_jacobian_match_reference_golden():
    g = load_golden("camera_model")
    pts, th = torch.from_numpy(g["points_3d"]), torch.from_numpy(g["params16"])
    u, v = davo_b200.compute_distorted_camera_model(pts, th)
    assert np.allclose(u.numpy(), g["u"], rtol=1e-12, atol=1e-12)
    assert np.allclose(v.numpy(), g["v"], rtol=1e-12, atol=1e-12)
    J, u2, v2 = davo_b200.compute_distorted_camera_model_and_jacobian(pts[1:], th[1:])
    assert J.shape == (5, 24, 16)
    assert np.allclose(J.numpy(), g["J_autograd"], rtol=1e-9, atol=1e-11)
    assert np.array_equal(u2.numpy(), u.numpy()[1:]) and np.array_equal(v2.numpy(), v.numpy()[1:])
    u32, v32 = davo_b200.compute_distorted_camera_model(pts[1:].float(), th[1:].float())
    assert np.allclose(u32.numpy(), g["u32"][1:], rtol=2e-5, atol=2e-6)
    assert np.allclose(v32.numpy(), g["v32"][1:], rtol=2e-5, atol=2e-6)
    with pytest.raises(ValueError):
        davo_b200.compute_distorted_camera_model(pts, th[:, :10])


@pytest.mark.parametrize("P", [3, 16, 45, 300])
def test_least_squares_matches_oracle(P):
    rng = np.random.default_rng(P)
    B, F, N = 5, 2, 37
    res = rng.standard_normal((B, F, N, 2))
    jac = rng.standard_normal((B, F, N, 2, P))
    w = rng.uniform(0, 1, size=(B, F, N, 1))
    for weights in (None, w):
        e, g = c_oracle.least_squares(res, jac, weights)
        tw = None if weights is None else torch.from_numpy(weights)
        eg = davo_b200.find_error(torch.from_numpy(res), tw)
        gg = davo_b200.find_error_gradient(torch.from_numpy(res), torch.from_numpy(jac), tw)
        assert np.allclose(eg.numpy(), e, rtol=1e-12)
        assert np.allclose(gg.numpy(), g, rtol=1e-10, atol=1e-12)
        # and against the reference's formula written with torch ops (solvers/least_squares_utils.py:16-48)
        sq = torch.from_numpy(res).square() if tw is None else tw * torch.from_numpy(res).square()
        assert np.allclose(eg.numpy(), sq.sum((1, 2, 3)).numpy(), rtol=1e-12)
    assert torch.equal(davo_b200.find_residuals(torch.ones(2, 3), torch.full((2, 3), 0.25)), torch.full((2, 3), 0.75))


@pytest.mark.parametrize("n", [3, 10, 34])
def test_bfgs_update_matches_reference_golden(n):
    g = load_golden("bfgs_update")
    H, s, y = (torch.from_numpy(g[k + str(n)]) for k in ("H", "s", "y"))
    out = davo_b200.BFGSSolver.update_inverse_hessian(H, s, y)
    assert np.allclose(out.numpy(), g[f"Hout{n}"], rtol=1e-12, atol=1e-13)
    assert torch.equal(out[:2], H[:2])  # curvature <= 0: unchanged bit for bit (test_bfgs_solver.py:335-361)
    assert np.array_equal(out.numpy(), c_oracle.bfgs_update(g[f"H{n}"], g[f"s{n}"], g[f"y{n}"]))  # same op order
    sc = davo_b200.BFGSSolver.scale_initial_inverse_hessian(s, y)
    assert sc.shape == (s.shape[0], 1)
    assert np.allclose(sc.squeeze(-1).numpy(), g[f"scale{n}"], rtol=1e-13)
    out32 = davo_b200.BFGSSolver.update_inverse_hessian(H.float(), s.float(), y.float())
    assert np.allclose(out32.numpy()[2:], g[f"Hout{n}_f32"][2:], rtol=2e-3, atol=2e-3)


def test_bfgs_update_known_answer():
    """tests/autograd_solvers/test_bfgs_solver.py:307-332 with its literal numbers."""
    step = torch.tensor([-1.26262069, -0.78272035, 0.98543104], dtype=torch.float64)
    dg = torch.tensor([0.15339519, -0.28944666, 0.54194925], dtype=torch.float64)
    H = torch.tensor([[2.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 3.0]], dtype=torch.float64)
    c = (step * dg).sum()
    left = torch.eye(3) - (step[:, None] * dg[None, :]) / c
    right = torch.eye(3) - (dg[:, None] * step[None, :]) / c
    expected = left @ H @ right + step[:, None] * step[None, :] / c
    result = davo_b200.BFGSSolver.update_inverse_hessian(H, step, dg)
    assert result.shape == (3, 3)
    assert torch.isclose(expected, result).all()


# ---- line search ------------------------------------------------------------------------------------

@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("strong", [False, True])
def test_line_search_reference_cases(dt, strong):
    """test_wolffe_conditions.py:214-305: alpha < 1, alpha > 1, alpha == 0.25 exactly, alpha ~ 0."""
    g = load_golden("line_search")
    name = "float32" if dt == torch.float32 else "float64"
    tg = torch.tensor(g["dist_targets"], dtype=dt)
    d = torch.tensor(g["dist_dirs"], dtype=dt)
    x = torch.zeros(4, 2, dtype=dt)
    obj = davo_b200.AnalyticObjective("distance", (4,), 2, dtype=dt, target=tg)
    f0, gr = obj.evaluate(x)
    alpha, probes = davo_b200.line_search_wolfe_conditions(x, d, f0, gr, obj, strong=strong, return_probes=True)
    assert alpha.shape == (4,)
    assert np.array_equal(alpha.numpy(), g[f"dist_alpha_{name}_{int(strong)}"])
    assert np.array_equal(probes.numpy(), g[f"dist_probes_{name}_{int(strong)}"])
    assert alpha[0] < 1.0 and alpha[1] > 1.0 and alpha[2] == 0.25
    assert torch.isclose(alpha[3], torch.tensor(0.0, dtype=dt))


@pytest.mark.parametrize("name", ["sphere", "log_sphere", "rosenbrock", "cosine", "x2_sine"])
@pytest.mark.parametrize("strong", [False, True])
def test_line_search_analytic_matches_reference_golden(name, strong):
    g = load_golden("line_search")
    x, d = torch.from_numpy(g[f"{name}_x"]), torch.from_numpy(g[f"{name}_d"])
    obj = davo_b200.AnalyticObjective(name, (x.shape[0],), x.shape[1], dtype=torch.float64)
    f0, gr = obj.evaluate(x)
    alpha, probes = davo_b200.line_search_wolfe_conditions(x, d, f0, gr, obj, sufficient_decrease=0.1, curvature=0.6,
                                                           strong=strong, return_probes=True)
    assert np.array_equal(probes.numpy(), g[f"{name}_probes_{int(strong)}"])
    assert np.allclose(alpha.numpy(), g[f"{name}_alpha_{int(strong)}"], rtol=1e-13)
    # the returned alpha satisfies both Wolfe inequalities (test_wolffe_conditions.py:152-211)
    f1, g1 = obj.evaluate(x + alpha[:, None] * d)
    g0 = (d * gr.cpu()).sum(-1)
    ok = alpha > 0
    assert torch.all((f1.cpu() <= f0.cpu() + 0.1 * alpha * g0 + 1e-12)[ok])
    dphi = (d * g1.cpu()).sum(-1)
    if strong:
        assert torch.all((dphi.abs() <= -0.6 * g0 + 1e-9)[ok & (probes < 1000)])


def test_line_search_calibration_matches_reference_golden():
    g = load_golden("line_search")
    obj = davo_b200.DistortionObjective(torch.from_numpy(g["d10_points"]), torch.from_numpy(g["d10_obs"]),
                                        torch.from_numpy(g["d10_pose"]))
    x, d = torch.from_numpy(g["d10_x"]), torch.from_numpy(g["d10_d"])
    f0, gr = obj.evaluate(x)
    alpha, probes = davo_b200.line_search_wolfe_conditions(x, d, f0, gr, obj, strong=True, return_probes=True)
    assert np.array_equal(probes.numpy(), g["d10_probes"])
    assert np.allclose(alpha.numpy(), g["d10_alpha"], rtol=1e-12)


def test_line_search_warns_on_bad_constants():
    obj = davo_b200.AnalyticObjective("sphere", (2,), 3, dtype=torch.float64)
    x = torch.ones(2, 3, dtype=torch.float64)
    f0, gr = obj.evaluate(x)
    with pytest.warns(UserWarning):  # wolfe_conditions.py:65-69
        davo_b200.line_search_wolfe_conditions(x, -gr.cpu(), f0, gr, obj, sufficient_decrease=0.9, curvature=0.1)


# ---- analytic solves: the reference's solver tests, mirrored ------------------------------------------

@pytest.mark.parametrize("name", ["sphere", "sphere_offset", "log_sphere", "rosenbrock", "cosine", "x2_sine"])
def test_analytic_solves_match_reference_golden(name):
    g = load_golden("analytic_solves")
    x0 = torch.from_numpy(g[f"{name}_x0"])
    obj = davo_b200.AnalyticObjective(name, (x0.shape[0],), x0.shape[1], dtype=torch.float64)
    info = davo_b200.BFGSSolver(error_threshold=1e-6).eval()(x0, obj, return_info=True)
    same = info.iterations.numpy() == g[f"{name}_float64_iters"]
    assert same.mean() >= 0.9
    assert np.array_equal(info.evaluations.numpy()[same], g[f"{name}_float64_fevals"][same])
    assert np.array_equal(info.reason.numpy(), g[f"{name}_float64_reason"])
    assert np.allclose(info.parameters.numpy()[same], g[f"{name}_float64_x"][same], rtol=1e-6, atol=1e-8)
    oracle = c_oracle.solve(name, g[f"{name}_x0"], error_threshold=1e-6)
    assert np.array_equal(info.iterations.numpy(), oracle["iters"])
    assert np.allclose(info.parameters.numpy(), oracle["x"], rtol=1e-6, atol=1e-8)  # FMA vs no FMA


def test_solver_reference_behaviours():
    """Batch dimensions (test_bfgs_solver.py:150-160), iteration cap (:134-147), monotone in budget (:190-200)."""
    rng = np.random.default_rng(0)
    x0 = torch.tensor(rng.normal(0.0, 1.0, size=(3, 8, 4)))
    obj = davo_b200.AnalyticObjective("sphere", (3, 8), 4, dtype=torch.float64)
    out = davo_b200.BFGSSolver(error_threshold=1e-6).eval()(x0, obj)
    assert out.shape == (3, 8, 4) and out.dtype == torch.float64 and out.device == x0.device
    assert torch.isclose(out, torch.zeros_like(out), atol=1e-3).all()
    start = torch.tensor([[-1.2, 1.0]], dtype=torch.float64)
    ros = davo_b200.AnalyticObjective("rosenbrock", (1,), 2, dtype=torch.float64)
    errs = []
    for iters in (1, 2, 4, 8, 16, 64):
        info = davo_b200.BFGSSolver(error_threshold=1e-9, iterations=iters).eval()(start, ros, return_info=True)
        assert int(info.iterations[0]) <= iters
        errs.append(float(info.cost[0]))
    assert all(b <= a + 1e-12 for a, b in zip(errs, errs[1:]))
    assert errs[-1] < 1e-6


def test_solver_default_training_mode_and_bad_width():
    obj = davo_b200.AnalyticObjective("sphere", (2,), 3, dtype=torch.float32)
    x0 = torch.ones(2, 3)
    out = davo_b200.BFGSSolver()(x0, obj)  # training mode with drop-path 0.1 (the reference default)
    assert out.shape == x0.shape and not out.requires_grad
    assert davo_b200.BFGSSolver().eval()(x0.clone().requires_grad_(True), obj).requires_grad  # create_graph path
    with pytest.raises(ValueError):
        davo_b200.BFGSSolver().eval()(torch.ones(2, 4), obj)
    out = davo_b200.BFGSSolver(drop_path_p=0.0, training_error_threshold=1e-2)(x0, obj)  # training thresholds
    assert out.shape == (2, 3)


# ---- calibration solves: gates G64 / G32 ---------------------------------------------------------------

DISTORT_CASES = ["solve_cfg2_f64", "solve_cfg2_pose_f64", "solve_cfg4_f64", "solve_cfg2_noisy_f64",
                 "solve_cfg2_f32", "solve_cfg2_f32_thr1e-7", "solve_cfg4_f32"]


@pytest.mark.parametrize("name", DISTORT_CASES)
def test_solve_gate_against_reference(name):
    """Gates G64 / G32.  north_star tolerances — parameters rel <= 1e-4, cost rel <= 1e-5,
    identical accepted-step counts on >= 99 % — against the reference's own outputs (golden fixtures), or the
    reference's self-consistency band where the reference does not meet them against itself (float32 near the
    noise floor, ill-conditioned config 4, noisy data)."""
    g = load_golden(name)
    batch = golden_batch(g["meta"])
    kw = g["meta"]["solver_kwargs"]
    got = gpu_solve(batch, **kw)
    band = reference_band(g, kw["error_threshold"])
    m = compare_solves(got, g, kw["error_threshold"])
    mo = compare_solves(got, c_oracle.solve_batch(batch, **kw), kw["error_threshold"])
    print(name, "kernel vs reference ", summary(m))
    print(name, "reference vs itself ", summary(band))
    print(name, "kernel vs oracle    ", summary(mo))
    record_parity(name, "host inputs (one warp per problem)", kernel_vs_reference=m, reference_vs_itself=band,
                  kernel_vs_oracle=mo)
    assert_within_band(m, band)
    assert_within_band(mo, band)
    if name in ("solve_cfg2_f64", "solve_cfg2_pose_f64"):  # the strict gate, met outright
        assert m["steps_equal"] >= 0.99 and m["reason_equal"] >= 0.99
        assert m["dtheta_p99"] <= 1e-4 and m["dcost_p99"] <= 1e-5


JOINT_CASES = ["solve_cfg3_small_f64", "solve_cfg3_f64", "solve_cfg3_f32"]


@pytest.mark.parametrize("name", JOINT_CASES)
def test_joint_solve_gate_against_reference(name):
    """BASELINE config 3: intrinsics + a 6-DoF pose per view (n = 10 + 6V), warp-per-problem wide solver."""
    g = load_golden(name)
    batch = golden_batch(g["meta"])
    kw = g["meta"]["solver_kwargs"]
    got = gpu_solve(batch, **kw)
    band = reference_band(g, kw["error_threshold"])
    m = compare_solves(got, g, kw["error_threshold"])
    print(name, "kernel vs reference ", summary(m))
    print(name, "reference vs itself ", summary(band))
    record_parity(name, "one CTA per problem", kernel_vs_reference=m, reference_vs_itself=band)
    # the noise-free optimum's cost is ~1e-11 (below the threshold) and is rounding noise of the worse
    # conditioned pose parameters: cost is compared with a floor of 5e-2 of max(cost, threshold) here
    assert_within_band(m, band, cost_floor=5e-2)
    if name == "solve_cfg3_small_f64":
        assert m["steps_equal"] >= 0.99 and m["dtheta_p99"] <= 1e-4


@pytest.mark.parametrize("dt,rtol", [(np.float32, 5e-4), (np.float64, 1e-10)])
@pytest.mark.parametrize("N,V", [(5, 1), (24, 3), (64, 4), (50, 9)])
def test_joint_cost_gradient_matches_oracle(dt, rtol, N, V):
    b = davo_b200.synthetic.make_joint(13, N, V, seed=N + V, dtype=dt)
    rng = np.random.default_rng(N)
    x = (b.x0 + 0.01 * rng.standard_normal(b.x0.shape)).astype(dt)
    obj = davo_b200.JointPoseObjective(torch.from_numpy(b.points_3d), torch.from_numpy(b.obs))
    cost, grad = obj.evaluate(torch.from_numpy(x))
    fo, go = c_oracle.eval_cost_grad("joint", x, b.points_3d, b.obs, N=N, V=V)
    assert np.allclose(cost.cpu().numpy(), fo, rtol=rtol)
    scale = np.abs(go).max(axis=1, keepdims=True)
    assert np.all(np.abs(grad.cpu().numpy() - go) <= rtol * scale + 1e-30)


def test_joint_cost_gradient_matches_reference_autograd_golden():
    g = load_golden("camera_model")
    obj = davo_b200.JointPoseObjective(torch.from_numpy(g["joint_points"]), torch.from_numpy(g["joint_obs"]))
    cost, grad = obj.evaluate(torch.from_numpy(g["joint_x"]))
    assert np.allclose(cost.cpu().numpy(), g["joint_cost"], rtol=1e-11)
    assert np.allclose(grad.cpu().numpy(), g["joint_grad"], rtol=1e-9, atol=1e-11)


def test_joint_weighted_and_line_search_match_oracle():
    b = davo_b200.synthetic.make_joint(12, 36, 2, seed=9, dtype=np.float64)
    rng = np.random.default_rng(9)
    w = rng.uniform(0.0, 1.5, size=(12, 2, 36))
    obj = davo_b200.JointPoseObjective(torch.from_numpy(b.points_3d), torch.from_numpy(b.obs), weights=torch.from_numpy(w))
    x = torch.from_numpy(b.x0)
    cost, grad = obj.evaluate(x)
    fo, go = c_oracle.eval_cost_grad("joint", b.x0, b.points_3d, b.obs, weights=w, N=36, V=2)
    assert np.allclose(cost.cpu().numpy(), fo, rtol=1e-11)
    assert np.allclose(grad.cpu().numpy(), go, rtol=1e-9, atol=1e-11)
    d = -1e-3 * go
    alpha, probes = davo_b200.line_search_wolfe_conditions(x, torch.from_numpy(d), cost, grad, obj, strong=True,
                                                           return_probes=True)
    ao, po = c_oracle.line_search("joint", b.x0, d, fo, go, b.points_3d, b.obs, w, N=36, V=2, strong=True)
    assert np.array_equal(probes.numpy(), po)
    assert np.allclose(alpha.numpy(), ao, rtol=1e-12)


def test_ill_conditioned_problems_do_not_stall_in_float32():
    """Regression: a BFGS update that replaced y^T H by (H y)^T ("H is symmetric") let the rounding asymmetry of
    H grow in float32 until these config-4 problems, which the reference algorithm solves in ~170-280 steps,
    spent all 1000 iterations with ~39 probes per line search.  Trajectories of single float32 problems are
    chaotic here (the reference agrees with itself on 43 % of step counts), so the check is on population
    statistics against the oracle; measured on the first 2048 problems: reference 143.5 steps / 3.56 % capped,
    oracle 145.0 / 3.56 %, this kernel 147.3 / 3.86 %."""
    batch = davo_b200.synthetic.make_distort10(65536, 256, seed=0xB200, dtype=np.float32, ill_conditioned=True,
                                               pathological=0.02)
    kw = dict(error_threshold=1e-5, iterations=1000)
    pick = [60403, 48468] + list(range(0, 4094))
    sub = davo_b200.synthetic.CalibrationBatch(batch.model, batch.points_3d[pick], batch.obs[pick], batch.pose[pick],
                                               batch.x0[pick], batch.truth[pick], 1)
    got = gpu_solve(sub, **kw)
    ref = c_oracle.solve_batch(sub, **kw)
    assert got["iters"][0] < 400 and got["iters"][1] < 400, got["iters"][:2]
    assert got["reason"][0] == 0 and got["reason"][1] == 0
    capped_gpu, capped_ref = (got["reason"] == 2).mean(), (ref["reason"] == 2).mean()
    assert capped_gpu <= capped_ref + 0.01, (capped_gpu, capped_ref)
    assert abs(got["iters"].mean() - ref["iters"].mean()) <= 0.05 * ref["iters"].mean()
    ev_gpu, ev_ref = (got["fevals"] - got["iters"]).astype(np.int64), (ref["fevals"] - ref["iters"]).astype(np.int64)
    assert abs(np.median(ev_gpu) - np.median(ev_ref)) <= 0.05 * np.median(ev_ref)
    assert np.percentile(ev_gpu, 99) <= 1.25 * np.percentile(ev_ref, 99), (np.percentile(ev_gpu, 99), np.percentile(ev_ref, 99))
    # a line search that bisects down to lo == hi on every one of 1000 iterations (~40 probes each) does occur in
    # the reference algorithm too (oracle: 1 of these 65536 problems, 25 799 probes); it must stay that rare
    assert (ev_gpu > 5000).sum() <= 3, np.sort(ev_gpu)[-6:]


@pytest.mark.parametrize("N", [1, 7, 33, 100])
def test_solve_ragged_match_counts(N):
    """N not a multiple of the warp width, down to a single match (under-determined: must not hang)."""
    batch = davo_b200.synthetic.make_distort10(40, N, seed=100 + N, dtype=np.float64)
    kw = dict(error_threshold=1e-12, iterations=60)
    got = gpu_solve(batch, **kw)
    ref = c_oracle.solve_batch(batch, **kw)
    m = compare_solves(got, ref, 1e-12)
    print(N, summary(m))
    assert m["steps_equal"] >= 0.9
    assert m["reason_equal"] >= 0.9


def test_solve_weighted_matches_oracle():
    batch = davo_b200.synthetic.make_distort10(64, 96, seed=77, dtype=np.float64)
    rng = np.random.default_rng(77)
    w = (rng.uniform(size=(64, 96)) > 0.2).astype(np.float64)
    kw = dict(error_threshold=1e-12, iterations=200)
    got = gpu_solve(batch, weights=torch.from_numpy(w), **kw)
    staged = c_oracle.stage(batch.points_3d, batch.obs, batch.pose)
    ref = c_oracle.solve("distort10", batch.x0, staged, weights=w, N=96, **kw)
    m = compare_solves(got, ref, 1e-12)
    print(summary(m))
    assert m["steps_equal"] >= 0.98 and m["dtheta_p99"] <= 1e-6


def test_streamed_host_path_equals_resident_path():
    """Host inputs go through the chunked copy/stage/solve pipeline; results must be bit-identical to staging
    everything first, for any chunk size (including one that does not divide the batch)."""
    batch = davo_b200.synthetic.make_distort10(1000, 64, seed=4, dtype=np.float32)
    rng = np.random.default_rng(4)
    w = torch.from_numpy((rng.uniform(size=(1000, 64)) > 0.1).astype(np.float32))
    pts, obs, x0 = (torch.from_numpy(a) for a in (batch.points_3d, batch.obs, batch.x0))
    for weights in (None, w):
        resident = davo_b200.DistortionObjective(pts.cuda(), obs.cuda(), weights=None if weights is None else weights.cuda())
        assert resident.is_staged
        solver = davo_b200.BFGSSolver(error_threshold=1e-6).eval()
        ref = solver(x0, resident, return_info=True)
        for chunk in (16384, 333, 64):
            lazy = davo_b200.DistortionObjective(pts.pin_memory(), obs.pin_memory(), weights=weights)
            assert not lazy.is_staged
            solver.stream_chunk = chunk
            got = solver(x0, lazy, return_info=True)
            assert lazy.is_staged and torch.equal(lazy.data0, resident.data0)
            for a, b in zip(got, ref):
                assert torch.equal(a, b)


def test_solve_full_size_properties():
    """BASELINE config 2 at full size (64K x 256, float32): size-independent properties.
    (1) solving is idempotent: a converged solution re-submitted retires at once with 0 steps;
    (2) permuting the problems permutes the outputs; (3) every output row is written.
    The batch is device resident, so every solve below takes the same route (two problems per warp + the
    straggler launch); a 4096-problem batch is solved one warp per problem, whose sums associate differently:
    it must agree within the float32 band, not bit for bit."""
    B = 65536
    batch = davo_b200.synthetic.make_distort10(B, 256, seed=0xB200, dtype=np.float32)
    obj = davo_b200.DistortionObjective(torch.from_numpy(batch.points_3d).cuda(), torch.from_numpy(batch.obs).cuda())
    solver = davo_b200.BFGSSolver(error_threshold=1e-5).eval()
    x0 = torch.from_numpy(batch.x0).cuda()
    info = solver(x0, obj, return_info=True)
    assert info.converged.float().mean() > 0.99
    assert int(info.iterations.min()) >= 1 and int(info.iterations.max()) <= 1000
    again = solver(info.parameters, obj, return_info=True)
    conv = info.converged
    assert torch.equal(again.parameters[conv], info.parameters[conv])
    assert int(again.iterations[conv].max()) == 0
    assert torch.equal(again.cost[conv], info.cost[conv])
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).cuda()
    sub = davo_b200.DistortionObjective.from_staged(obj.data0[perm])
    pinfo = solver(x0[perm], sub, return_info=True)
    assert torch.equal(pinfo.parameters, info.parameters[perm])
    assert torch.equal(pinfo.iterations, info.iterations[perm])
    assert torch.equal(pinfo.cost, info.cost[perm])
    small = davo_b200.DistortionObjective.from_staged(obj.data0[:4096])
    sinfo = solver(x0[:4096], small, return_info=True)
    same = (sinfo.iterations == info.iterations[:4096]).float().mean()
    assert same >= 0.97, same  # the reference agrees with itself on 99.7 % at this threshold
    rel = ((sinfo.parameters - info.parameters[:4096]).abs() / info.parameters[:4096].abs().clamp(min=1.0)).max(dim=1).values
    assert float(rel.median()) <= 1e-4


@pytest.mark.parametrize("dt,N", [(np.float64, 50), (np.float64, 33), (np.float32, 50), (np.float64, 96)])
def test_two_per_warp_kernel_ragged_and_generic_match_counts(dt, N):
    """Batches large enough for the two-problems-per-warp kernel with N not a multiple of 32 (a lane's pair may
    lack its second match) and with a run-time N (only N = 256 has a compile-time instantiation), against the oracle."""
    B = 16384
    batch = davo_b200.synthetic.make_distort10(B, N, seed=300 + N, dtype=dt)
    kw = dict(error_threshold=1e-12 if dt == np.float64 else 1e-5, iterations=80)
    obj = davo_b200.DistortionObjective(torch.from_numpy(batch.points_3d).cuda(), torch.from_numpy(batch.obs).cuda())
    info = davo_b200.BFGSSolver(**kw).eval()(torch.from_numpy(batch.x0).cuda(), obj, return_info=True)
    got = dict(x=info.parameters.cpu().numpy(), cost=info.cost.cpu().numpy(), iters=info.iterations.cpu().numpy(),
               fevals=info.evaluations.cpu().numpy(), reason=info.reason.cpu().numpy())
    ref = c_oracle.solve_batch(batch, **kw)
    m = compare_solves(got, ref, kw["error_threshold"])
    print(dt.__name__, N, summary(m))
    if dt == np.float64:
        assert m["steps_equal"] >= 0.99 and m["reason_equal"] >= 0.99 and m["dtheta_p99"] <= 1e-6
    else:
        assert m["steps_equal"] >= 0.9 and m["reason_equal"] >= 0.97 and m["dtheta_median"] <= 1e-4


@pytest.mark.parametrize("name", ["solve_cfg2_f64", "solve_cfg2_f32", "solve_cfg4_f32"])
def test_two_per_warp_kernel_matches_reference_golden(name):
    """The golden problems of configs 2 and 4 tiled to >= 15K problems (the two-problems-per-warp route + the
    straggler launch): every copy of a problem gives bit-identical outputs whatever half-warp, partner problem and queue
    position it meets, and copy 0 is held to the same gate against the REFERENCE's results as the small batch."""
    g = load_golden(name)
    batch = golden_batch(g["meta"])
    kw = g["meta"]["solver_kwargs"]
    B0 = batch.B
    reps = -(-15360 // B0)
    tile = lambda a: np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1)))
    big = davo_b200.synthetic.CalibrationBatch(batch.model, tile(batch.points_3d), tile(batch.obs), tile(batch.pose),
                                               tile(batch.x0), tile(batch.truth), 1)
    dt = TDT[big.x0.dtype]
    obj = davo_b200.DistortionObjective(torch.from_numpy(big.points_3d).cuda(), torch.from_numpy(big.obs).cuda(),
                                        torch.from_numpy(big.pose).cuda(), dtype=dt)
    info = davo_b200.BFGSSolver(**kw).eval()(torch.from_numpy(big.x0).cuda(), obj, return_info=True)
    x = info.parameters.cpu().numpy().reshape(reps, B0, -1)
    it = info.iterations.cpu().numpy().reshape(reps, B0)
    co = info.cost.cpu().numpy().reshape(reps, B0)
    for r in range(1, reps):
        assert np.array_equal(x[r], x[0], equal_nan=True) and np.array_equal(it[r], it[0])
        assert np.array_equal(co[r], co[0], equal_nan=True)
    got = dict(x=x[0], cost=co[0], iters=it[0], fevals=info.evaluations.cpu().numpy()[:B0],
               reason=info.reason.cpu().numpy()[:B0])
    m = compare_solves(got, g, kw["error_threshold"])
    band = reference_band(g, kw["error_threshold"])
    print(name, "two-per-warp kernel vs reference", summary(m))
    record_parity(name, "two problems per warp (tiled to >= 15K problems)", kernel_vs_reference=m,
                  reference_vs_itself=band)
    assert_within_band(m, band)
    if name == "solve_cfg2_f64":
        assert m["steps_equal"] >= 0.995 and m["dtheta_max"] <= 1e-6


@pytest.mark.parametrize("thr", [1e-5, 1e-7])
def test_full_size_batch_slice_against_oracle(thr):
    """BASELINE config 2 at full size (64K x 256, float32, the bench's own inputs) on the route the bench takes
    (two problems per warp + straggler launch): rows 0..4095 of the 64K-problem launch against the C oracle on
    the same 4096 problems, held to the reference's own float32 self-consistency band (fixture solve_cfg2_f32 for
    threshold 1e-5; solve_cfg2_f32_thr1e-7 for the bench's 1e-7, where the reference is chaotic against itself)."""
    B, S = 65536, 4096
    batch = davo_b200.synthetic.make_distort10(B, 256, seed=0xB200, dtype=np.float32)
    kw = dict(error_threshold=thr, iterations=1000)
    obj = davo_b200.DistortionObjective(torch.from_numpy(batch.points_3d).cuda(), torch.from_numpy(batch.obs).cuda())
    info = davo_b200.BFGSSolver(**kw).eval()(torch.from_numpy(batch.x0).cuda(), obj, return_info=True)
    got = dict(x=info.parameters[:S].cpu().numpy(), cost=info.cost[:S].cpu().numpy(),
               iters=info.iterations[:S].cpu().numpy(), fevals=info.evaluations[:S].cpu().numpy(),
               reason=info.reason[:S].cpu().numpy())
    ref = c_oracle.solve_batch(batch.slice(0, S), **kw)
    m = compare_solves(got, ref, thr)
    g = load_golden("solve_cfg2_f32" if thr == 1e-5 else "solve_cfg2_f32_thr1e-7")
    band = reference_band(g, thr)
    print(f"thr {thr}: 64K launch rows 0..4095 vs oracle", summary(m))
    print(f"thr {thr}: reference vs itself            ", summary(band))
    record_parity(f"cfg2_64K_slice_f32_thr{thr:g}", "two problems per warp (64K-problem launch, rows 0..4095)",
                  kernel_vs_oracle=m, reference_vs_itself=band)
    assert_within_band(m, band)


@pytest.mark.parametrize("B", [14209, 20001, 65535])
def test_two_per_warp_kernel_odd_batch_sizes(B):
    """Batch sizes that leave a half-warp without a partner problem at the end of the queue: every row is written
    exactly once (outputs pre-filled with a sentinel), and the first rows agree with a small batch of the same
    problems (solved one warp per problem) within the float32 band."""
    batch = davo_b200.synthetic.make_distort10(B, 64, seed=77, dtype=np.float32)
    obj = davo_b200.DistortionObjective(torch.from_numpy(batch.points_3d).cuda(), torch.from_numpy(batch.obs).cuda())
    solver = davo_b200.BFGSSolver(error_threshold=1e-5).eval()
    from davo_b200.solvers import SolveBuffers
    buf = SolveBuffers.allocate(B, 10, torch.float32, obj.device)
    buf.x.fill_(float("nan")); buf.cost.fill_(float("nan")); buf.iterations.fill_(-7); buf.reason.fill_(-7)
    x0 = torch.from_numpy(batch.x0).cuda()
    solver.solve_into(x0, obj, out=buf)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(buf.x).all()) and bool(torch.isfinite(buf.cost).all())
    assert int(buf.iterations.min()) >= 1 and int(buf.reason.min()) >= 0 and int(buf.reason.max()) <= 3
    assert float(buf.converged.float().mean()) > 0.99
    small = davo_b200.DistortionObjective.from_staged(obj.data0[:2048])
    sinfo = solver(x0[:2048], small, return_info=True)
    assert float((sinfo.iterations == buf.iterations[:2048]).float().mean()) >= 0.97


def test_stragglers_are_handed_off_and_solved():
    """Config 4 (ill-conditioned) at full size: problems that pass the evaluation cap of the two-per-warp launch are
    re-solved by the second launch; every row is written, no hand-off flag survives, and the population agrees with
    the one-warp-per-problem kernel (which solves the same problems in 4K chunks through the streamed path)."""
    B = 65536
    batch = davo_b200.synthetic.make_distort10(B, 256, seed=0xB200, dtype=np.float32, ill_conditioned=True,
                                               pathological=0.02)
    kw = dict(error_threshold=1e-5, iterations=1000)
    pts, obs, x0 = (torch.from_numpy(a) for a in (batch.points_3d, batch.obs, batch.x0))
    dev = davo_b200.BFGSSolver(**kw).eval()(x0.cuda(), davo_b200.DistortionObjective(pts.cuda(), obs.cuda()),
                                             return_info=True)
    reason = dev.reason.cpu().numpy()
    assert reason.min() >= 0 and reason.max() <= 3          # no internal hand-off value escapes
    evals = (dev.evaluations - dev.iterations).cpu().numpy()
    assert (dev.evaluations.cpu().numpy() > 4096).sum() >= 4  # the stragglers exist and were solved to the end
    host = davo_b200.BFGSSolver(**kw).eval()(x0, davo_b200.DistortionObjective(pts, obs), return_info=True)
    hr = host.reason.numpy()
    assert np.abs(np.bincount(reason, minlength=4) - np.bincount(hr, minlength=4)).max() <= 0.01 * B
    assert abs(evals.mean() - (host.evaluations - host.iterations).numpy().mean()) <= 0.03 * evals.mean()
    assert abs(dev.iterations.float().mean().item() - host.iterations.float().mean().item()) <= 0.03 * dev.iterations.float().mean().item()


def test_forward_model_is_differentiable_in_the_parameters():
    """compute_distorted_camera_model(points, theta) back-propagates to theta through the analytic Jacobian; the
    fixture holds autograd's Jacobian of the reference forward model."""
    g = load_golden("camera_model")
    pts = torch.from_numpy(g["points_3d"][1:])
    th = torch.from_numpy(g["params16"][1:]).clone().requires_grad_(True)
    u, v = davo_b200.compute_distorted_camera_model(pts, th)
    rng = np.random.default_rng(3)
    a, b = torch.from_numpy(rng.standard_normal(u.shape)), torch.from_numpy(rng.standard_normal(v.shape))
    ((u * a).sum() + (v * b).sum()).backward()
    expect = np.einsum("bi,bij->bj", np.concatenate([a.numpy(), b.numpy()], axis=1), g["J_autograd"])
    assert np.allclose(th.grad.numpy(), expect, rtol=1e-9, atol=1e-10)


# ---- the entry script's bundle-adjustment objective ---------------------------------

BA_SHAPES = {"a": (4, 8), "b": (2, 5), "c": (3, 11), "d": (6, 7), "e": (3, 20), "f": (4, 30)}  # e, f: n = 75, 111 > 64


@pytest.mark.parametrize("tag", sorted(BA_SHAPES))
def test_angle_ba_cost_gradient_match_reference_autograd_golden(tag):
    """networks/calibration_network.py:58-67 evaluated by the reference's own functions + autograd (fixture) vs the
    CUDA evaluator; rows 1-4 hit the Taylor branches, elu's negative branch and a zero rotation."""
    g = load_golden("angle_ba")
    x, obs, vis = g[f"{tag}_x"], g[f"{tag}_obs"], g[f"{tag}_vis"]
    obj = davo_b200.AngleDistanceObjective(torch.from_numpy(obs), torch.from_numpy(vis))
    cost, grad = obj.evaluate(torch.from_numpy(x))
    assert np.allclose(cost.cpu().numpy(), g[f"{tag}_cost"], rtol=1e-12)
    scale = np.abs(g[f"{tag}_grad"]).max(axis=1, keepdims=True)
    assert np.all(np.abs(grad.cpu().numpy() - g[f"{tag}_grad"]) <= 1e-11 * scale)
    obj32 = davo_b200.AngleDistanceObjective(torch.from_numpy(obs).float(), torch.from_numpy(vis).float())
    cost32, grad32 = obj32.evaluate(torch.from_numpy(x).float())
    assert np.allclose(cost32.cpu().numpy(), g[f"{tag}_cost32"], rtol=5e-6)
    # float32: the reference's own gradient is only good to ~3e-7 of its largest entry (vs float64); allow 1e-4
    assert np.all(np.abs(grad32.cpu().numpy() - g[f"{tag}_grad"]) <= 1e-4 * scale)


@pytest.mark.parametrize("tag", sorted(BA_SHAPES))
def test_angle_ba_line_search_matches_reference_golden(tag):
    g = load_golden("angle_ba")
    x, obs, vis = g[f"{tag}_x"], g[f"{tag}_obs"], g[f"{tag}_vis"]
    obj = davo_b200.AngleDistanceObjective(torch.from_numpy(obs), torch.from_numpy(vis))
    alpha, probes = davo_b200.line_search_wolfe_conditions(
        torch.from_numpy(x), torch.from_numpy(g[f"{tag}_d"]), torch.from_numpy(g[f"{tag}_cost"]),
        torch.from_numpy(g[f"{tag}_grad"]), obj, strong=True, return_probes=True)
    assert np.array_equal(probes.numpy(), g[f"{tag}_probes"])
    assert np.allclose(alpha.numpy(), g[f"{tag}_alpha"], rtol=1e-10)


def test_angle_ba_all_visible_default_and_bool_mask():
    """visibility_mask=None means every point is visible; a bool mask is accepted like the driver passes it."""
    b = davo_b200.synthetic.make_angle_ba(10, 6, 3, seed=9, dtype=np.float64)
    x = torch.from_numpy(b.x0)
    ones = np.ones_like(b.weights)
    c0, g0 = davo_b200.AngleDistanceObjective(torch.from_numpy(b.obs)).evaluate(x)
    fo, go = c_oracle.eval_cost_grad("angle_ba", b.x0, b.obs, None, ones, N=6, V=3)
    assert np.allclose(c0.cpu().numpy(), fo, rtol=1e-12)
    assert np.allclose(g0.cpu().numpy(), go, rtol=1e-9, atol=1e-12)
    c1, g1 = davo_b200.AngleDistanceObjective(torch.from_numpy(b.obs), torch.from_numpy(b.weights > 0)).evaluate(x)
    fo, go = c_oracle.eval_cost_grad("angle_ba", b.x0, b.obs, None, b.weights, N=6, V=3)
    assert np.allclose(c1.cpu().numpy(), fo, rtol=1e-12)
    assert np.allclose(g1.cpu().numpy(), go, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", ["solve_ba_f64_30steps", "solve_ba_small_f64", "solve_ba_f32_30steps", "solve_ba_f64",
                                  "solve_ba_f32", "solve_ba_n75_f64", "solve_ba_n111_f64"])
def test_angle_ba_solve_matches_reference_golden(name):
    """Capped at 30-40 accepted steps in float64: step for step the reference's trajectory.  Full-length runs of
    this non-smooth objective are chaotic in the reference itself: population-level gate against its own band."""
    g = load_golden(name)
    batch = golden_batch(g["meta"])
    kw = g["meta"]["solver_kwargs"]
    got = gpu_solve(batch, **kw)
    m = compare_solves(got, g, kw["error_threshold"])
    band = reference_band(g, kw["error_threshold"])
    print(name, "kernel vs reference", summary(m))
    print(name, "reference vs itself", summary(band))
    record_parity(name, "one warp per problem (wide solver)", kernel_vs_reference=m, reference_vs_itself=band)
    assert_within_band(m, band)
    if name in ("solve_ba_f64_30steps", "solve_ba_small_f64", "solve_ba_n75_f64", "solve_ba_n111_f64"):
        assert m["steps_equal"] == 1.0 and m["fevals_equal"] >= 0.95 and m["reason_equal"] == 1.0
        assert m["dtheta_max"] <= (1e-4 if "small" in name else 1e-6)  # north_star: 1e-4 (flat directions at V=2, N=5)
    ref = c_oracle.solve_batch(batch, **kw)
    mo = compare_solves(got, ref, kw["error_threshold"])
    print(name, "kernel vs oracle", summary(mo))
    if name.endswith("steps") or "small" in name or "_n" in name:
        assert mo["steps_equal"] >= 0.98


def test_angle_ba_large_batch_properties():
    """16K bundle-adjustment problems (4 views x 8 points, the entry script's shape, float32): every row is written,
    permuting the problems permutes the outputs bit for bit, the error never increases, and the float32 population
    looks like the float64 one."""
    B = 16384
    b = davo_b200.synthetic.make_angle_ba(B, 8, 4, seed=0xB207, dtype=np.float32)
    obs, vis, x0 = (torch.from_numpy(a).cuda() for a in (b.obs, b.weights, b.x0))
    solver = davo_b200.BFGSSolver(error_threshold=1e-4, iterations=200).eval()
    obj = davo_b200.AngleDistanceObjective(obs, vis)
    info = solver(x0, obj, return_info=True)
    start, _ = obj.evaluate(x0, want_grad=False)
    assert bool(torch.isfinite(info.parameters).all()) and int(info.iterations.min()) >= 1
    assert bool((info.cost <= start * (1 + 1e-5)).all())
    assert float(info.cost.median()) < 0.05 * float(start.median())
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3)).cuda()
    pinfo = solver(x0[perm], davo_b200.AngleDistanceObjective(obs[perm], vis[perm]), return_info=True)
    assert torch.equal(pinfo.parameters, info.parameters[perm])
    assert torch.equal(pinfo.iterations, info.iterations[perm]) and torch.equal(pinfo.cost, info.cost[perm])
    sub = slice(0, 2048)
    i64 = solver(x0[sub].double(), davo_b200.AngleDistanceObjective(obs[sub].double(), vis[sub].double()), return_info=True)
    assert abs(float(i64.cost.median()) - float(info.cost[sub].median())) <= 0.2 * float(i64.cost.median())


def test_angle_ba_rejects_bad_shapes():
    obs = torch.zeros(3, 4, 8, 2)
    obj = davo_b200.AngleDistanceObjective(obs)
    with pytest.raises(ValueError):
        davo_b200.BFGSSolver().eval()(torch.zeros(3, 44), obj)  # calibration_pinhole_camera_model.py:51-56
    with pytest.raises(ValueError):
        davo_b200.AngleDistanceObjective(torch.zeros(3, 1, 8, 2))  # a single view has no relative pose
    with pytest.raises(NotImplementedError):
        big = davo_b200.AngleDistanceObjective(torch.zeros(2, 4, 40, 2))  # n = 141 > 128
        davo_b200.BFGSSolver().eval()(torch.zeros(2, 141), big)


def test_calibration_network_forward_refines_its_initial_guess():
    """CalibrationNetwork.forward (networks/calibration_network.py:54-73): MLP guess -> BFGS on the bundle-adjustment
    objective; equals calling the solver on that guess directly, and lowers the error."""
    torch.manual_seed(0)
    b = davo_b200.synthetic.make_angle_ba(48, 8, 4, seed=12, dtype=np.float32)
    net = davo_b200.CalibrationNetwork(4, 8).cuda().eval()
    obs, vis = torch.from_numpy(b.obs).cuda(), torch.from_numpy(b.weights).cuda() > 0
    with torch.no_grad():
        guess = net.estimate(obs.reshape(-1, 64))  # the fused tcgen05 kernel (tests/test_gpu_secant.py pins it to the modules)
        params, err = net(obs, vis, return_error=True)
    assert params.shape == (48, 45) and err.shape == (48,) and params.device.type == "cuda"
    obj = davo_b200.AngleDistanceObjective(obs, vis)
    start, _ = obj.evaluate(guess, want_grad=False)
    direct = davo_b200.BFGSSolver(error_threshold=1e-7).eval()(guess, obj, return_info=True)
    assert torch.equal(direct.parameters, params) and torch.equal(direct.cost, err)
    assert bool((err <= start).all()) and float(err.median()) < 0.5 * float(start.median())
    net.train()
    assert net(obs, vis).requires_grad  # training mode differentiates through the solve (tests/test_gpu_training.py)


def test_randomised_sweep_against_oracle():
    """tools/fuzz_parity.py for 20 s: random batch sizes on both sides of the two-per-warp threshold, ragged and tiny
    match counts, weights, both precisions, all four objectives, iteration caps 0..200 — no case may leave the band
    (float64: >= 90 % identical step counts and median |dtheta| <= 1e-6 per batch)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "20", "7"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "SUSPECT" not in r.stdout, r.stdout[-3000:]
    assert " runs in " in r.stdout


# ---- regression tests for round-1 review findings ---------------------------------------------------------------

def test_lazily_staged_weighted_objective_keeps_its_weights_in_line_search_and_solve():
    """A host-input (lazily staged) DistortionObjective with weights: the descriptor must say has_weights before
    the weights are on the device (line search and solve_into build it before touching data0)."""
    b = davo_b200.synthetic.make_distort10(48, 64, seed=91, dtype=np.float64)
    rng = np.random.default_rng(91)
    w = rng.uniform(0.0, 2.0, size=(48, 64))
    w[:, ::5] = 0.0
    pts, obs, x0 = (torch.from_numpy(a) for a in (b.points_3d, b.obs, b.x0))
    resident = davo_b200.DistortionObjective(pts.cuda(), obs.cuda(), weights=torch.from_numpy(w).cuda())
    cost, grad = resident.evaluate(x0)
    d = -1e-2 * grad
    want_alpha, want_probes = davo_b200.line_search_wolfe_conditions(x0, d, cost, grad, resident, strong=True,
                                                                     return_probes=True)
    lazy = davo_b200.DistortionObjective(pts, obs, weights=torch.from_numpy(w))
    assert not lazy.is_staged and lazy.has_weights()
    alpha, probes = davo_b200.line_search_wolfe_conditions(x0, d, cost, grad, lazy, strong=True, return_probes=True)
    assert torch.equal(alpha, want_alpha) and torch.equal(probes, want_probes)
    unweighted = davo_b200.DistortionObjective(pts.cuda(), obs.cuda())
    a0, _ = davo_b200.line_search_wolfe_conditions(x0, d, cost, grad, unweighted, strong=True, return_probes=True)
    assert not torch.equal(a0, want_alpha)  # the weights matter for this input: a dropped flag would show
    lazy2 = davo_b200.DistortionObjective(pts, obs, weights=torch.from_numpy(w))
    solver = davo_b200.BFGSSolver(error_threshold=1e-12, iterations=50).eval()
    got = solver.solve_into(x0.cuda(), lazy2)
    ref = solver.solve_into(x0.cuda(), resident)
    assert torch.equal(got.x, ref.x) and torch.equal(got.iterations, ref.iterations)


def test_least_squares_wrappers_are_differentiable_like_the_reference():
    """find_error / find_error_gradient are plain differentiable torch ops in the reference
    (solvers/least_squares_utils.py:16-48); the kernel-backed versions carry the same gradients."""
    rng = np.random.default_rng(3)
    B, F, N, P = 4, 2, 9, 5
    r = torch.tensor(rng.standard_normal((B, F, N, 2)), requires_grad=True)
    J = torch.tensor(rng.standard_normal((B, F, N, 2, P)), requires_grad=True)
    w = torch.tensor(rng.uniform(0.1, 1.0, size=(B, F, N, 1)), requires_grad=True)
    up = torch.tensor(rng.standard_normal(B))
    upg = torch.tensor(rng.standard_normal((B, P)))
    for weights in (None, w):
        ref_e = (r.square() if weights is None else r.square() * weights).sum(dim=(-3, -2, -1))
        ref_g = (2.0 * (r if weights is None else r * weights).unsqueeze(-1) * J).sum(dim=(-4, -3, -2))
        inputs = [r, J] + ([] if weights is None else [w])
        want_e = torch.autograd.grad((ref_e * up).sum(), [r] + inputs[2:])
        want_g = torch.autograd.grad((ref_g * upg).sum(), inputs)
        e = davo_b200.find_error(r, weights)
        g = davo_b200.find_error_gradient(r, J, weights)
        assert torch.allclose(e, ref_e.detach(), rtol=1e-12) and torch.allclose(g, ref_g.detach(), rtol=1e-11, atol=1e-12)
        got_e = torch.autograd.grad((e * up).sum(), [r] + inputs[2:])
        got_g = torch.autograd.grad((g * upg).sum(), inputs)
        for a, b_ in zip(got_e + got_g, want_e + want_g):
            assert torch.allclose(a, b_, rtol=1e-11, atol=1e-12)
    with pytest.raises(NotImplementedError):
        davo_b200.compute_distorted_camera_model(torch.zeros(1, 2, 3, requires_grad=True), torch.zeros(1, 16))
