"""CPU: the C oracle (oracle/calib_oracle.c) against fixtures written by the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins the oracle; the GPU tests then compare CUDA to it."""
import numpy as np
import pytest

from conftest import golden_batch, load_golden
from oracle import c_oracle
from parity import assert_within_band, compare_solves, reference_band, summary


def test_forward_model_matches_reference():
    g = load_golden("camera_model")
    u, v = c_oracle.project(g["points_3d"], g["params16"])
    # row 0 hits the z' == 0 guard (distorted_camera_model.py:57): a = x/1e-8 is huge there
    assert np.allclose(u, g["u"], rtol=1e-12, atol=1e-12)
    assert np.allclose(v, g["v"], rtol=1e-12, atol=1e-12)
    u32, v32 = c_oracle.project(g["points_3d"].astype(np.float32), g["params16"].astype(np.float32))
    assert np.allclose(u32[1:], g["u32"][1:], rtol=2e-5, atol=2e-6)
    assert np.allclose(v32[1:], g["v32"][1:], rtol=2e-5, atol=2e-6)


def test_jacobian_matches_autograd_of_reference_forward():
    g = load_golden("camera_model")
    J, _, _ = c_oracle.project(g["points_3d"][1:], g["params16"][1:], jacobian=True)
    assert J.shape == g["J_autograd"].shape
    assert np.allclose(J, g["J_autograd"], rtol=1e-10, atol=1e-11)


def test_cost_and_gradient_match_autograd():
    g = load_golden("camera_model")
    staged = c_oracle.stage(g["d10_points"], g["d10_obs"], g["d10_pose"])
    f, gr = c_oracle.eval_cost_grad("distort10", g["d10_x"], staged, N=staged.shape[1])
    assert np.allclose(f, g["d10_cost"], rtol=1e-12)
    assert np.allclose(gr, g["d10_grad"], rtol=1e-10, atol=1e-12)
    V, N = g["joint_obs"].shape[1], g["joint_obs"].shape[2]
    f, gr = c_oracle.eval_cost_grad("joint", g["joint_x"], g["joint_points"], g["joint_obs"], N=N, V=V)
    assert np.allclose(f, g["joint_cost"], rtol=1e-12)
    assert np.allclose(gr, g["joint_grad"], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("n", [3, 10, 34])
def test_bfgs_update_and_scale(n):
    g = load_golden("bfgs_update")
    H = c_oracle.bfgs_update(g[f"H{n}"], g[f"s{n}"], g[f"y{n}"])
    assert np.allclose(H, g[f"Hout{n}"], rtol=1e-12, atol=1e-13)
    # rows 0 (negative curvature) and 1 (zero curvature): H untouched, bit for bit (test_bfgs_solver.py:335-361)
    assert np.array_equal(H[:2], g[f"H{n}"][:2])
    assert np.allclose(c_oracle.bfgs_initial_scale(g[f"s{n}"], g[f"y{n}"]), g[f"scale{n}"], rtol=1e-13)
    H32 = c_oracle.bfgs_update(*[g[k + str(n)].astype(np.float32) for k in ("H", "s", "y")])
    assert np.allclose(H32[2:], g[f"Hout{n}_f32"][2:], rtol=2e-3, atol=2e-3)


def test_bfgs_update_known_answer():
    """tests/autograd_solvers/test_bfgs_solver.py:307-332, eq. 6.17 written out."""
    g = load_golden("bfgs_update")
    s, y, H = g["kat_s"][0], g["kat_y"][0], g["kat_H"][0]
    c = s @ y
    left = np.eye(3) - np.outer(s, y) / c
    right = np.eye(3) - np.outer(y, s) / c
    expected = left @ H @ right + np.outer(s, s) / c
    got = c_oracle.bfgs_update(g["kat_H"], g["kat_s"], g["kat_y"])[0]
    assert np.allclose(got, expected)
    assert np.allclose(got, g["kat_Hout"][0], rtol=1e-13)


@pytest.mark.parametrize("dt", ["float32", "float64"])
@pytest.mark.parametrize("strong", [False, True])
def test_line_search_reference_cases(dt, strong):
    """test_wolffe_conditions.py:214-305: alpha < 1, alpha > 1, alpha == 0.25, alpha ~ 0."""
    g = load_golden("line_search")
    np_dt = np.dtype(dt)
    x = np.zeros((4, 2), np_dt)
    tg, d = g["dist_targets"].astype(np_dt), g["dist_dirs"].astype(np_dt)
    f0, gr = c_oracle.eval_cost_grad("distance", x, tg)
    a, probes = c_oracle.line_search("distance", x, d, f0, gr, tg, strong=strong)
    ref = g[f"dist_alpha_{dt}_{int(strong)}"]
    assert np.array_equal(a, ref)
    assert np.array_equal(probes, g[f"dist_probes_{dt}_{int(strong)}"])
    assert a[0] < 1.0 and a[1] > 1.0 and a[2] == 0.25 and abs(a[3]) < 1e-8


@pytest.mark.parametrize("name", ["sphere", "log_sphere", "rosenbrock", "cosine", "x2_sine"])
@pytest.mark.parametrize("strong", [False, True])
def test_line_search_analytic(name, strong):
    g = load_golden("line_search")
    x, d = g[f"{name}_x"], g[f"{name}_d"]
    f0, gr = c_oracle.eval_cost_grad(name, x)
    a, probes = c_oracle.line_search(name, x, d, f0, gr, sufficient_decrease=0.1, curvature=0.6, strong=strong)
    assert np.array_equal(probes, g[f"{name}_probes_{int(strong)}"])
    assert np.allclose(a, g[f"{name}_alpha_{int(strong)}"], rtol=1e-14)


def test_line_search_calibration_objective():
    g = load_golden("line_search")
    staged = c_oracle.stage(g["d10_points"], g["d10_obs"], g["d10_pose"])
    f0, gr = c_oracle.eval_cost_grad("distort10", g["d10_x"], staged, N=staged.shape[1])
    a, probes = c_oracle.line_search("distort10", g["d10_x"], g["d10_d"], f0, gr, staged, N=staged.shape[1], strong=True)
    assert np.array_equal(probes, g["d10_probes"])
    assert np.allclose(a, g["d10_alpha"], rtol=1e-12)


@pytest.mark.parametrize("name", ["sphere", "sphere_offset", "log_sphere", "rosenbrock", "cosine", "x2_sine"])
def test_analytic_solves(name):
    g = load_golden("analytic_solves")
    x0 = g[f"{name}_x0"]
    r = c_oracle.solve(name, x0.astype(np.float64), error_threshold=1e-6)
    # the hand-written gradients differ from autograd's in the last bits, which can move one count
    same = r["iters"] == g[f"{name}_float64_iters"]
    assert same.mean() >= 0.9
    assert np.array_equal(r["fevals"][same], g[f"{name}_float64_fevals"][same])
    assert np.array_equal(r["reason"], g[f"{name}_float64_reason"])
    assert np.allclose(r["x"][same], g[f"{name}_float64_x"][same], rtol=1e-6, atol=1e-8)
    assert np.allclose(r["cost"], g[f"{name}_float64_cost"], atol=2e-6)
    r32 = c_oracle.solve(name, x0.astype(np.float32), error_threshold=1e-6)
    assert (r32["iters"] == g[f"{name}_float32_iters"]).mean() >= 0.6  # fp32 trajectories are chaotic near the floor


BA_SHAPES = {"a": (4, 8), "b": (2, 5), "c": (3, 11), "d": (6, 7), "e": (3, 20), "f": (4, 30)}


@pytest.mark.parametrize("tag", sorted(BA_SHAPES))
def test_angle_ba_cost_gradient_line_search_match_reference(tag):
    """The entry script's objective (networks/calibration_network.py:58-67): the oracle's hand-written reverse
    mode against torch.autograd of the reference's own functions, and the line search on it."""
    g = load_golden("angle_ba")
    V, N = BA_SHAPES[tag]
    assert g[f"{tag}_unbatched_diff"] <= 1e-12  # batched (keepdim fix) == unmodified reference, one problem at a time
    x, obs, vis = g[f"{tag}_x"], g[f"{tag}_obs"], g[f"{tag}_vis"]
    f, gr = c_oracle.eval_cost_grad("angle_ba", x, obs, None, vis, N=N, V=V)
    assert np.allclose(f, g[f"{tag}_cost"], rtol=1e-13)
    assert np.all(np.abs(gr - g[f"{tag}_grad"]) <= 1e-12 * np.abs(g[f"{tag}_grad"]).max(axis=1, keepdims=True))
    f32, gr32 = c_oracle.eval_cost_grad("angle_ba", x.astype(np.float32), obs.astype(np.float32), None,
                                        vis.astype(np.float32), N=N, V=V)
    assert np.allclose(f32, g[f"{tag}_cost32"], rtol=2e-6)
    assert np.all(np.abs(gr32 - g[f"{tag}_grad32"]) <= 2e-5 * np.abs(g[f"{tag}_grad32"]).max(axis=1, keepdims=True))
    a, probes = c_oracle.line_search("angle_ba", x, g[f"{tag}_d"], f, gr, obs, None, vis, N=N, V=V, strong=True)
    assert np.array_equal(probes, g[f"{tag}_probes"])
    assert np.allclose(a, g[f"{tag}_alpha"], rtol=1e-12)


F64_CASES = ["solve_cfg2_f64", "solve_cfg2_pose_f64", "solve_cfg3_small_f64", "solve_cfg3_f64", "solve_cfg4_f64",
             "solve_cfg2_noisy_f64", "solve_ba_f64_30steps", "solve_ba_small_f64", "solve_ba_f64", "solve_ba_n75_f64",
             "solve_ba_n111_f64"]
F32_CASES = ["solve_cfg2_f32", "solve_cfg2_f32_thr1e-7", "solve_cfg3_f32", "solve_cfg4_f32", "solve_ba_f32_30steps",
             "solve_ba_f32"]


@pytest.mark.parametrize("name", F64_CASES + F32_CASES)
def test_solve_matches_reference_within_its_own_band(name):
    """Gates G64 / G32 for the oracle.  float64 on well-conditioned problems (configs 2, 3): identical
    accepted-step counts and termination reasons, parameters to ~1e-11.  Ill-conditioned (config 4), noisy
    and float32 runs are chaotic in the reference itself; there the bar is the reference's own band."""
    g = load_golden(name)
    batch = golden_batch(g["meta"])
    kw = g["meta"]["solver_kwargs"]
    r = c_oracle.solve_batch(batch, **kw)
    m = compare_solves(r, g, kw["error_threshold"])
    band = reference_band(g, kw["error_threshold"])
    print(name, "oracle vs reference", summary(m))
    print(name, "reference vs itself", summary(band))
    assert_within_band(m, band)
    if name in ("solve_cfg2_f64", "solve_cfg2_pose_f64", "solve_cfg3_small_f64", "solve_ba_f64_30steps",
                "solve_ba_small_f64", "solve_ba_n75_f64", "solve_ba_n111_f64"):
        assert m["steps_equal"] == 1.0 and m["fevals_equal"] == 1.0 and m["reason_equal"] == 1.0
        # the 2-view, 5-point bundle adjustment has nearly flat directions: the reference reproduces itself to
        # 9e-6 there under a permutation of the points; the north_star tolerance is 1e-4
        assert m["dtheta_max"] <= (1e-4 if name == "solve_ba_small_f64" else 1e-6)
