"""GPU: the training-mode / differentiable solve through BFGSSolver -> C-ABI
(davo_solve_training, davo_solve_backward) against
  * tests/golden/training.npz: the UNMODIFIED reference's BFGSSolver.train() with requires_grad parameters
    (create_graph=True path), returned parameters and d(sum(w * x_out))/d x0 (oracle/make_golden.py gen_training);
  * oracle/train_oracle.py (numpy restatement) for drop-path and return_second_last trajectories."""
import numpy as np
import pytest
import torch

import davo_b200
from conftest import load_golden
from oracle import train_oracle
from test_training_oracle import CASES, ILL_CONDITIONED, SETTINGS, golden_problem, relative_gradient_error

pytestmark = pytest.mark.gpu


def gpu_objective(g, case, dtype=torch.float64):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    if case == "d10":
        return davo_b200.DistortionObjective(t(g["d10_points"]), t(g["d10_obs"]), t(g["d10_pose"]), dtype=dtype)
    if case == "joint":
        return davo_b200.JointPoseObjective(t(g["joint_points"]), t(g["joint_obs"]), dtype=dtype)
    if case == "ba":
        return davo_b200.AngleDistanceObjective(t(g["ba_obs"]), t(g["ba_vis"]), dtype=dtype)
    x0 = g[f"{case}_x0"]
    return davo_b200.AnalyticObjective(case, x0.shape[:-1], x0.shape[-1], dtype=dtype)


def train_solver(skw, **extra):
    return davo_b200.BFGSSolver(drop_path_p=0.0, training_iterations=skw["training_iterations"],
                                training_error_threshold=skw["training_error_threshold"],
                                return_second_last=skw.get("return_second_last", False), **extra).train()


@pytest.mark.parametrize("setting", SETTINGS)
@pytest.mark.parametrize("case", CASES)
def test_training_solve_matches_reference_create_graph_path(case, setting):
    g = load_golden("training")
    skw = g["meta"]["settings"][setting]
    obj = gpu_objective(g, case)
    x0 = torch.from_numpy(g[f"{case}_x0"]).requires_grad_(True)
    w = torch.from_numpy(g[f"{case}_w"])
    x = train_solver(skw)(x0, obj)
    assert x.requires_grad and x.shape == x0.shape and x.device == x0.device
    (grad,) = torch.autograd.grad((x * w).sum(), x0)
    want_x, want_g = g[f"{case}_{setting}_x"], g[f"{case}_{setting}_grad_x0"]
    assert np.allclose(x.detach().numpy(), want_x, rtol=1e-7, atol=1e-9), np.abs(x.detach().numpy() - want_x).max()
    err = relative_gradient_error(grad.numpy(), want_g)
    print(case, setting, "relative gradient error", err)
    assert err <= ILL_CONDITIONED.get((case, setting), 1e-6), err


@pytest.mark.parametrize("case", ["d10", "ba", "rosenbrock"])
def test_training_solve_float32_inputs(case):
    """float32 problem sets: forward in float32, backward up-cast to float64; gradients agree with the float64
    reference to float32 accuracy on short chains."""
    g = load_golden("training")
    skw = g["meta"]["settings"]["k3"]
    obj = gpu_objective(g, case, dtype=torch.float32)
    x0 = torch.from_numpy(g[f"{case}_x0"]).float().requires_grad_(True)
    w = torch.from_numpy(g[f"{case}_w"]).float()
    x = train_solver(skw)(x0, obj)
    (grad,) = torch.autograd.grad((x * w).sum(), x0)
    assert grad.dtype == torch.float32
    want_x, want_g = g[f"{case}_k3_x"], g[f"{case}_k3_grad_x0"]
    assert np.allclose(x.detach().numpy(), want_x, rtol=2e-3, atol=2e-4)
    assert relative_gradient_error(grad.double().numpy(), want_g) <= 5e-2


def test_drop_path_matches_numpy_restatement_and_is_seeded():
    """drop-path (bfgs_solver.py:122-125): the kernel's counter-based draws are restated in numpy
    (train_oracle.drop_path_uniform); every problem stops at the iteration the restatement says, with the
    parameters the restated forward reaches, and torch.manual_seed reproduces a run."""
    g = load_golden("training")
    case = "d10"
    obj = gpu_objective(g, case)
    x0 = torch.from_numpy(g[f"{case}_x0"])
    B = x0.shape[0]
    p = 0.3
    solver = davo_b200.BFGSSolver(drop_path_p=p, training_iterations=12, training_error_threshold=1e-12).train()
    torch.manual_seed(1234)
    info = solver(x0, obj, return_info=True)
    torch.manual_seed(1234)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    torch.manual_seed(1234)
    again = solver(x0, obj, return_info=True)
    assert torch.equal(info.parameters, again.parameters) and torch.equal(info.iterations, again.iterations)
    prob = golden_problem(g, case)
    dropped = 0
    for b in range(B):
        drop = lambda k, b=b: not (train_oracle.drop_path_uniform(seed, b, k) > np.float32(p))
        x, traj = train_oracle.solve_forward(prob.row(b), g[f"{case}_x0"][b], error_threshold=1e-12, iterations=12, drop=drop)
        assert int(info.iterations[b]) == len(traj)
        assert np.allclose(info.parameters[b].numpy(), x, rtol=1e-7, atol=1e-9)
        if len(traj) < 12:
            dropped += 1
            assert int(info.reason[b]) == 5
    assert dropped >= 1
    # a different seed retires different problems
    torch.manual_seed(99)
    other = solver(x0, obj, return_info=True)
    assert not torch.equal(other.iterations, info.iterations)


def test_drop_path_statistics():
    """P(still updating after k iterations) = (1 - p)^(k + 1) for problems that never converge."""
    B = 20000
    obj = davo_b200.AnalyticObjective("sphere_offset", (B,), 3, dtype=torch.float32)
    torch.manual_seed(3)
    x0 = torch.randn(B, 3) * 50.0
    solver = davo_b200.BFGSSolver(drop_path_p=0.1, training_iterations=1, training_error_threshold=1e-12).train()
    info = solver(x0, obj, return_info=True)
    frac_dropped_at_0 = float((info.iterations == 0).float().mean())
    assert abs(frac_dropped_at_0 - 0.1) < 4 * np.sqrt(0.1 * 0.9 / B)


def test_differentiable_eval_mode_and_no_grad():
    """create_graph = parameters.requires_grad also in eval mode (bfgs_solver.py:85); under torch.no_grad or with
    plain parameters nothing is recorded and the result carries no graph."""
    g = load_golden("training")
    obj = gpu_objective(g, "joint")
    x0 = torch.from_numpy(g["joint_x0"])
    solver = davo_b200.BFGSSolver(iterations=5, error_threshold=1e-12).eval()
    plain = solver(x0, obj)
    assert not plain.requires_grad
    xg = x0.clone().requires_grad_(True)
    with torch.no_grad():
        assert not solver(xg, obj).requires_grad
    out = solver(xg, obj)
    assert out.requires_grad
    assert np.allclose(out.detach().numpy(), plain.numpy(), rtol=1e-9, atol=1e-11)
    (grad,) = torch.autograd.grad(out.square().sum(), xg)
    assert torch.isfinite(grad).all() and float(grad.abs().max()) > 0


def test_gradient_matches_finite_differences_of_the_solve():
    """d x_out / d x0 against central differences of the (training-mode, fixed-length) solve itself."""
    g = load_golden("training")
    obj = gpu_objective(g, "d10")
    x0 = torch.from_numpy(g["d10_x0"])[:2].contiguous()
    obj = davo_b200.DistortionObjective(torch.from_numpy(g["d10_points"][:2]), torch.from_numpy(g["d10_obs"][:2]),
                                        torch.from_numpy(g["d10_pose"][:2]))
    solver = davo_b200.BFGSSolver(drop_path_p=0.0, training_iterations=3, training_error_threshold=1e-14).train()
    w = torch.from_numpy(g["d10_w"])[:2]
    xg = x0.clone().requires_grad_(True)
    (grad,) = torch.autograd.grad((solver(xg, obj) * w).sum(), xg)
    fd = torch.zeros_like(x0)
    h = 1e-6
    for c in range(x0.shape[1]):
        e = torch.zeros_like(x0)
        e[:, c] = h
        fd[:, c] = ((solver(x0 + e, obj) - solver(x0 - e, obj)) * w).sum(dim=1) / (2 * h)
    # alpha is piecewise constant in x0: the finite difference is exact unless a bisection decision flips inside +-h
    assert np.allclose(grad.numpy(), fd.numpy(), rtol=2e-4, atol=1e-6 * float(fd.abs().max()))


def test_calibration_network_trains_through_the_solve():
    """CalibrationNetwork.train() (networks/calibration_network.py:54-73 driven as Trainer.fit drives it): the loss
    back-propagates through the solve into the MLP's weights and one optimiser step changes them."""
    torch.manual_seed(0)
    b = davo_b200.synthetic.make_angle_ba(32, 8, 4, seed=12, dtype=np.float64)
    net = davo_b200.CalibrationNetwork(4, 8).double().cuda().train()
    net.solver.training_iterations = 6
    obs, vis = torch.from_numpy(b.obs).cuda(), torch.from_numpy(b.weights).cuda() > 0
    opt = torch.optim.SGD(net.parameters(), lr=1e-3)
    params, err = net(obs, vis, return_error=True)
    assert params.requires_grad and params.shape == (32, 45)
    truth = torch.from_numpy(b.truth).cuda()
    loss = (params - truth).square().mean()
    loss.backward()
    grads = [p.grad for p in net.initial_estimator.parameters()]
    assert all(gr is not None and torch.isfinite(gr).all() for gr in grads)
    assert any(float(gr.abs().max()) > 0 for gr in grads)
    before = [p.detach().clone() for p in net.initial_estimator.parameters()]
    opt.step()
    assert any(not torch.equal(a, p.detach()) for a, p in zip(before, net.initial_estimator.parameters()))


def _observation_case(g, case):
    """(objective, observations leaf, x0, w) of one fixture case with observations that require grad."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    obs = t(g[f"{case}_obs"]).requires_grad_(True)
    if case == "d10":
        obj = davo_b200.DistortionObjective(t(g["d10_points"]), obs, t(g["d10_pose"]))
    elif case == "joint":
        obj = davo_b200.JointPoseObjective(t(g["joint_points"]), obs)
    else:
        obj = davo_b200.AngleDistanceObjective(obs, t(g["ba_vis"]))
    return obj, obs, t(g[f"{case}_x0"]), t(g[f"{case}_w"])


@pytest.mark.parametrize("setting", SETTINGS)
@pytest.mark.parametrize("case", ["d10", "joint", "ba"])
def test_gradient_with_respect_to_observations(case, setting):
    """d(sum(w * x_out)) / d observations of every camera objective against the reference's autograd (DISTORT10:
    analytic J v; JOINT and ANGLE_BA: from the evaluations of the Hessian-vector difference)."""
    g = load_golden("training")
    skw = g["meta"]["settings"][setting]
    obj, obs, x0, w = _observation_case(g, case)
    x0 = x0.clone().requires_grad_(True)
    x = train_solver(skw)(x0, obj)
    g_x0, g_obs = torch.autograd.grad((x * w).sum(), [x0, obs])
    assert g_obs.shape == obs.shape
    assert relative_gradient_error(g_x0.numpy(), g[f"{case}_{setting}_grad_x0"]) <= ILL_CONDITIONED.get((case, setting), 1e-6)
    want = g[f"{case}_{setting}_grad_obs"]
    err = np.abs(g_obs.numpy() - want).max() / np.abs(want).max()
    print(case, setting, "observation gradient error", err)
    assert err <= ILL_CONDITIONED.get((case, setting), 1e-6)
    # data gradient alone (parameters do not require grad)
    x = train_solver(skw)(x0.detach(), obj)
    (g_only,) = torch.autograd.grad((x * w).sum(), [obs])
    assert torch.equal(g_only, g_obs)
