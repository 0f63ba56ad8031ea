"""GPU: the secant-zoom line-search variant and interpolate_alpha through the Python surface ->
C-ABI, against the reference function's fixture and the C oracle's restatement of the variant."""
import numpy as np
import pytest
import torch

import davo_b200
from conftest import load_golden
from oracle import c_oracle

pytestmark = pytest.mark.gpu
TDT = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


@pytest.mark.parametrize("name,dt", [("f32", np.float32), ("f64", np.float64)])
def test_interpolate_alpha_matches_reference_fixture(name, dt):
    g = load_golden("interpolate_alpha")
    ins = [torch.from_numpy(g[f"{name}_in"][i]).requires_grad_(True) for i in range(4)]
    out = davo_b200.interpolate_alpha(*ins)
    assert out.shape == ins[0].shape and out.dtype == ins[0].dtype
    assert np.array_equal(out.detach().numpy(), g[f"{name}_out"], equal_nan=True)      # bit for bit
    grads = torch.autograd.grad((out * torch.from_numpy(g[f"{name}_grad_out"])).sum(), ins)
    for got, want in zip(grads, g[f"{name}_grads"]):
        assert np.allclose(got.numpy(), want, rtol=4 * np.finfo(dt).eps, atol=0, equal_nan=True)


def test_interpolate_alpha_reference_known_answers_and_shapes():
    """tests/utils/test_interpolate_alpha.py:6-62."""
    g = load_golden("interpolate_alpha")
    k = torch.from_numpy(g["kat_in"]).float()
    assert torch.equal(davo_b200.interpolate_alpha(k[:, 0], k[:, 1], k[:, 2], k[:, 3]), torch.from_numpy(g["kat_out"]).float())
    a1, a2, v1, v2 = (torch.randn(5, 1, 2, 3) for _ in range(4))
    r = davo_b200.interpolate_alpha(a1, a2, v1, v2)
    assert r.shape == (5, 1, 2, 3)
    assert torch.all(r >= torch.minimum(a1, a2)) and torch.all(r <= torch.maximum(a1, a2))


def test_interpolate_alpha_gradcheck():
    """tests/utils/test_interpolate_alpha.py:65-99 (interior zeros; the backward is the reference's custom one)."""
    torch.manual_seed(0)
    a1, a2, v1 = (torch.randn(100, dtype=torch.double) for _ in range(3))
    t = (0.1 + 0.8 * torch.rand(100, dtype=torch.double)) * (a2 - a1) + a1
    v2 = (a2 - t) * v1 / (a1 - t)
    ins = [x.clone().requires_grad_(True) for x in (a1, a2, v1, v2)]
    assert torch.autograd.gradcheck(davo_b200.interpolate_alpha, ins, eps=1e-6, atol=1e-4)


def _objective(batch, dt):
    if batch.model == "distort10":
        return davo_b200.DistortionObjective(torch.from_numpy(batch.points_3d), torch.from_numpy(batch.obs), dtype=dt)
    return davo_b200.JointPoseObjective(torch.from_numpy(batch.points_3d), torch.from_numpy(batch.obs), dtype=dt)


@pytest.mark.parametrize("model", ["distort10", "joint", "log_sphere"])
def test_secant_line_search_matches_oracle(model):
    rng = np.random.default_rng(11)
    if model == "distort10":
        b = davo_b200.synthetic.make_distort10(300, 64, seed=5, dtype=np.float64)
        obj, x = _objective(b, torch.float64), b.x0
        okw = dict(data0=c_oracle.stage(b.points_3d, b.obs, None), N=b.N)
    elif model == "joint":
        b = davo_b200.synthetic.make_joint(200, 32, 2, seed=6, dtype=np.float64)
        obj, x = _objective(b, torch.float64), b.x0
        okw = dict(data0=b.points_3d, data1=b.obs, N=b.N, V=b.views)
    else:
        x = rng.normal(0, 3, (300, 5))
        obj = davo_b200.AnalyticObjective(model, (300,), 5, dtype=torch.float64)
        okw = {}
    f0, g = c_oracle.eval_cost_grad(model, x, **okw)
    d = -g * rng.uniform(0.05, 20.0, (x.shape[0], 1))
    want_a, want_p = c_oracle.line_search(model, x, d, f0, g, strong=True, zoom_interpolation=True, **okw)
    T = torch.from_numpy
    got_a, got_p = davo_b200.line_search_wolfe_conditions(T(x), T(d), T(f0), T(g), obj, strong=True, return_probes=True,
                                                          zoom_interpolation=True)
    same = got_p.numpy() == want_p
    assert same.mean() >= 0.99
    assert np.allclose(got_a.numpy()[same], want_a[same], rtol=1e-9, atol=1e-12)
    plain_a, plain_p = davo_b200.line_search_wolfe_conditions(T(x), T(d), T(f0), T(g), obj, strong=True, return_probes=True)
    # the secant step usually needs fewer probes than bisection, not always (log-sphere with random step scales)
    print(model, "probes: secant", int(got_p.sum()), "bisection", int(plain_p.sum()))
    assert int(got_p.sum()) <= 1.1 * int(plain_p.sum())


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_secant_solve_matches_oracle_and_saves_evaluations(dt):
    b = davo_b200.synthetic.make_distort10(1024, 128, seed=21, dtype=dt)
    thr = 1e-5 if dt == np.float32 else 1e-10
    want = c_oracle.solve("distort10", b.x0, c_oracle.stage(b.points_3d, b.obs, None), N=b.N, error_threshold=thr,
                          zoom_interpolation=True)
    obj = _objective(b, TDT[np.dtype(dt)])
    solver = davo_b200.BFGSSolver(error_threshold=thr).eval()
    plain = solver(torch.from_numpy(b.x0), obj, return_info=True)
    solver.zoom_interpolation = True
    got = solver(torch.from_numpy(b.x0), obj, return_info=True)
    same = got.iterations.numpy() == want["iters"]
    print("identical steps", same.mean(), "evals secant", int(got.evaluations.sum()), "bisection", int(plain.evaluations.sum()))
    assert same.mean() >= (0.99 if dt == np.float64 else 0.9)
    tol = 1e-6 if dt == np.float64 else 2e-2
    assert np.allclose(got.parameters.numpy()[same], want["x"][same], rtol=tol, atol=tol)
    assert float(got.converged.float().mean()) >= float(plain.converged.float().mean()) - 0.02
    assert int(got.evaluations.sum()) <= 1.05 * int(plain.evaluations.sum())


# ---- the fused initial-guess network (csrc/mlp_kernels.cu) ------------------------------------------------------

def _reference_mlp(net, x):
    """The torch modules in float64 on the CPU (no GPU library call: the GPU test run must show no cuBLAS kernel)."""
    import copy
    ref = copy.deepcopy(net.initial_estimator).double().cpu().eval()
    with torch.no_grad():
        return ref(x.double().cpu())


@pytest.mark.parametrize("views,points,hidden,B", [(4, 8, -1, 300), (4, 8, -1, 65536), (2, 4, 64, 129), (3, 12, 144, 1000)])
def test_fused_initial_estimator_matches_torch_modules(views, points, hidden, B):
    torch.manual_seed(views * 100 + points)
    net = davo_b200.CalibrationNetwork(views, points, hidden_size=hidden).cuda()
    # non-trivial BatchNorm statistics and affine parameters
    for bn in (net.initial_estimator[2], net.initial_estimator[5]):
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.normal_(0, 0.2)
    net.eval()
    x = torch.randn(B, 2 * views * points, device="cuda")
    with torch.no_grad():
        got = net.estimate(x)
    want = _reference_mlp(net, x)
    assert got.shape == want.shape and got.dtype == torch.float32
    err = float((got.double().cpu() - want).abs().max() / want.abs().max())
    print("fused MLP", (views, points, hidden, B), "max error relative to max |output|", err)
    assert err <= 1e-5
    # weights changed in place -> the packed copy is rebuilt
    with torch.no_grad():
        net.initial_estimator[6].bias.add_(1.0)
        again = net.estimate(x)
    assert float((again - got - 1.0).abs().max()) <= 1e-5
    # training mode and grad mode run the torch modules
    assert net.estimate(x[:64]).requires_grad


# ---- N beyond the shared-memory slab: the matches are read from global memory (Distort10WideObjective<kGlobal>) -----

@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_distortion_objective_with_very_many_matches(dt):
    """The specialised kernels stage a problem's matches in a warp's shared-memory slab (N <= ~3 500); beyond that the
    C-ABI falls back to the generic solver, and beyond ITS slab (~14 000 in float32) to reading the matches from global
    memory: no N is refused.  Cost, gradient, line search and solve against the oracle."""
    for N in (6000, 20000):
        b = davo_b200.synthetic.make_distort10(6, N, seed=N, dtype=dt)
        obj = _objective(b, TDT[np.dtype(dt)])
        staged = c_oracle.stage(b.points_3d, b.obs, None)
        f_ref, g_ref = c_oracle.eval_cost_grad("distort10", b.x0, staged, N=N)
        f, g = obj.evaluate(torch.from_numpy(b.x0))
        tol = 2e-4 if dt == np.float32 else 1e-11
        assert np.allclose(f.cpu().numpy(), f_ref, rtol=tol) and np.allclose(g.cpu().numpy(), g_ref, rtol=tol, atol=tol * np.abs(g_ref).max())
        thr = 1e-5 * N / 256 if dt == np.float32 else 1e-10 * N / 256
        want = c_oracle.solve("distort10", b.x0, staged, N=N, error_threshold=thr)
        got = davo_b200.BFGSSolver(error_threshold=thr).eval()(torch.from_numpy(b.x0), obj, return_info=True)
        assert bool(got.converged.all()) and want["converged"].all()
        assert np.allclose(got.parameters.numpy(), want["x"], rtol=2e-2 if dt == np.float32 else 1e-6, atol=2e-3 if dt == np.float32 else 1e-7)
        if dt == np.float64:
            assert np.array_equal(got.iterations.numpy(), want["iters"])
            d = -g_ref
            a_ref, p_ref = c_oracle.line_search("distort10", b.x0, d, f_ref, g_ref, staged, N=N, strong=True)
            a, pr = davo_b200.line_search_wolfe_conditions(torch.from_numpy(b.x0), torch.from_numpy(d), torch.from_numpy(f_ref),
                                                           torch.from_numpy(g_ref), obj, strong=True, return_probes=True)
            assert np.array_equal(pr.numpy(), p_ref) and np.allclose(a.numpy(), a_ref, rtol=1e-9)


def test_joint_objective_with_more_than_nine_views():
    """n = 10 + 6 V > 64: the CTA-per-problem solve stops at 9 views; 10 .. 19 views run on the one-warp-per-problem
    wide solver with four components per lane.  Cost, gradient and solve against the oracle."""
    b = davo_b200.synthetic.make_joint(24, 40, 12, seed=77, dtype=np.float64)
    assert b.x0.shape[1] == 82
    obj = _objective(b, torch.float64)
    okw = dict(data0=b.points_3d, data1=b.obs, N=b.N, V=b.views)
    f_ref, g_ref = c_oracle.eval_cost_grad("joint", b.x0, **okw)
    f, g = obj.evaluate(torch.from_numpy(b.x0))
    assert np.allclose(f.cpu().numpy(), f_ref, rtol=1e-11) and np.allclose(g.cpu().numpy(), g_ref, rtol=1e-9, atol=1e-11 * np.abs(g_ref).max())
    want = c_oracle.solve("joint", b.x0, b.points_3d, b.obs, N=b.N, V=b.views, error_threshold=1e-10, iterations=300)
    got = davo_b200.BFGSSolver(error_threshold=1e-10, iterations=300).eval()(torch.from_numpy(b.x0), obj, return_info=True)
    same = got.iterations.numpy() == want["iters"]
    assert same.mean() >= 0.9
    assert np.allclose(got.parameters.numpy()[same], want["x"][same], rtol=1e-6, atol=1e-8)
    with pytest.raises(NotImplementedError):
        big = davo_b200.synthetic.make_joint(2, 8, 20, seed=1, dtype=np.float64)   # n = 130 > 128
        davo_b200.BFGSSolver().eval()(torch.from_numpy(big.x0), _objective(big, torch.float64))
