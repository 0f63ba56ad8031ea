"""The training-mode restatement (oracle/train_oracle.py) against the unmodified reference's create_graph path:
returned parameters and d(sum(w * x_out))/d x0 of tests/golden/training.npz (oracle/make_golden.py gen_training)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import c_oracle, train_oracle

CASES = ["rosenbrock", "log_sphere", "sphere", "cosine", "d10", "joint", "ba"]
SETTINGS = ["k3", "k8", "k25thr", "k8second"]


# Chains whose reference gradient has exploded (|d x_out / d x0| of 1e5 .. 1e7: 25 BFGS updates with 1/(y.s) factors
# differentiated through): the reference's own autograd result is then sensitive to rounding at the 1e-5 .. 1 level
# (with an EXACT Hessian-vector product this restatement still differs by 3.3e-5 on rosenbrock).  Looser bound /
# parameters only.
# log_sphere after 8 iterations has converged (y.s ~ 1e-12: the update's 1/(y.s) factors are ~1e12 and the true
# gradient ~1e-6); the absolute error there is 4e-7.
ILL_CONDITIONED = {("rosenbrock", "k25thr"): 1e-3, ("ba", "k25thr"): np.inf, ("log_sphere", "k8"): 1e-4,
                   ("log_sphere", "k8second"): 1e-4}


def golden_problem(g, case):
    if case == "d10":
        staged = c_oracle.stage(g["d10_points"], g["d10_obs"], g["d10_pose"])
        return train_oracle.Problem("distort10", staged, N=staged.shape[1])
    if case == "joint":
        return train_oracle.Problem("joint", g["joint_points"], g["joint_obs"], N=g["joint_points"].shape[1],
                                    V=g["joint_obs"].shape[1])
    if case == "ba":
        return train_oracle.Problem("angle_ba", g["ba_obs"], None, g["ba_vis"], N=g["ba_obs"].shape[2], V=g["ba_obs"].shape[1])
    return train_oracle.Problem(case)


def relative_gradient_error(got, want):
    """max over problems of |got - want| / max |want| (per problem; the gradient of a problem whose chain has
    exploded is compared at its own scale)."""
    scale = np.maximum(np.abs(want).max(axis=1, keepdims=True), 0.05)  # |w| ~ 1: a converged problem's gradient is ~0
    return float((np.abs(got - want) / scale).max())


@pytest.mark.parametrize("setting", SETTINGS)
@pytest.mark.parametrize("case", CASES)
def test_training_restatement_matches_reference_autograd(case, setting):
    g = load_golden("training")
    skw = g["meta"]["settings"][setting]
    x, grad = train_oracle.solve_with_grad(
        golden_problem(g, case), g[f"{case}_x0"], g[f"{case}_w"], error_threshold=skw["training_error_threshold"],
        iterations=skw["training_iterations"], second_last=skw.get("return_second_last", False))
    want_x, want_g = g[f"{case}_{setting}_x"], g[f"{case}_{setting}_grad_x0"]
    assert np.allclose(x, want_x, rtol=1e-7, atol=1e-9), np.abs(x - want_x).max()
    err = relative_gradient_error(grad, want_g)
    print(case, setting, "relative gradient error", err)
    assert err <= ILL_CONDITIONED.get((case, setting), 1e-6), err
