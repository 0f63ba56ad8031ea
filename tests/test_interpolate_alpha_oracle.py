"""CPU: the oracle's restatement of interpolate_alpha (utils/func_interpolate_alpha.py) and of the secant-zoom line
search against tests/golden/interpolate_alpha.npz (written from the unmodified reference function by
oracle/make_golden.py gen_interpolate_alpha) and the reference tests' known answers
(tests/utils/test_interpolate_alpha.py:29-62)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import c_oracle


def test_known_answers():
    g = load_golden("interpolate_alpha")
    k = g["kat_in"]
    out = c_oracle.interpolate_alpha(*[k[:, i].astype(np.float32) for i in range(4)])
    assert np.array_equal(out, g["kat_out"].astype(np.float32))  # the reference tests run in float32 and use ==
    out = c_oracle.interpolate_alpha(*[k[:, i] for i in range(4)])
    assert np.allclose(out, g["kat_out"], rtol=1e-15)


@pytest.mark.parametrize("name,dt", [("f32", np.float32), ("f64", np.float64)])
def test_forward_and_backward_match_reference(name, dt):
    g = load_golden("interpolate_alpha")
    ins = [g[f"{name}_in"][i] for i in range(4)]
    res = c_oracle.interpolate_alpha(*ins, grad_out=g[f"{name}_grad_out"])
    assert np.array_equal(res[0], g[f"{name}_out"], equal_nan=True)
    for got, want in zip(res[1:], g[f"{name}_grads"]):
        assert np.allclose(got, want, rtol=4 * np.finfo(dt).eps, atol=0, equal_nan=True)
    lo, hi = np.minimum(ins[0], ins[1]), np.maximum(ins[0], ins[1])
    ok = np.isfinite(res[0])
    assert np.all(res[0][ok] >= lo[ok]) and np.all(res[0][ok] <= hi[ok])   # test_result_is_between_alphas


def test_secant_zoom_line_search_satisfies_strong_wolfe_with_fewer_probes():
    """The composition (wolfe_conditions.py state machine + interpolate_alpha zoom step) has no reference
    implementation; it is held to the reference line-search tests' property (test_wolffe_conditions.py:152-211): the
    returned step satisfies both strong Wolfe inequalities, and on a quadratic-like problem it needs fewer probes."""
    rng = np.random.default_rng(5)
    B, n = 400, 6
    x = rng.normal(0, 3, (B, n))
    f0, g = c_oracle.eval_cost_grad("log_sphere", x)
    d = -g * rng.uniform(0.2, 30.0, (B, 1))
    c1, c2 = 0.1, 0.6
    res = {}
    for z in (False, True):
        a, probes = c_oracle.line_search("log_sphere", x, d, f0, g, sufficient_decrease=c1, curvature=c2, strong=True,
                                         zoom_interpolation=z)
        f1, g1 = c_oracle.eval_cost_grad("log_sphere", x + a[:, None] * d)
        g0 = (d * g).sum(1)
        assert np.all(f1 <= f0 + c1 * a * g0 + 1e-12)
        assert np.all(np.abs((d * g1).sum(1)) <= c2 * np.abs(g0) + 1e-12)
        res[z] = probes
    assert res[True].sum() < res[False].sum()
    print("probes: bisection", res[False].sum(), "secant", res[True].sum())
