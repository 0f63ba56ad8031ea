"""Shared comparison of two solve results (dicts with x, cost, iters, fevals, reason)."""
import numpy as np


def compare_solves(got, ref, error_threshold):
    """Metrics of SURVEY.md §8(d) 'parity protocol'."""
    x, xr = np.asarray(got["x"], np.float64), np.asarray(ref["x"], np.float64)
    both_finite = np.isfinite(x).all(1) & np.isfinite(xr).all(1)
    dth = np.abs(x - xr) / np.maximum(np.abs(xr), 1.0)
    dth = np.where(both_finite[:, None], dth, 0.0).max(axis=1)
    c, cr = np.asarray(got["cost"], np.float64), np.asarray(ref["cost"], np.float64)
    fin = np.isfinite(c) & np.isfinite(cr)
    dcost = np.where(fin, np.abs(c - cr) / np.maximum(np.abs(cr), error_threshold), 0.0)
    return dict(
        steps_equal=float((np.asarray(got["iters"]) == np.asarray(ref["iters"])).mean()),
        fevals_equal=float((np.asarray(got["fevals"]) == np.asarray(ref["fevals"])).mean()),
        reason_equal=float((np.asarray(got["reason"]) == np.asarray(ref["reason"])).mean()),
        finite_agree=float((np.isfinite(x).all(1) == np.isfinite(xr).all(1)).mean()),
        dtheta=dth, dtheta_median=float(np.median(dth)), dtheta_p99=float(np.quantile(dth, 0.99)),
        dtheta_max=float(dth.max()),
        dcost=dcost, dcost_median=float(np.median(dcost)), dcost_p99=float(np.quantile(dcost, 0.99)),
        dcost_max=float(dcost.max()),
    )


def summary(m):
    return {k: v for k, v in m.items() if not isinstance(v, np.ndarray)}


def reference_band(golden, error_threshold):
    """How well the reference reproduces ITSELF when only the order of the matches (the floating-point
    summation order) changes: fixtures carry a second reference run on permuted matches (SURVEY.md
    Appendix B).  A kernel cannot be asked to agree with the reference more tightly than this."""
    perm = {k[5:]: v for k, v in golden.items() if k.startswith("perm_")}
    return compare_solves(perm, golden, error_threshold)


def assert_within_band(m, band, steps_slack=0.04, reason_slack=0.03, tol_factor=4.0, theta_floor=1e-4,
                       cost_floor=1e-5):
    """north_star tolerances (identical steps >= 99 %, dtheta <= 1e-4, dcost <= 1e-5) wherever the reference
    itself meets them; otherwise at least as tight as the reference's own band (with statistical slack)."""
    assert m["steps_equal"] >= min(0.99, band["steps_equal"] - steps_slack), (summary(m), summary(band))
    assert m["reason_equal"] >= min(0.99, band["reason_equal"] - reason_slack), (summary(m), summary(band))
    assert m["dtheta_p99"] <= max(theta_floor, tol_factor * band["dtheta_p99"]), (summary(m), summary(band))
    assert m["dcost_p99"] <= max(cost_floor, tol_factor * band["dcost_p99"]), (summary(m), summary(band))
