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
