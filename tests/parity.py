"""Shared comparison of two solve results (dicts with x, cost, iters, fevals, reason)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# The achieved parity numbers of every gated fixture are written here by the GPU suite (and mirrored into
# gpurun_out/, the only directory that travels back from the GPU box); the committed copy is the record.
PARITY_RECORD = os.environ.get("DAVO_PARITY_RECORD", os.path.join(ROOT, "profiles", "parity_r2.json"))
_RECORD_KEYS = ("steps_equal", "fevals_equal", "reason_equal", "dtheta_median", "dtheta_p99", "dtheta_max",
                "dcost_median", "dcost_p99", "dcost_max")


def record_parity(fixture, route, **comparisons):
    """Merge {fixture: {route: {comparison: metrics}}} into the parity record.  `comparisons` maps a label
    ("kernel_vs_reference", "reference_vs_itself", "kernel_vs_oracle") to a compare_solves() result."""
    for path in (PARITY_RECORD, os.path.join(ROOT, "gpurun_out", os.path.basename(PARITY_RECORD))):
        try:
            with open(path) as fh:
                data = json.load(fh)
        except (OSError, ValueError):
            data = {}
        entry = data.setdefault(fixture, {}).setdefault(route, {})
        for label, m in comparisons.items():
            entry[label] = {k: (float(m[k]) if not isinstance(m[k], (int, str)) else m[k])
                            for k in m if k in _RECORD_KEYS or k in ("problems", "note")}
            entry[label]["problems"] = int(len(m["dtheta"])) if "dtheta" in m else m.get("problems")
        try:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "w") as fh:
                json.dump(data, fh, indent=1, sort_keys=True)
        except OSError:
            pass


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


def assert_within_band(m, band, steps_slack=0.0, reason_slack=0.0, tol_factor=2.0, theta_floor=1e-4,
                       cost_floor=1e-5):
    """north_star tolerances (identical steps >= 99 %, dtheta <= 1e-4, dcost <= 1e-5) wherever the reference
    itself meets them; otherwise at least as tight as the reference's own band: the fraction of identical step
    counts / termination reasons may fall short of the reference's own self-agreement only by the 3-sigma
    sampling error of the two binomial estimates being compared, medians / p99 by a factor of `tol_factor`."""
    B = len(m["dtheta"])
    Bb = len(band["dtheta"]) if "dtheta" in band else B
    for key, slack in (("steps_equal", steps_slack), ("reason_equal", reason_slack)):
        p = band[key]
        # 3 sigma of the difference of two binomial estimates: the kernel's over B problems, the band's over Bb
        slack = max(slack, 3.0 * np.sqrt(max(p * (1.0 - p), 1e-4) * (1.0 / B + 1.0 / Bb)))
        assert m[key] >= min(0.99, p - slack), (key, summary(m), summary(band))
    # Where the reference's own tail is already of order one the trajectories are chaotic (float32 at its
    # noise floor, ill-conditioned config 4): a p99 of a heavy tail is not a stable statistic, compare medians.
    for key, floor in (("dtheta", theta_floor), ("dcost", cost_floor)):
        stat = "_p99" if band[key + "_p99"] <= 1e-2 else "_median"
        assert m[key + stat] <= max(floor, tol_factor * band[key + stat]), (key + stat, summary(m), summary(band))
