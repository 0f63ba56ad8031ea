"""numpy restatement of the differentiable / training-mode solve — TEST INFRASTRUCTURE (checker only).

Forward: BFGSSolver.forward in training mode for one problem at a time (autograd_solvers/bfgs_solver.py:80-215:
training thresholds :88-93, drop-path :122-125, return_second_last :196-212), recording the iterates.
Backward: reverse mode through the recorded iterates, which is what `create_graph=True` (:85,133-135) makes
torch.autograd do: the step length alpha is a constant (the line search detaches everything,
line_search/wolfe_conditions.py:70-73), every gradient g_k = grad f(x_k) is a differentiable function of x_k
(its vector-Jacobian product is a Hessian-vector product), and the eq. 6.20 scale (:217-233), the eq. 6.17 update
(:235-303) and InverseCurvature's custom backward (utils/func_inverse_curvature.py:22-37) are differentiated as
written.  Hessian-vector products are central differences of the analytic gradient (the product kernel does the
same: fourth-order, accurate to ~1e-10), so agreement with the reference's autograd is ~1e-7 relative or better on
well-conditioned chains, not exact.

Built on the C oracle's evaluator and line search (c_oracle).  Pinned by tests/golden/training.npz, which
oracle/make_golden.py writes from the unmodified reference (tests/test_training_oracle.py).
"""
from __future__ import annotations

import numpy as np

from . import c_oracle


class Problem:
    """One problem's view of a batched objective (rows b of the data arrays)."""

    def __init__(self, model, data0=None, data1=None, weights=None, N=0, V=1):
        self.model, self.N, self.V = model, N, V
        self.data0, self.data1, self.weights = data0, data1, weights

    def row(self, b):
        pick = lambda a: None if a is None else np.ascontiguousarray(a[b:b + 1])
        return Problem(self.model, pick(self.data0), pick(self.data1), pick(self.weights), self.N, self.V)

    def eval(self, x):
        f, g = c_oracle.eval_cost_grad(self.model, np.ascontiguousarray(x[None]), self.data0, self.data1, self.weights,
                                       N=self.N, V=self.V)
        return float(f[0]), g[0]

    def line_search(self, x, d, f, g, c1, c2):
        a, _ = c_oracle.line_search(self.model, x[None], d[None], np.array([f]), g[None], self.data0, self.data1,
                                    self.weights, N=self.N, V=self.V, sufficient_decrease=c1, curvature=c2, strong=True)
        return float(a[0])

    def hvp(self, x, v, rel_step=5e-7):
        """Hessian-vector product by a fourth-order central difference of the analytic gradient:
        (-g(x+2hv) + 8 g(x+hv) - 8 g(x-hv) + g(x-2hv)) / 12h, perturbation |hv| = rel_step * (1 + |x|)."""
        nv = np.linalg.norm(v)
        if nv == 0.0:
            return np.zeros_like(x)
        h = rel_step * (1.0 + np.linalg.norm(x)) / nv
        g = lambda t: self.eval(x + t * h * v)[1]
        return (-g(2.0) + 8.0 * g(1.0) - 8.0 * g(-1.0) + g(-2.0)) / (12.0 * h)


def bfgs_update(H, s, y):
    """eq. 6.17 as bfgs_solver.py:263-303 writes it; returns (H+, intermediates for the backward pass)."""
    sy = float(s @ y)
    rho = 1.0 / sy if sy > 0.0 else 0.0
    b = H.T @ y            # (y^T H)
    c = H @ y
    w = y * rho
    a = s * rho
    q = float(b @ w)
    Hn = H + (1.0 + q) * np.outer(a, s) - np.outer(a, b) - np.outer(c, a)
    return Hn, (rho, a, b, c, w, q)


def solve_forward(prob: Problem, x0, *, error_threshold, iterations, minimum_step=1e-8, c1=1e-4, c2=0.9,
                  drop=None, second_last=False):
    """One problem.  drop(k) -> True retires the problem at the top of iteration k (drop-path).
    Returns (x_out, trajectory) with trajectory = list of (x_k, g_k, alpha_k applied)."""
    x = np.array(x0, dtype=np.float64)
    traj = []
    H = np.eye(len(x))
    g_prev = d = s = None
    for k in range(iterations):
        if drop is not None and drop(k):
            break
        f, g = prob.eval(x)
        if not (f > error_threshold):
            break
        if k == 0:
            d = -g
        else:
            y = g - g_prev
            if k == 1:
                den = max(float(y @ y), 1e-5)
                H = max(float(s @ y) / den, 1e-4) * H
            H, _ = bfgs_update(H, s, y)
            d = -(H @ g)
        alpha = prob.line_search(x, d, f, g, c1, c2)
        s = alpha * d
        stop = not (np.linalg.norm(s) > minimum_step)
        applied = not (second_last and stop)   # return_second_last: the step that retires the problem is not applied
        traj.append((x.copy(), g.copy(), alpha if applied else 0.0))
        if applied:
            x = x + s
        g_prev = g
        if stop:
            break
    return x, traj


def solve_backward(prob: Problem, traj, xbar_out):
    """d loss / d x0 given d loss / d x_out, through the recorded iterates."""
    K = len(traj)
    n = len(xbar_out)
    xbar = np.array(xbar_out, dtype=np.float64)
    if K == 0:
        return xbar
    X = [t[0] for t in traj]
    G = [t[1] for t in traj]
    A = [t[2] for t in traj]
    # replay: H_k (after the update at iteration k), d_k
    Hs, D = [None] * K, [None] * K
    H = np.eye(n)
    for k in range(K):
        if k == 0:
            D[0] = -G[0]
        else:
            s, y = A[k - 1] * D[k - 1], G[k] - G[k - 1]
            if k == 1:
                H = max(float(s @ y) / max(float(y @ y), 1e-5), 1e-4) * H
            H, _ = bfgs_update(H, s, y)
            Hs[k] = H
            D[k] = -(H @ G[k])
    sbar_next = np.zeros(n)
    gbar_carry = np.zeros(n)
    Hbar = np.zeros((n, n))
    for k in range(K - 1, -1, -1):
        sbar = sbar_next + xbar                    # x_{k+1} = x_k + s_k  and  H_{k+1} = U(H_k, s_k, y_k)
        dbar = A[k] * sbar                         # s_k = alpha_k d_k, alpha_k constant
        gbar = gbar_carry.copy()                   # from y_k = g_{k+1} - g_k
        if k == 0:
            gbar -= dbar                           # d_0 = -g_0
        else:
            H, g = Hs[k], G[k]
            gbar -= H.T @ dbar                     # d_k = -H_k g_k
            Hbar = Hbar - np.outer(dbar, g)
            s, y = A[k - 1] * D[k - 1], G[k] - G[k - 1]
            if k == 1:
                sy, yy = float(s @ y), float(y @ y)
                den = max(yy, 1e-5)
                scale = max(sy / den, 1e-4)
                Hp = scale * np.eye(n)
            else:
                Hp = Hs[k - 1]
            _, (rho, a, b, c, w, q) = bfgs_update(Hp, s, y)
            Gm = Hbar
            Gs, Gta, Gb, Ga, Gtc = Gm @ s, Gm.T @ a, Gm @ b, Gm @ a, Gm.T @ c
            qbar = float(a @ Gs)
            abar = (1.0 + q) * Gs - Gb - Gtc
            sbar_p = (1.0 + q) * Gta
            bbar = -Gta + qbar * w
            cbar = -Ga
            wbar = qbar * b
            Hpbar = Gm + np.outer(y, bbar) + np.outer(cbar, y)
            ybar = Hp @ bbar + Hp.T @ cbar + rho * wbar
            rhobar = float(wbar @ y) + float(abar @ s)
            sbar_p = sbar_p + rho * abar
            t = -rho * rho * rhobar               # InverseCurvature.backward
            sbar_p = sbar_p + t * y
            ybar = ybar + t * s
            if k == 1:
                scalebar = float(np.trace(Hpbar))  # H'_0 = scale * I
                if sy / den >= 1e-4:               # clamp(min=1e-4) passes the gradient where it did not clamp
                    numbar = scalebar / den
                    denbar = -scalebar * sy / (den * den)
                    sbar_p = sbar_p + numbar * y
                    ybar = ybar + numbar * s
                    if yy >= 1e-5:
                        ybar = ybar + denbar * 2.0 * y
                Hbar = np.zeros((n, n))
            else:
                Hbar = Hpbar
            gbar = gbar + ybar
            gbar_carry = -ybar
            sbar_next = sbar_p
        xbar = xbar + prob.hvp(X[k], gbar)         # g_k = grad f(x_k)
    return xbar


def solve_with_grad(prob_batch: Problem, x0, w, **kw):
    """Batch driver for tests: x_out[B,n] and d sum(w * x_out) / d x0."""
    xs, grads = [], []
    for b in range(x0.shape[0]):
        prob = prob_batch.row(b)
        x, traj = solve_forward(prob, x0[b], **kw)
        xs.append(x)
        grads.append(solve_backward(prob, traj, w[b]))
    return np.stack(xs), np.stack(grads)


def drop_path_uniform(seed, problem, iteration):
    """The uniform in [0, 1) the training kernel draws for (problem, iteration) — csrc/train_params.cuh
    drop_path_bits: Philox4x32-10 with counter (problem, 0, iteration, 'DROP'), key = seed; word 0, top 24 bits.
    The problem keeps updating iff u > drop_path_p (bfgs_solver.py:122-125)."""
    from . import gen_oracle
    c = np.array([[problem & 0xFFFFFFFF, 0, iteration, 0x44524F50]], dtype=np.uint64)
    w = gen_oracle.philox4x32(c, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return float(np.float32(int(w[0, 0]) >> 8) * np.float32(1.0 / 16777216.0))
