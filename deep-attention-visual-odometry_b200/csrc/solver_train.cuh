// solver_train.cuh — the backward pass of the differentiable (create_graph=True) solve, one warp per problem.
//
// What the reference does: with `parameters.requires_grad` BFGSSolver.forward evaluates every gradient with
// create_graph=True (autograd_solvers/bfgs_solver.py:85,133-135), so torch.autograd later differentiates the whole
// unrolled iteration: x_{k+1} = x_k + alpha_k d_k (:191-199), d_0 = -g_0 (:152-155), d_k = -H_k g_k (:173-176),
// H_k = U(H_{k-1}, s_{k-1}, y_{k-1}) (eq. 6.17, :235-303) with the eq. 6.20 scale at k = 1 (:159-167, :217-233) and
// InverseCurvature's custom backward (utils/func_inverse_curvature.py:22-37).  The step length alpha_k is a constant:
// the line search detaches everything it touches (line_search/wolfe_conditions.py:70-73).  Every g_k = grad f(x_k) is
// a differentiable function of x_k; its vector-Jacobian product is a Hessian-vector product.
//
// Here: the training-mode forward (solver_wide.cuh, TrainRecorder) keeps (x_k, g_k, alpha_k) per accepted step; this
// kernel (a) replays the inverse-Hessian recursion from the recorded gradients, parking every H_k and d_k in a
// caller-provided scratch area, and (b) runs the reverse sweep over k = K-1 .. 0 with the adjoint of each statement
// written out by hand (oracle/train_oracle.py is the numpy statement of the same sweep, pinned to the reference's
// autograd by tests/golden/training.npz).  Hessian-vector products are fourth-order central differences of the
// analytic gradient kernel, (-g(x+2hv) + 8 g(x+hv) - 8 g(x-hv) + g(x-2hv)) / 12h with |hv| = rel_step (1 + |x|):
// ~1e-10 relative in float64, which is why this pass always runs in float64 (the Python wrapper up-casts a float32
// trajectory).
#pragma once
#include "davo_common.cuh"
#include "solver_wide.cuh"
#include "objectives_wide.cuh"
#include "objectives_joint.cuh"
#include "objectives_ba.cuh"

namespace davo {

// JOINT and ANGLE_BA observations [V, N, 2]: taken from the evaluations of the Hessian-vector difference (see
// DataGradient in objectives_wide.cuh).
template <typename T>
struct DataGradient<JointObjective<T, true>, T> {
    static constexpr bool kSupported = true;
    __device__ static int elements(const SolveParams<T>& p) { return 2 * p.V * p.N; }
    __device__ static void analytic(JointObjective<T, true>&, const T*, const T*, T*, int) {}
    __device__ static void arm(JointObjective<T, true>& obj, T* out, T coef) { obj.dgrad = out; obj.dcoef = coef; }
};
template <typename T>
struct DataGradient<AngleBAObjective<T, 0, 0, true>, T> {
    static constexpr bool kSupported = true;
    __device__ static int elements(const SolveParams<T>& p) { return 2 * p.V * p.N; }
    __device__ static void analytic(AngleBAObjective<T, 0, 0, true>&, const T*, const T*, T*, int) {}
    __device__ static void arm(AngleBAObjective<T, 0, 0, true>& obj, T* out, T coef) { obj.dgrad = out; obj.dcoef = coef; }
};

template <typename T>
struct BackwardWorkspace {
    static constexpr int kVecs = 28;
    T* vec[kVecs];
    T* M;    // n x ld: H during the replay, Hbar during the reverse sweep
    int ld;
    __host__ __device__ static size_t bytes(int n) {
        return sizeof(T) * ((size_t)kVecs * wide_vec(n) + (size_t)n * (n | 1));
    }
    __device__ void carve(unsigned char* base, int n) {
        T* p = reinterpret_cast<T*>(base);
        const int v = wide_vec(n);
        for (int i = 0; i < kVecs; ++i) vec[i] = p + (size_t)i * v;
        M = p + (size_t)kVecs * v;
        ld = n | 1;
    }
};

namespace train {

template <typename T>
__device__ __forceinline__ T dot(const T* a, const T* b, int n, int lane) {
    T acc = T(0);
    for (int c = lane; c < n; c += 32) acc = fma_t(a[c], b[c], acc);
    return warp_allreduce(acc);
}
// The products below are what the backward pass spends its time in, and with two warps per scheduler their cost is
// latency, not throughput: every routine keeps several independent sums going (two partial sums per product, several
// products per sweep over the matrix) instead of one chain of n dependent FMAs per product per sweep.

// out_k = M v_k, k < R (lane c walks row c of M once for all R products)
template <typename T, int R>
__device__ __forceinline__ void mv_multi(const T* M, int ld, const T* const (&v)[R], T* const (&out)[R], int n, int lane) {
    __syncwarp();
    for (int c = lane; c < n; c += 32) {
        T e[R], o[R];
#pragma unroll
        for (int k = 0; k < R; ++k) e[k] = o[k] = T(0);
        const T* row = M + (size_t)c * ld;
        int j = 0;
        for (; j + 2 <= n; j += 2) {
            const T m0 = row[j], m1 = row[j + 1];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                e[k] = fma_t(m0, v[k][j], e[k]);
                o[k] = fma_t(m1, v[k][j + 1], o[k]);
            }
        }
        if (j < n) {
            const T m0 = row[j];
#pragma unroll
            for (int k = 0; k < R; ++k) e[k] = fma_t(m0, v[k][j], e[k]);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) out[k][c] = e[k] + o[k];
    }
    __syncwarp();
}
// out_k = M^T v_k, k < R (lane c walks column c)
template <typename T, int R>
__device__ __forceinline__ void mtv_multi(const T* M, int ld, const T* const (&v)[R], T* const (&out)[R], int n, int lane) {
    __syncwarp();
    for (int c = lane; c < n; c += 32) {
        T e[R], o[R];
#pragma unroll
        for (int k = 0; k < R; ++k) e[k] = o[k] = T(0);
        const T* col = M + c;
        int j = 0;
        for (; j + 2 <= n; j += 2) {
            const T m0 = col[(size_t)j * ld], m1 = col[(size_t)(j + 1) * ld];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                e[k] = fma_t(m0, v[k][j], e[k]);
                o[k] = fma_t(m1, v[k][j + 1], o[k]);
            }
        }
        if (j < n) {
            const T m0 = col[(size_t)j * ld];
#pragma unroll
            for (int k = 0; k < R; ++k) e[k] = fma_t(m0, v[k][j], e[k]);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) out[k][c] = e[k] + o[k];
    }
    __syncwarp();
}
// out_r = M u (row walk) and out_c = M^T w (column walk) in one loop
template <typename T>
__device__ __forceinline__ void mv_mtv(const T* M, int ld, const T* u, const T* w, T* out_r, T* out_c, int n, int lane) {
    __syncwarp();
    for (int c = lane; c < n; c += 32) {
        T r0 = T(0), r1 = T(0), c0 = T(0), c1 = T(0);
        const T* row = M + (size_t)c * ld;
        const T* col = M + c;
        int j = 0;
        for (; j + 2 <= n; j += 2) {
            r0 = fma_t(row[j], u[j], r0);
            c0 = fma_t(col[(size_t)j * ld], w[j], c0);
            r1 = fma_t(row[j + 1], u[j + 1], r1);
            c1 = fma_t(col[(size_t)(j + 1) * ld], w[j + 1], c1);
        }
        if (j < n) {
            r0 = fma_t(row[j], u[j], r0);
            c0 = fma_t(col[(size_t)j * ld], w[j], c0);
        }
        out_r[c] = r0 + r1;
        out_c[c] = c0 + c1;
    }
    __syncwarp();
}
template <typename T>
__device__ __forceinline__ void mv(const T* M, int ld, const T* v, T* out, int n, int lane) {
    const T* const vs[1] = {v};
    T* const os[1] = {out};
    mv_multi<T, 1>(M, ld, vs, os, n, lane);
}
template <typename T>
__device__ __forceinline__ void mtv(const T* M, int ld, const T* v, T* out, int n, int lane) {
    const T* const vs[1] = {v};
    T* const os[1] = {out};
    mtv_multi<T, 1>(M, ld, vs, os, n, lane);
}

// eq. 6.17 intermediates from (Hp, s, y): rho, a = s rho, b = Hp^T y, c = Hp y, w = y rho, q = b . w.
// Hp == nullptr stands for scale * I (the k = 1 update).
template <typename T>
__device__ __forceinline__ void update_terms(const T* Hp, int ld, T scale, const T* s, const T* y, T* a, T* bv, T* cv,
                                             T* w, T& rho, T& q, int n, int lane) {
    const T sy = dot(s, y, n, lane);
    rho = sy > T(0) ? T(1) / sy : T(0);   // func_inverse_curvature.py:8-11
    if (Hp) {
        mv_mtv(Hp, ld, y, y, cv, bv, n, lane);   // cv = Hp y, bv = Hp^T y
    } else {
        __syncwarp();
        for (int c = lane; c < n; c += 32) bv[c] = cv[c] = scale * y[c];
    }
    __syncwarp();
    for (int c = lane; c < n; c += 32) {
        a[c] = s[c] * rho;
        w[c] = y[c] * rho;
    }
    __syncwarp();
    q = dot(bv, w, n, lane);
}

}  // namespace train

// One problem's backward pass.  `obj` is bound to problem b.
template <typename T, typename Obj>
__device__ __forceinline__ void backward_one(Obj& obj, const SolveParams<T>& p, const BackwardParams<T>& bp, int b,
                                             BackwardWorkspace<T>& ws, int lane) {
    using namespace train;
    const int n = p.n, ld = ws.ld;
    const int K = bp.traj_len[b];
    T* xbar = ws.vec[0];  T* sbar_next = ws.vec[1]; T* gcarry = ws.vec[2]; T* dbar = ws.vec[3];
    T* gbar = ws.vec[4];  T* s = ws.vec[5];  T* y = ws.vec[6];  T* a = ws.vec[7];  T* bv = ws.vec[8];
    T* cv = ws.vec[9];    T* w = ws.vec[10]; T* Gs = ws.vec[11]; T* Gta = ws.vec[12]; T* Gb = ws.vec[13];
    T* Ga = ws.vec[14];   T* Gtc = ws.vec[15]; T* abar = ws.vec[16]; T* sbar_p = ws.vec[17]; T* bbar = ws.vec[18];
    T* cbar = ws.vec[19]; T* wbar = ws.vec[20]; T* ybar = ws.vec[21]; T* xt = ws.vec[22]; T* gt = ws.vec[23];
    T* hacc = ws.vec[24]; T* gk = ws.vec[25]; T* gprev = ws.vec[26]; T* dprev = ws.vec[27];
    T* M = ws.M;
    const size_t per_step = (size_t)n * n + n;
    T* base = bp.scratch + (size_t)bp.scratch_offset[b] * per_step;
    const size_t row0 = (size_t)bp.traj_offset[b];
    const T* X = bp.traj_x + row0 * n;
    const T* G = bp.traj_g + row0 * n;
    const T* A = bp.traj_alpha + row0;
    __syncwarp();
    for (int c = lane; c < n; c += 32) xbar[c] = bp.grad_out[(size_t)b * n + c];
    using DG = DataGradient<Obj, T>;
    T* dout = (DG::kSupported && bp.grad_data) ? bp.grad_data + (size_t)b * DG::elements(p) : nullptr;
    if (dout)
        for (int i = lane; i < DG::elements(p); i += 32) dout[i] = T(0);
    __syncwarp();
    if (K > 0) {
        // ---- (a) replay: H_k and d_k for every recorded step ----------------------------------------------
        for (int i = 0; i < n; ++i)
            for (int j = lane; j < n; j += 32) M[i * ld + j] = (i == j) ? T(1) : T(0);
        for (int k = 0; k < K; ++k) {
            __syncwarp();
            T* Hk = base + (size_t)k * per_step;
            T* Dk = Hk + (size_t)n * n;
            for (int c = lane; c < n; c += 32) gk[c] = G[(size_t)k * n + c];
            __syncwarp();
            if (k == 0) {
                for (int c = lane; c < n; c += 32) xt[c] = -gk[c];           // bfgs_solver.py:152-155
            } else {
                const T ap = A[k - 1];
                for (int c = lane; c < n; c += 32) {
                    s[c] = ap * dprev[c];
                    y[c] = gk[c] - gprev[c];
                }
                __syncwarp();
                if (k == 1) {                                                  // :159-167, :217-233
                    const T sy = dot(s, y, n, lane), yy = dot(y, y, n, lane);
                    const T den = yy < T(1e-5) ? T(1e-5) : yy;
                    T sc = sy / den;
                    sc = sc < T(1e-4) ? T(1e-4) : sc;
                    for (int c = lane; c < n; c += 32) M[c * ld + c] = sc;
                    __syncwarp();
                }
                T rho, q;
                update_terms(M, ld, T(0), s, y, a, bv, cv, w, rho, q, n, lane);
                const T onepq = T(1) + q;
                for (int i = 0; i < n; ++i) {                                  // :263-303
                    const T ai = a[i], ci = cv[i];
                    for (int j = lane; j < n; j += 32) {
                        const T h = M[i * ld + j] + onepq * ai * s[j] - ai * bv[j] - ci * a[j];
                        M[i * ld + j] = h;
                        Hk[(size_t)i * n + j] = h;
                    }
                }
                mv(M, ld, gk, xt, n, lane);                                    // :173-176
                for (int c = lane; c < n; c += 32) xt[c] = -xt[c];
            }
            __syncwarp();
            for (int c = lane; c < n; c += 32) {
                Dk[c] = xt[c];
                dprev[c] = xt[c];
                gprev[c] = gk[c];
            }
        }
        __syncwarp();
        // ---- (b) reverse sweep ----------------------------------------------------------------------------
        for (int i = 0; i < n; ++i)
            for (int j = lane; j < n; j += 32) M[i * ld + j] = T(0);          // Hbar
        for (int c = lane; c < n; c += 32) sbar_next[c] = gcarry[c] = T(0);
        __syncwarp();
        for (int k = K - 1; k >= 0; --k) {
            const T ak = A[k];
            const T* Hk = base + (size_t)k * per_step;
            for (int c = lane; c < n; c += 32) {
                dbar[c] = ak * (sbar_next[c] + xbar[c]);   // x_{k+1} = x_k + s_k, H_{k+1} = U(H_k, s_k, y_k); s_k = alpha_k d_k
                gbar[c] = gcarry[c];                        // from y_k = g_{k+1} - g_k
                gk[c] = G[(size_t)k * n + c];
            }
            __syncwarp();
            if (k == 0) {
                for (int c = lane; c < n; c += 32) gbar[c] -= dbar[c];         // d_0 = -g_0
            } else {
                mtv(Hk, n, dbar, xt, n, lane);                                 // d_k = -H_k g_k
                for (int c = lane; c < n; c += 32) gbar[c] -= xt[c];
                for (int i = 0; i < n; ++i) {
                    const T di = dbar[i];
                    for (int j = lane; j < n; j += 32) M[i * ld + j] -= di * gk[j];
                }
                const T ap = A[k - 1];
                const T* Dp = base + (size_t)(k - 1) * per_step + (size_t)n * n;
                for (int c = lane; c < n; c += 32) {
                    s[c] = ap * Dp[c];
                    y[c] = gk[c] - G[(size_t)(k - 1) * n + c];
                }
                __syncwarp();
                T sy = T(0), yy = T(0), den = T(1), scale = T(1);
                const T* Hp = nullptr;
                if (k == 1) {
                    sy = dot(s, y, n, lane);
                    yy = dot(y, y, n, lane);
                    den = yy < T(1e-5) ? T(1e-5) : yy;
                    scale = sy / den;
                    scale = scale < T(1e-4) ? T(1e-4) : scale;
                } else {
                    Hp = base + (size_t)(k - 1) * per_step;
                }
                T rho, q;
                update_terms(Hp, n, scale, s, y, a, bv, cv, w, rho, q, n, lane);
                {
                    const T* const rv[3] = {s, bv, a};
                    T* const ro[3] = {Gs, Gb, Ga};
                    mv_multi<T, 3>(M, ld, rv, ro, n, lane);     // Hbar s, Hbar b, Hbar a
                    const T* const cvv[2] = {a, cv};
                    T* const co[2] = {Gta, Gtc};
                    mtv_multi<T, 2>(M, ld, cvv, co, n, lane);   // Hbar^T a, Hbar^T c
                }
                const T qbar = dot(a, Gs, n, lane);
                const T onepq = T(1) + q;
                for (int c = lane; c < n; c += 32) {
                    abar[c] = onepq * Gs[c] - Gb[c] - Gtc[c];
                    sbar_p[c] = onepq * Gta[c];
                    bbar[c] = -Gta[c] + qbar * w[c];
                    cbar[c] = -Ga[c];
                    wbar[c] = qbar * bv[c];
                }
                __syncwarp();
                // ybar = Hp bbar + Hp^T cbar + rho wbar
                if (Hp) {
                    mv_mtv(Hp, n, bbar, cbar, xt, gt, n, lane);   // Hp bbar, Hp^T cbar
                    for (int c = lane; c < n; c += 32) ybar[c] = xt[c] + gt[c] + rho * wbar[c];
                } else {
                    for (int c = lane; c < n; c += 32) ybar[c] = scale * (bbar[c] + cbar[c]) + rho * wbar[c];
                }
                __syncwarp();
                const T rhobar = dot(wbar, y, n, lane) + dot(abar, s, n, lane);
                const T t = -rho * rho * rhobar;                               // func_inverse_curvature.py:22-37
                for (int c = lane; c < n; c += 32) {
                    sbar_p[c] += rho * abar[c] + t * y[c];
                    ybar[c] += t * s[c];
                }
                __syncwarp();
                if (k == 1) {
                    // H'_0 = scale I: scalebar = trace(Hbar + y bbar^T + cbar y^T)
                    T tr = T(0);
                    for (int c = lane; c < n; c += 32) tr += M[c * ld + c] + y[c] * bbar[c] + cbar[c] * y[c];
                    const T scalebar = warp_allreduce(tr);
                    if (sy / den >= T(1e-4)) {                                 // clamp(min=1e-4) passes the gradient
                        const T numbar = scalebar / den;
                        const T denbar = -scalebar * sy / (den * den);
                        const T two_denbar = (yy >= T(1e-5)) ? T(2) * denbar : T(0);  // clamp(min=1e-5) on y.y
                        for (int c = lane; c < n; c += 32) {
                            sbar_p[c] += numbar * y[c];
                            ybar[c] += numbar * s[c] + two_denbar * y[c];
                        }
                    }
                    for (int i = 0; i < n; ++i)
                        for (int j = lane; j < n; j += 32) M[i * ld + j] = T(0);
                } else {
                    for (int i = 0; i < n; ++i) {                              // Hbar <- Hbar + y bbar^T + cbar y^T
                        const T yi = y[i], ci = cbar[i];
                        for (int j = lane; j < n; j += 32) M[i * ld + j] += yi * bbar[j] + ci * y[j];
                    }
                }
                __syncwarp();
                for (int c = lane; c < n; c += 32) {
                    gbar[c] += ybar[c];
                    gcarry[c] = -ybar[c];
                    sbar_next[c] = sbar_p[c];
                }
            }
            __syncwarp();
            // xbar += (d g_k / d x_k)^T gbar: Hessian-vector product at x_k
            for (int c = lane; c < n; c += 32) gk[c] = X[(size_t)k * n + c];   // gk now holds x_k
            for (int c = n + lane; c < wide_vec(n); c += 32) gk[c] = T(0);
            __syncwarp();
            if (dout) DG::analytic(obj, gk, gbar, dout, lane);
            const T nv = sqrt(dot(gbar, gbar, n, lane));
            if (nv > T(0) && isfinite(nv)) {
                const T nx = sqrt(dot(gk, gk, n, lane));
                const T h = bp.rel_step * (T(1) + nx) / nv;
                for (int c = lane; c < n; c += 32) hacc[c] = T(0);
                const T steps[4] = {T(2), T(1), T(-1), T(-2)};
                const T coef[4] = {T(-1), T(8), T(-8), T(1)};
                const T inv = T(1) / (T(12) * h);
#pragma unroll 1
                for (int e = 0; e < 4; ++e) {
                    __syncwarp();
                    for (int c = lane; c < n; c += 32) xt[c] = gk[c] + steps[e] * h * gbar[c];
                    for (int c = n + lane; c < wide_vec(n); c += 32) xt[c] = T(0);
                    __syncwarp();
                    DG::arm(obj, dout, coef[e] * inv);   // the same difference gives d (g . gbar) / d data
                    obj.eval(xt, gt);
                    __syncwarp();
                    for (int c = lane; c < n; c += 32) hacc[c] += coef[e] * gt[c];
                }
                DG::arm(obj, nullptr, T(0));
                __syncwarp();
                for (int c = lane; c < n; c += 32) xbar[c] += hacc[c] * inv;
            }
            __syncwarp();
        }
    }
    for (int c = lane; c < n; c += 32) bp.grad_x0[(size_t)b * n + c] = xbar[c];
    __syncwarp();
}

}  // namespace davo
