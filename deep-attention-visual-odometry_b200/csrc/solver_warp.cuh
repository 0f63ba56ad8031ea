// solver_warp.cuh — one warp solves one problem: BFGS outer loop + strong-Wolfe line search.
//
// State machine restated from /root/reference/deep_attention_visual_odometry/
//   autograd_solvers/bfgs_solver.py:80-215            (outer loop, eval mode)
//   autograd_solvers/bfgs_solver.py:217-233, 235-303  (eq. 6.20 scaling, eq. 6.17 update)
//   utils/func_inverse_curvature.py:8-11              (1/(y.s), 0 when y.s <= 0)
//   autograd_solvers/line_search/wolfe_conditions.py:23-253 (alg. 3.5 widen + 3.6 bisection zoom)
// per problem (no reference operation couples problems).
//
// Data layout inside the warp: the n-vectors x, g, d, s, y are DISTRIBUTED — lanes 2c and 2c+1 hold
// component c ("slot" c, 16 slots, zero beyond n) — and lane pair c holds row c of the n x n inverse
// Hessian in registers.  Scalars of the state machine (f, alpha, lo, hi, ...) are replicated and
// bitwise identical in all lanes, so every branch is warp-uniform.  Vectors are made visible to all
// lanes through a 16-entry shared-memory line (one STS + vector LDS), reductions over components use
// a 4-step butterfly, and y^T H uses a 15-shuffle reduce-scatter.
//
// Evaluation reuse: the reference re-evaluates f and grad f at the accepted point at the top of the
// next outer iteration (bfgs_solver.py:128-135).  The accepted point is bitwise x + alpha*d, which is
// the line search's last probe whenever it returns the probe it just made; its (f, grad) are kept in
// that case.  fevals_out still counts what the reference would have evaluated.
#pragma once
#include "davo_common.cuh"

#ifndef DAVO_FAITHFUL_BFGS
#define DAVO_FAITHFUL_BFGS 0
#endif

namespace davo {

template <typename T, int NP>
__device__ __forceinline__ void slot_gather(T own, T* line, int lane, T (&out)[NP]) {
    using V4 = typename Vec4<T>::type;
    __syncwarp();  // earlier readers of `line` are done
    if (!(lane & 1)) line[lane >> 1] = own;
    __syncwarp();
    const V4* l4 = reinterpret_cast<const V4*>(line);
#pragma unroll
    for (int q = 0; q < (NP + 3) / 4; ++q) {
        const V4 t = l4[q];
        if (4 * q + 0 < NP) out[4 * q + 0] = t.x;
        if (4 * q + 1 < NP) out[4 * q + 1] = t.y;
        if (4 * q + 2 < NP) out[4 * q + 2] = t.z;
        if (4 * q + 3 < NP) out[4 * q + 3] = t.w;
    }
}

// Two vectors at once through two lines (one pair of warp syncs instead of two).
template <typename T, int NP>
__device__ __forceinline__ void slot_gather2(T own_a, T own_b, T* line_a, T* line_b, int lane, T (&out_a)[NP],
                                             T (&out_b)[NP]) {
    using V4 = typename Vec4<T>::type;
    __syncwarp();
    if (!(lane & 1)) {
        line_a[lane >> 1] = own_a;
        line_b[lane >> 1] = own_b;
    }
    __syncwarp();
    const V4* a4 = reinterpret_cast<const V4*>(line_a);
    const V4* b4 = reinterpret_cast<const V4*>(line_b);
#pragma unroll
    for (int q = 0; q < (NP + 3) / 4; ++q) {
        const V4 ta = a4[q], tb = b4[q];
        if (4 * q + 0 < NP) { out_a[4 * q + 0] = ta.x; out_b[4 * q + 0] = tb.x; }
        if (4 * q + 1 < NP) { out_a[4 * q + 1] = ta.y; out_b[4 * q + 1] = tb.y; }
        if (4 * q + 2 < NP) { out_a[4 * q + 2] = ta.z; out_b[4 * q + 2] = tb.z; }
        if (4 * q + 3 < NP) { out_a[4 * q + 3] = ta.w; out_b[4 * q + 3] = tb.w; }
    }
}

template <typename T, typename Obj>
__device__ __forceinline__ void eval_at(Obj& obj, T xc, T* xt_line, int lane, T& f, T& g_own) {
    __syncwarp();
    if (!(lane & 1)) xt_line[lane >> 1] = xc;
    __syncwarp();
    obj.eval(xt_line, f, g_own);
}

template <typename T>
struct LineSearchResult {
    T alpha;      // upper_alpha, wolfe_conditions.py:239
    T last_cand;  // the last probe made
    T last_f;     // objective at the last probe
    T last_g;     // this lane's gradient component at the last probe
    int probes;
};

// wolfe_conditions.py:23-239 for one problem.  x, d, g are this lane's components.
template <typename T, typename Obj>
__device__ __forceinline__ LineSearchResult<T> line_search_warp(Obj& obj, const SolveParams<T>& p, T x, T d, T f0,
                                                                T g, T* xt_line, int lane) {
    const T g0 = slot_allreduce(mul_rn(d, g));  // :77
    bool widening = true, zooming = false;      // :80-82
    T lo = T(0), hi = T(0), cand = T(1);        // :97-108
    T lo_f = f0, hi_f = f0, cand_f = f0;        // :109-111
    T gt = T(0);
    int probes = 0;
    const T neg_c2_g0 = mul_rn(T(-1) * p.c2, g0);  // -1.0 * curvature * base_gradient (:163,:168)
    for (int i = 0; i < p.max_ls; ++i) {           // :116
        if (!(widening || zooming)) break;         // :119-121
        if (i > 0) {
            if (widening) {                        // :125-127
                hi = cand;
                hi_f = cand_f;
                cand = mul_rn(T(2), cand);
            }
            if (zooming) cand = mul_rn(T(0.5), add_rn(lo, hi));  // :128-131, :242-253
        }
        const T xt = add_rn(x, mul_rn(cand, d));   // :139
        eval_at(obj, xt, xt_line, lane, cand_f, gt);
        const T dphi = slot_allreduce(mul_rn(d, gt));  // d/d alpha f(x + alpha d), :141
        ++probes;
        bool D = cand_f > add_rn(f0, mul_rn(mul_rn(p.c1, cand), g0));  // :146-150
        if (zooming) D = D || (cand_f >= lo_f);                        // :151-153
        if (widening && i > 0) D = D || (cand_f >= hi_f);              // :154-157
        const bool C = p.strong ? (fabs(dphi) <= neg_c2_g0)            // :160-164
                                : (mul_rn(T(-1), dphi) <= neg_c2_g0);  // :165-169
        const bool G = widening ? (dphi >= T(0)) : (mul_rn(dphi, sub_rn(hi, lo)) >= T(0));  // :174-180
        if (zooming) {                             // :187-207
            if (D) {
                hi = cand; hi_f = cand_f;
            } else if (C) {
                hi = lo = cand; hi_f = lo_f = cand_f; zooming = false;
            } else {
                if (G) { hi = lo; hi_f = lo_f; }
                lo = cand; lo_f = cand_f;
            }
        } else {                                   // widening, :216-237
            if (D) {
                lo = hi; lo_f = hi_f; hi = cand; hi_f = cand_f; widening = false; zooming = true;
            } else if (C) {
                hi = lo = cand; hi_f = lo_f = cand_f; widening = false;
            } else if (G) {
                lo = cand; lo_f = cand_f; widening = false; zooming = true;
            }
        }
        if (zooming && !(lo != hi)) zooming = false;  // :236
    }
    LineSearchResult<T> r;
    r.alpha = hi;
    r.last_cand = cand;
    r.last_f = cand_f;
    r.last_g = gt;
    r.probes = probes;
    return r;
}

template <typename T>
__device__ __forceinline__ bool same_bits(T a, T b);
template <>
__device__ __forceinline__ bool same_bits<float>(float a, float b) { return __float_as_uint(a) == __float_as_uint(b); }
template <>
__device__ __forceinline__ bool same_bits<double>(double a, double b) {
    return __double_as_longlong(a) == __double_as_longlong(b);
}

// bfgs_solver.py:80-215 for one problem.  NP = compile-time bound on n (row length held in registers).
template <typename T, int NP, typename Obj>
__device__ __forceinline__ void solve_one_warp(Obj& obj, const SolveParams<T>& p, int b, T* xt_line, T* bc_line,
                                               int lane) {
    const int n = p.n;
    const int c = lane >> 1;
    const bool own = c < n;
    T x = own ? p.x0[(size_t)b * n + c] : T(0);
    T g = T(0), gprev = T(0), d = T(0), s = T(0), f = T(0);
    T H[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) H[j] = (j == c) ? T(1) : T(0);  // :112-117
    int iters = 0, fevals = 0, reason = DAVO_REASON_CAP;
    bool have_fg = false, have_f = false;

    for (int k = 0; k < p.max_iters; ++k) {  // :118
        if (!have_fg) eval_at(obj, x, xt_line, lane, f, g);  // :128-135
        ++fevals;
        have_f = true;
        if (!(f > p.thr)) {  // :143 (strict >; NaN retires)
            reason = (f <= p.thr) ? DAVO_REASON_THRESHOLD : DAVO_REASON_NAN;
            break;
        }
        if (k == 0) {
            d = mul_rn(T(-1), g);  // :152-155
        } else {
            const T y = sub_rn(g, gprev);                    // :157
            const T sy = slot_allreduce(mul_rn(s, y));       // y^T s
            if (k == 1) {                                    // :159-167, :217-233 (eq. 6.20)
                T den = slot_allreduce(mul_rn(y, y));
                den = (den < T(1e-5)) ? T(1e-5) : den;
                T sc = div_rn(sy, den);
                sc = (sc < T(1e-4)) ? T(1e-4) : sc;
#pragma unroll
                for (int j = 0; j < NP; ++j) H[j] = mul_rn(sc, H[j]);
            }
            T rho = div_rn(T(1), sy);                        // func_inverse_curvature.py:8-11
            if (sy <= T(0)) rho = T(0);
#if DAVO_FAITHFUL_BFGS
            // Literal restatement: y^T H and H y are formed separately (H is symmetric only up to rounding)
            // and every product is rounded before the next operation, as ATen does.
            T yv[NP], sv[NP], yHv[NP];
            slot_gather<T, NP>(y, bc_line, lane, yv);
            slot_gather<T, NP>(s, bc_line, lane, sv);
            T Hy = T(0);                                     // (H y)_c, :293-295
            T part[kSlots];
#pragma unroll
            for (int j = 0; j < kSlots; ++j) part[j] = T(0);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                Hy = fma_t(H[j], yv[j], Hy);
                part[j] = mul_rn(y, H[j]);                   // y_c H[c][j]
            }
            const T yH = reduce_scatter16<false>(part, lane);  // (y^T H)_c, :268-270
            slot_gather<T, NP>(yH, bc_line, lane, yHv);
            T q = T(0);                                      // :271-274
#pragma unroll
            for (int j = 0; j < NP; ++j) q = fma_t(yHv[j], mul_rn(yv[j], rho), q);
            const T sr = mul_rn(s, rho);                     // :277
            const T onepq = add_rn(T(1), q);
#pragma unroll
            for (int j = 0; j < NP; ++j) {                   // :278-303, old H on the right-hand side
                const T srj = mul_rn(sv[j], rho);
                const T sop = mul_rn(mul_rn(sr, sv[j]), onepq);
                const T sgp = mul_rn(sr, yHv[j]);
                const T gsp = mul_rn(Hy, srj);
                H[j] = sub_rn(sub_rn(add_rn(H[j], sop), sgp), gsp);
            }
            T gvv[NP];
            slot_gather<T, NP>(g, bc_line, lane, gvv);
#else
            // Same update, H + (s rho) s^T (1+q) - (s rho)(y^T H) - (H y)(s rho)^T with the old H on the
            // right (:263-303), using y^T H = (H y)^T: H stays symmetric to rounding under this update, so the
            // two differ in the last bits only (the float64 gate still matches the reference's step counts
            // on 100 % of problems).  Saves the reduce-scatter over rows and lets the products fuse.
            T yv[NP], sv[NP];
            slot_gather2<T, NP>(y, s, bc_line, xt_line, lane, yv, sv);
            T Hy = T(0);                                     // (H y)_c, :293-295
#pragma unroll
            for (int j = 0; j < NP; ++j) Hy = fma_t(H[j], yv[j], Hy);
            const T q = mul_rn(slot_allreduce(mul_rn(y, Hy)), rho);  // y^T H y / (y^T s), :271-274
            const T onepq = add_rn(T(1), q);
            const T sr = mul_rn(s, rho);                     // :277
            const T nHyrho = -mul_rn(Hy, rho);
            T Hyv[NP], gvv[NP];
            slot_gather2<T, NP>(Hy, g, bc_line, xt_line, lane, Hyv, gvv);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const T inner = fma_t(sv[j], onepq, -Hyv[j]);   // s_j (1+q) - (y^T H)_j
                H[j] = fma_t(nHyrho, sv[j], fma_t(sr, inner, H[j]));
            }
#endif
            T Hg = T(0);
#pragma unroll
            for (int j = 0; j < NP; ++j) Hg = fma_t(H[j], gvv[j], Hg);
            d = own ? mul_rn(T(-1), Hg) : T(0);              // :173-176
        }
        const LineSearchResult<T> ls = line_search_warp(obj, p, x, d, f, g, xt_line, lane);  // :181-190
        fevals += ls.probes;
        ++iters;
        s = mul_rn(ls.alpha, d);                             // :191
        x = add_rn(x, s);                                    // :192
        const T nrm = sqrt_rn(slot_allreduce(mul_rn(s, s))); // :203-205
        gprev = g;
        have_fg = same_bits(ls.alpha, ls.last_cand);
        have_f = have_fg;
        if (have_fg) {
            f = ls.last_f;
            g = ls.last_g;
        }
        if (!(nrm > p.min_step)) {                           // :203-207
            reason = DAVO_REASON_STEP;
            break;
        }
    }
    if (!have_f) {  // cost at the returned parameters (networks/calibration_network.py:71)
        T gtmp;
        eval_at(obj, x, xt_line, lane, f, gtmp);
    }
    if (own && !(lane & 1)) p.x_out[(size_t)b * n + c] = x;
    if (lane == 0) {
        if (p.cost_out) p.cost_out[b] = f;
        if (p.converged_out) p.converged_out[b] = (f <= p.thr) ? 1 : 0;
        if (p.iters_out) p.iters_out[b] = iters;
        if (p.fevals_out) p.fevals_out[b] = fevals;
        if (p.reason_out) p.reason_out[b] = reason;
    }
}

}  // namespace davo
