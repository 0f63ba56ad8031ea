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

// bfgs_solver.py:80-215 driving wolfe_conditions.py:23-239 for one problem.  NP = compile-time bound on n
// (row length of H held in registers).
//
// The outer loop and the line search are written as ONE loop around ONE objective evaluation: the
// evaluator is the bulk of the kernel's code, and with a separate call site in the outer iteration, the line
// search and the final cost it was inlined three times (57 KB of SASS, beyond the instruction cache).
// `mode` says what the evaluation that is about to run is for:
//   kEvalOuter  (f, grad) at x for the top of an outer iteration          bfgs_solver.py:128-135
//   kEvalProbe  a line-search probe at x + cand * d                       wolfe_conditions.py:134-143
//   kEvalFinal  the cost at the returned parameters                       networks/calibration_network.py:71
template <typename T, int NP, typename Obj>
__device__ __forceinline__ void solve_one_warp(Obj& obj, const SolveParams<T>& p, int b, T* xt_line, T* bc_line,
                                               T* scratch, int lane) {
    enum { kEvalOuter, kEvalProbe, kEvalFinal };
    const int n = p.n;
    const int c = lane >> 1;
    const bool own = c < n;
    T x = own ? p.x0[(size_t)b * n + c] : T(0);
    T g = T(0), gprev = T(0), d = T(0), s = T(0), f = T(0);
    T H[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) H[j] = (j == c) ? T(1) : T(0);  // :112-117
    int iters = 0, fevals = 0, reason = DAVO_REASON_CAP, k = 0;
    // line-search state (wolfe_conditions.py:77-114)
    bool widening = false, zooming = false;
    T lo = T(0), hi = T(0), cand = T(1), lo_f = T(0), hi_f = T(0), cand_f = T(0), f0 = T(0), g0 = T(0),
      neg_c2_g0 = T(0), gt = T(0);
    int ls_i = 0;
    int mode = (p.max_iters > 0) ? kEvalOuter : kEvalFinal;

    for (;;) {
        const T pt = (mode == kEvalProbe) ? add_rn(x, mul_rn(cand, d)) : x;  // wolfe_conditions.py:139
        T fe, ge;
        eval_at(obj, pt, xt_line, lane, fe, ge);
        if (mode == kEvalFinal) {
            f = fe;
            break;
        }
        bool start_iteration = false;
        if (mode == kEvalOuter) {
            f = fe;
            g = ge;
            start_iteration = true;
        } else {
            // ---- one line-search probe has been evaluated: wolfe_conditions.py:143-237 ----
            cand_f = fe;
            gt = ge;
            ++fevals;
            const T dphi = slot_allreduce(mul_rn(d, gt));                  // d/d alpha f(x + alpha d), :141
            bool D = cand_f > add_rn(f0, mul_rn(mul_rn(p.c1, cand), g0)); // :146-150
            if (zooming) D = D || (cand_f >= lo_f);                       // :151-153
            if (widening && ls_i > 0) D = D || (cand_f >= hi_f);          // :154-157
            const bool C = p.strong ? (fabs(dphi) <= neg_c2_g0)           // :160-164
                                    : (mul_rn(T(-1), dphi) <= neg_c2_g0); // :165-169
            const bool G = widening ? (dphi >= T(0)) : (mul_rn(dphi, sub_rn(hi, lo)) >= T(0));  // :174-180
            if (zooming) {                                                // :187-207
                if (D) {
                    hi = cand; hi_f = cand_f;
                } else if (C) {
                    hi = lo = cand; hi_f = lo_f = cand_f; zooming = false;
                } else {
                    if (G) { hi = lo; hi_f = lo_f; }
                    lo = cand; lo_f = cand_f;
                }
            } else {                                                      // widening, :216-237
                if (D) {
                    lo = hi; lo_f = hi_f; hi = cand; hi_f = cand_f; widening = false; zooming = true;
                } else if (C) {
                    hi = lo = cand; hi_f = lo_f = cand_f; widening = false;
                } else if (G) {
                    lo = cand; lo_f = cand_f; widening = false; zooming = true;
                }
            }
            if (zooming && !(lo != hi)) zooming = false;                  // :236
            ++ls_i;
            if ((widening || zooming) && ls_i < p.max_ls) {               // :116-121: another probe
                if (widening) {                                           // :125-127
                    hi = cand; hi_f = cand_f; cand = mul_rn(T(2), cand);
                }
                if (zooming) cand = mul_rn(T(0.5), add_rn(lo, hi));       // :128-131, :242-253
                continue;
            }
            // ---- line search finished, alpha = upper_alpha (:239): take the step, bfgs_solver.py:191-207 ----
            ++iters;
            s = mul_rn(hi, d);                                            // :191
            x = add_rn(x, s);                                             // :192
            const T nrm = sqrt_rn(slot_allreduce(mul_rn(s, s)));          // :203-205
            gprev = g;
            // The accepted point is bitwise x + alpha d: when the search returns the probe it has just made,
            // that probe's (f, grad) ARE the next outer iteration's evaluation.
            const bool reuse = same_bits(hi, cand);
#if DAVO_TRACE
            if (b == p.trace_problem && lane == 0 && k < p.trace_capacity) {
                T* rec = p.trace + 8 * (size_t)k;
                rec[0] = T(k); rec[1] = f0; rec[2] = g0; rec[3] = hi; rec[4] = T(ls_i); rec[5] = nrm; rec[6] = cand_f;
                rec[7] = reuse ? T(1) : T(0);
            }
#endif
            ++k;
            const bool stop_step = !(nrm > p.min_step);                   // :203-207 (strict >)
            if (stop_step || k >= p.max_iters) {                          // :118
                reason = stop_step ? DAVO_REASON_STEP : DAVO_REASON_CAP;
                if (reuse) {
                    f = cand_f;
                    break;
                }
                mode = kEvalFinal;
                continue;
            }
            if (!reuse) {
                mode = kEvalOuter;
                continue;
            }
            f = cand_f;
            g = gt;
            start_iteration = true;
        }
        if (start_iteration) {
            // ---- top of an outer iteration with (f, g) at x: bfgs_solver.py:136-190 ----
            ++fevals;  // the reference evaluates here even when we could reuse the probe
            if (!(f > p.thr)) {                                           // :143 (strict >; NaN retires)
                reason = (f <= p.thr) ? DAVO_REASON_THRESHOLD : DAVO_REASON_NAN;
                break;
            }
            if (k == 0) {
                d = mul_rn(T(-1), g);                                     // :152-155
            } else {
                const T y = sub_rn(g, gprev);                             // :157
                const T sy = slot_allreduce(mul_rn(s, y));                // y^T s
                if (k == 1) {                                             // :159-167, :217-233 (eq. 6.20)
                    T den = slot_allreduce(mul_rn(y, y));
                    den = (den < T(1e-5)) ? T(1e-5) : den;
                    T sc = div_rn(sy, den);
                    sc = (sc < T(1e-4)) ? T(1e-4) : sc;
#pragma unroll
                    for (int j = 0; j < NP; ++j) H[j] = mul_rn(sc, H[j]);
                }
                T rho = div_rn(T(1), sy);                                 // func_inverse_curvature.py:8-11
                if (sy <= T(0)) rho = T(0);
                T gvv[NP];
#if DAVO_FAITHFUL_BFGS
                // Literal restatement: y^T H and H y are formed separately (H is symmetric only up to
                // rounding) and every product is rounded before the next operation, as ATen does.
                T yv[NP], sv[NP], yHv[NP];
                slot_gather<T, NP>(y, bc_line, lane, yv);
                slot_gather<T, NP>(s, bc_line, lane, sv);
                T Hy = T(0);                                              // (H y)_c, :293-295
                T part[kSlots];
#pragma unroll
                for (int j = 0; j < kSlots; ++j) part[j] = T(0);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    Hy = fma_t(H[j], yv[j], Hy);
                    part[j] = mul_rn(y, H[j]);                            // y_c H[c][j]
                }
                const T yH = reduce_scatter16<false>(part, lane);         // (y^T H)_c, :268-270
                slot_gather<T, NP>(yH, bc_line, lane, yHv);
                T q = T(0);                                               // :271-274
#pragma unroll
                for (int j = 0; j < NP; ++j) q = fma_t(yHv[j], mul_rn(yv[j], rho), q);
                const T sr = mul_rn(s, rho);                              // :277
                const T onepq = add_rn(T(1), q);
#pragma unroll
                for (int j = 0; j < NP; ++j) {                            // :278-303, old H on the right
                    const T srj = mul_rn(sv[j], rho);
                    const T sop = mul_rn(mul_rn(sr, sv[j]), onepq);
                    const T sgp = mul_rn(sr, yHv[j]);
                    const T gsp = mul_rn(Hy, srj);
                    H[j] = sub_rn(sub_rn(add_rn(H[j], sop), sgp), gsp);
                }
                slot_gather<T, NP>(g, bc_line, lane, gvv);
#else
                // Same update, H + (s rho) s^T (1+q) - (s rho)(y^T H) - (H y)(s rho)^T with the old H on the
                // right (:263-303), same operands, but the products are allowed to fuse into FMAs.
                // NOTE: y^T H must be formed from the COLUMNS of H.  Replacing it by (H y)^T ("H is symmetric")
                // is wrong in float32: H is symmetric only up to rounding, the true update maps the
                // antisymmetric part A to V^T A V (bounded) while the shortcut adds rho s y^T (H^T - H), which
                // grows on ill-conditioned problems until the directions are useless (config 4: problems that
                // converge in ~250 iterations ran into the 1000-iteration cap with ~39 probes per search).
                // (y^T H)_c = sum_j y_j H[j][c] needs COLUMN c while lane pair c holds ROW c: the rows go through
                // the warp's scratch area (row pitch kRedPitch) and come back as columns, 3 stores + NP loads
                // instead of a 15-shuffle reduce-scatter with two selects per exchanged value.
                T yv[NP], sv[NP];
                slot_gather2<T, NP>(y, s, bc_line, xt_line, lane, yv, sv);
                if (c < NP) {  // every row the column walk below touches is written (rows >= n keep their identity row)
                    using V4 = typename Vec4<T>::type;
                    using V2 = typename Vec2<T>::type;
                    static_assert(NP <= kRedPitch && NP >= 8, "row layout below assumes 8 <= NP <= kRedPitch");
                    T* rowp = scratch + c * kRedPitch;
                    if (!(lane & 1)) {
                        V4 r0, r1;
                        r0.x = H[0]; r0.y = H[1]; r0.z = H[2]; r0.w = H[3];
                        r1.x = H[4]; r1.y = H[5]; r1.z = H[6]; r1.w = H[7];
                        reinterpret_cast<V4*>(rowp)[0] = r0;
                        reinterpret_cast<V4*>(rowp)[1] = r1;
                    } else {
#pragma unroll
                        for (int j = 8; j + 1 < NP; j += 2) {
                            V2 r;
                            r.x = H[j]; r.y = H[j + 1];
                            *reinterpret_cast<V2*>(rowp + j) = r;
                        }
                        if (NP & 1) rowp[NP - 1] = H[NP - 1];
                    }
                }
                __syncwarp();
                T Hy = T(0), yH = T(0);                                   // (H y)_c :293-295, (y^T H)_c :268-270
                {
                    const T* colp = scratch + (own ? c : 0);
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        Hy = fma_t(H[j], yv[j], Hy);
                        yH = fma_t(yv[j], colp[j * kRedPitch], yH);       // y_j H[j][c]; rows j >= n hold y_j = 0
                    }
                    if (!own) yH = T(0);
                }
                const T q = mul_rn(slot_allreduce(yH * y), rho);          // y^T H y / (y^T s), :271-274
                const T onepq = add_rn(T(1), q);
                const T sr = mul_rn(s, rho);                              // :277
                const T nHyrho = -mul_rn(Hy, rho);
                T yHv[NP];
                slot_gather2<T, NP>(yH, g, bc_line, xt_line, lane, yHv, gvv);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const T inner = fma_t(sv[j], onepq, -yHv[j]);         // s_j (1+q) - (y^T H)_j
                    H[j] = fma_t(nHyrho, sv[j], fma_t(sr, inner, H[j]));
                }
#endif
                T Hg = T(0);
#pragma unroll
                for (int j = 0; j < NP; ++j) Hg = fma_t(H[j], gvv[j], Hg);
                d = own ? mul_rn(T(-1), Hg) : T(0);                       // :173-176
            }
            // ---- line-search set-up, wolfe_conditions.py:77-114 ----
            f0 = f;
            g0 = slot_allreduce(mul_rn(d, g));                            // :77
            neg_c2_g0 = mul_rn(T(-1) * p.c2, g0);                         // -1.0 * curvature * base_gradient
            widening = true; zooming = false;                             // :80-82
            lo = T(0); hi = T(0); cand = T(1);                            // :97-108
            lo_f = f0; hi_f = f0; cand_f = f0;                            // :109-111
            ls_i = 0;
            mode = kEvalProbe;
        }
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
