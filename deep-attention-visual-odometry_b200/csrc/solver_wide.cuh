// solver_wide.cuh — one warp solves one problem with up to 64 parameters (the JOINT model: 10
// intrinsics + 6 pose parameters per view).  Same state machine as solver_warp.cuh, restated from
//   autograd_solvers/bfgs_solver.py:80-303, utils/func_inverse_curvature.py:8-11,
//   autograd_solvers/line_search/wolfe_conditions.py:23-253,
// but with the n-vectors and the n x n inverse Hessian in this warp's SHARED memory (n = 34 needs
// 4.6 KB for H; registers could not hold two rows per lane next to the match arithmetic).  Lane L
// works on components L and L + 32.  H rows are padded to an odd stride (n | 1 words), so a lane walking
// its row and 32 lanes reading one column are both bank-conflict free.  Scalars are replicated and bitwise
// identical across lanes; reductions over components are per-lane partial sums + a 5-step butterfly.
#pragma once
#include "davo_common.cuh"
#include "solver_warp.cuh"  // same_bits
#include "objectives.cuh"   // pfma / pk (packed pairs)
#include "train_params.cuh"

namespace davo {

constexpr int kWideMax = 128;  // lane L works on components L, L + 32, ... (kCols = 2 for n <= 64, 4 for n <= 128)

__host__ __device__ constexpr int wide_cols(int n) { return n <= 64 ? 2 : 4; }
// length of one n-vector in the per-warp workspace (>= 64: the sweeps park unused lanes on scratch words of xt)
__host__ __device__ constexpr int wide_vec(int n) { return n <= 64 ? 64 : 128; }

// Row stride of the inverse Hessian.  Run-time n: odd (a lane walking its row and 32 lanes reading one column are both
// bank-conflict free).  Compile-time n: the smallest stride >= n that is 4 (mod 8): rows are 16-byte aligned, so a lane
// reads and writes its row four values at a time, and 8 consecutive rows start in 8 different 16-byte bank groups
// (conflict-free 128-bit accesses); a column read is still one word per lane in consecutive banks.
__host__ __device__ constexpr int wide_ld_aligned(int n) { return ((n + 3) & ~7) + 4 >= n ? ((n + 3) & ~7) + 4 : ((n + 3) & ~7) + 12; }
__host__ __device__ constexpr int wide_ld_max(int n) { return wide_ld_aligned(n) > (n | 1) ? wide_ld_aligned(n) : (n | 1); }

#ifndef DAVO_WIDE_FAITHFUL
#define DAVO_WIDE_FAITHFUL 0  // 1: op-by-op rounding of the BFGS update (A/B builds)
#endif

template <typename T>
__device__ __forceinline__ T warp_allreduce(T v) {
    v += shfl_xor(v, 1);
    v += shfl_xor(v, 2);
    v += shfl_xor(v, 4);
    v += shfl_xor(v, 8);
    v += shfl_xor(v, 16);
    return v;
}

// Per-warp shared-memory workspace of the wide solver.
template <typename T>
struct WideWorkspace {
    T *x, *g, *gprev, *d, *s, *y, *yH, *Hy, *xt, *gt, *H;
    int ld;  // row stride of H
    __host__ __device__ static size_t bytes(int n) {
        return sizeof(T) * (10 * (size_t)wide_vec(n) + (size_t)n * wide_ld_max(n));
    }
    __device__ void carve(unsigned char* base, int n) {
        T* p = reinterpret_cast<T*>(base);
        const int v = wide_vec(n);
        x = p; g = x + v; gprev = g + v; d = gprev + v; s = d + v;
        y = s + v; yH = y + v; Hy = yH + v; xt = Hy + v; gt = xt + v;
        H = gt + v;
        ld = n | 1;  // odd row stride: a lane walking its row and 32 lanes reading one column are both conflict free
    }
};

// a . b over the n <= 64 components: lane L contributes components L and L + 32 (no loop: the kernel is bound
// by instruction issue, and a loop with a run-time trip count costs more than the two guarded products).
template <int kCols = 2, typename T>
__device__ __forceinline__ T wide_dot(const T* a, const T* b, int n, int lane) {
    T acc = (lane < n) ? mul_rn(a[lane], b[lane]) : T(0);
#pragma unroll
    for (int h = 1; h < kCols; ++h)
        if (lane + 32 * h < n) acc = add_rn(acc, mul_rn(a[lane + 32 * h], b[lane + 32 * h]));
    return warp_allreduce(acc);
}

// The zoom step.  Bisection (wolfe_conditions.py:128-131, :242-253) or, with `secant`, the older solvers/ generation's
// interpolate_alpha(lo, hi, phi'(lo), phi'(hi)) (solvers/line_search_strong_wolfe_conditions.py:147-155,
// utils/func_interpolate_alpha.py:15-33); a non-finite interpolant (phi' = NaN or inf at a bracket end) bisects.
template <typename T>
__device__ __forceinline__ T zoom_step(bool secant, T lo, T hi, T lo_d, T hi_d) {
    const T mid = mul_rn(T(0.5), add_rn(lo, hi));
    if (!secant) return mid;
    const T c = interpolate_alpha_value(lo, hi, lo_d, hi_d);
    return isfinite(c) ? c : mid;
}

// wolfe_conditions.py:23-239.  x, d, g live in shared memory; probes are evaluated at ws.xt with the
// gradient written to ws.gt.
template <int kCols = 2, typename T, typename Obj>
__device__ __forceinline__ LineSearchResult<T> line_search_wide(Obj& obj, const SolveParams<T>& p, const T* x,
                                                                const T* d, T f0, const T* g, T* xt, T* gt,
                                                                int lane) {
    const int n = Obj::kParams > 0 ? Obj::kParams : p.n;
    const T g0 = wide_dot<kCols>(d, g, n, lane);       // :77
    bool widening = true, zooming = false;      // :80-82
    T lo = T(0), hi = T(0), cand = T(1);        // :97-108
    T lo_f = f0, hi_f = f0, cand_f = f0;        // :109-111
    T lo_d = g0, hi_d = g0, cand_d = g0;        // secant zoom: phi' at the bracket ends and at the last probe
    int probes = 0;
    const T neg_c2_g0 = mul_rn(T(-1) * p.c2, g0);
    for (int i = 0; i < p.max_ls; ++i) {        // :116
        if (!(widening || zooming)) break;      // :119-121
        if (i > 0) {
            if (widening) { hi = cand; hi_f = cand_f; hi_d = cand_d; cand = mul_rn(T(2), cand); }  // :125-127
            if (zooming) cand = zoom_step(p.zoom != 0, lo, hi, lo_d, hi_d);           // :128-131
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < kCols; ++h)                                              // :139
            if (lane + 32 * h < n) xt[lane + 32 * h] = add_rn(x[lane + 32 * h], mul_rn(cand, d[lane + 32 * h]));
        __syncwarp();
        cand_f = obj.eval(xt, gt);
        const T dphi = wide_dot<kCols>(d, gt, n, lane);                                    // :141
        cand_d = dphi;
        ++probes;
        bool D = cand_f > add_rn(f0, mul_rn(mul_rn(p.c1, cand), g0));               // :146-150
        if (zooming) D = D || (cand_f >= lo_f);                                     // :151-153
        if (widening && i > 0) D = D || (cand_f >= hi_f);                           // :154-157
        const bool C = p.strong ? (fabs(dphi) <= neg_c2_g0) : (mul_rn(T(-1), dphi) <= neg_c2_g0);  // :160-169
        const bool G = widening ? (dphi >= T(0)) : (mul_rn(dphi, sub_rn(hi, lo)) >= T(0));         // :174-180
        if (zooming) {                                                              // :187-207
            if (D) { hi = cand; hi_f = cand_f; hi_d = dphi; }
            else if (C) { hi = lo = cand; hi_f = lo_f = cand_f; hi_d = lo_d = dphi; zooming = false; }
            else { if (G) { hi = lo; hi_f = lo_f; hi_d = lo_d; } lo = cand; lo_f = cand_f; lo_d = dphi; }
        } else {                                                                    // :216-237
            if (D) { lo = hi; lo_f = hi_f; lo_d = hi_d; hi = cand; hi_f = cand_f; hi_d = dphi; widening = false; zooming = true; }
            else if (C) { hi = lo = cand; hi_f = lo_f = cand_f; hi_d = lo_d = dphi; widening = false; }
            else if (G) { lo = cand; lo_f = cand_f; lo_d = dphi; widening = false; zooming = true; }
        }
        if (zooming && !(lo != hi)) zooming = false;                                // :236
    }
    LineSearchResult<T> r;
    r.alpha = hi; r.last_cand = cand; r.last_f = cand_f; r.last_g = T(0); r.probes = probes;
    return r;
}

// bfgs_solver.py:80-215 for one problem.
template <int kCols = 2, typename T, typename Obj, typename Rec = NoRecorder>
__device__ __forceinline__ void solve_one_wide(Obj& obj, const SolveParams<T>& p, int b, WideWorkspace<T>& ws,
                                               int lane, const Rec rec = Rec()) {
    // float32 with a compile-time n: 16-byte-aligned rows and the vectorised sweeps below (float64 gains nothing from
    // them — its packed operations are pairs of scalar ones and a row chunk is two 128-bit loads: 460 ms against 447)
    constexpr bool kVectorRows = Obj::kParams > 0 && sizeof(T) == 4;
    // an objective with a compile-time parameter count (Obj::kParams > 0) turns n and the row stride into
    // constants: the sweeps over H below unroll with immediate offsets
    const int n = Obj::kParams > 0 ? Obj::kParams : p.n, ld = kVectorRows ? wide_ld_aligned(Obj::kParams) : (Obj::kParams > 0 ? (Obj::kParams | 1) : ws.ld);
    T *x = ws.x, *g = ws.g, *gprev = ws.gprev, *gt = ws.gt;
    T *d = ws.d, *s = ws.s, *y = ws.y, *yH = ws.yH, *Hy = ws.Hy, *H = ws.H;
    __syncwarp();
    for (int c = lane; c < n; c += 32) {
        x[c] = p.x0[(size_t)b * n + c];
        s[c] = T(0);
        g[c] = T(0);
    }
    for (int i = 0; i < n; ++i)
        for (int j = lane; j < n; j += 32) H[i * ld + j] = (i == j) ? T(1) : T(0);  // :112-117
    __syncwarp();
    // The outer loop and the line search are ONE loop around ONE objective evaluation (as in solver_warp.cuh):
    // the evaluator is most of this kernel's code and with call sites in the outer iteration, the line search and
    // the final cost it was inlined three times (77 KB of SASS).  `mode` says what the evaluation about to run is for.
    enum { kEvalOuter, kEvalProbe, kEvalFinal };
    T f = T(0);
    int iters = 0, fevals = 0, reason = DAVO_REASON_CAP, k = 0;
    bool widening = false, zooming = false;                                    // wolfe_conditions.py:77-114
    T lo = T(0), hi = T(0), cand = T(1), lo_f = T(0), hi_f = T(0), cand_f = T(0), f0 = T(0), g0 = T(0),
      neg_c2_g0 = T(0);
    T lo_d = T(0), hi_d = T(0), cand_d = T(0);  // secant zoom only (dead code in the eval-mode instantiation)
    const bool secant = Rec::kActive && rec.secant();
    int ls_i = 0;
    int mode = (p.max_iters > 0) ? kEvalOuter : kEvalFinal;
    if (Rec::kActive && mode == kEvalOuter && rec.drop(b, 0)) {  // dropped before its first evaluation (:122-125)
        reason = DAVO_REASON_DROPPED;
        mode = kEvalFinal;
    }

    for (;;) {
        const T* pt = x;
        T* gdst = (mode == kEvalOuter) ? g : gt;
        if (mode == kEvalProbe) {
            __syncwarp();
#pragma unroll
            for (int h = 0; h < kCols; ++h)                                      // wolfe_conditions.py:139
                if (lane + 32 * h < n) ws.xt[lane + 32 * h] = add_rn(x[lane + 32 * h], mul_rn(cand, d[lane + 32 * h]));
            pt = ws.xt;
        }
        __syncwarp();
        const T fe = obj.eval(pt, gdst);
        if (mode == kEvalFinal) {                                              // networks/calibration_network.py:71
            f = fe;
            break;
        }
        bool start_iteration = false;
        if (mode == kEvalOuter) {                                              // bfgs_solver.py:128-135
            f = fe;
            start_iteration = true;
        } else {
            // ---- one line-search probe has been evaluated: wolfe_conditions.py:143-237 ----
            cand_f = fe;
            ++fevals;
            const T dphi = wide_dot<kCols>(d, gt, n, lane);                            // :141
            bool D = cand_f > add_rn(f0, mul_rn(mul_rn(p.c1, cand), g0));       // :146-150
            if (zooming) D = D || (cand_f >= lo_f);                             // :151-153
            if (widening && ls_i > 0) D = D || (cand_f >= hi_f);                // :154-157
            const bool C = p.strong ? (fabs(dphi) <= neg_c2_g0) : (mul_rn(T(-1), dphi) <= neg_c2_g0);  // :160-169
            const bool G = widening ? (dphi >= T(0)) : (mul_rn(dphi, sub_rn(hi, lo)) >= T(0));         // :174-180
            if (Rec::kActive) {  // phi' at the bracket ends follows every assignment of (lo, lo_f) / (hi, hi_f) below
                cand_d = dphi;
                if (zooming) {
                    if (D) { hi_d = dphi; }
                    else if (C) { hi_d = lo_d = dphi; }
                    else { if (G) hi_d = lo_d; lo_d = dphi; }
                } else {
                    if (D) { lo_d = hi_d; hi_d = dphi; }
                    else if (C) { hi_d = lo_d = dphi; }
                    else if (G) { lo_d = dphi; }
                }
            }
            if (zooming) {                                                      // :187-207
                if (D) { hi = cand; hi_f = cand_f; }
                else if (C) { hi = lo = cand; hi_f = lo_f = cand_f; zooming = false; }
                else { if (G) { hi = lo; hi_f = lo_f; } lo = cand; lo_f = cand_f; }
            } else {                                                            // :216-237
                if (D) { lo = hi; lo_f = hi_f; hi = cand; hi_f = cand_f; widening = false; zooming = true; }
                else if (C) { hi = lo = cand; hi_f = lo_f = cand_f; widening = false; }
                else if (G) { lo = cand; lo_f = cand_f; widening = false; zooming = true; }
            }
            if (zooming && !(lo != hi)) zooming = false;                        // :236
            ++ls_i;
            if ((widening || zooming) && ls_i < p.max_ls) {                     // :116-121: another probe
                if (widening) { hi = cand; hi_f = cand_f; if (Rec::kActive) hi_d = cand_d; cand = mul_rn(T(2), cand); }  // :125-127
                if (zooming) cand = zoom_step(secant, lo, hi, lo_d, hi_d);               // :128-131
                continue;
            }
            // ---- line search finished, alpha = upper_alpha (:239): take the step, bfgs_solver.py:191-207 ----
            ++iters;
            T nrm2 = T(0);
            if (!Rec::kActive) {
#pragma unroll
                for (int h = 0; h < kCols; ++h) {                               // :191-199
                    const int c = lane + 32 * h;
                    if (c < n) {
                        const T sc = mul_rn(hi, d[c]);
                        s[c] = sc;
                        x[c] = add_rn(x[c], sc);
                        nrm2 = add_rn(nrm2, mul_rn(sc, sc));
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < kCols; ++h) {
                    const int c = lane + 32 * h;
                    if (c < n) {
                        const T sc = mul_rn(hi, d[c]);
                        s[c] = sc;
                        nrm2 = add_rn(nrm2, mul_rn(sc, sc));
                    }
                }
            }
            const T nrm = sqrt_rn(warp_allreduce(nrm2));
            if (Rec::kActive) {
                // training mode: keep (x_k, g_k, alpha_k) for the backward pass; with return_second_last the step
                // that retires the problem on its length is not applied (:196-212)
                const bool withheld = rec.second_last() && !(nrm > p.min_step);
                __syncwarp();
                rec.record(b, iters - 1, n, x, g, withheld ? T(0) : hi, lane);
                if (withheld) {
                    reason = DAVO_REASON_STEP;
                    f = f0;  // the cost at the retained iterate x_k
                    break;
                }
#pragma unroll
                for (int h = 0; h < kCols; ++h)
                    if (lane + 32 * h < n) x[lane + 32 * h] = add_rn(x[lane + 32 * h], s[lane + 32 * h]);
            }
            // the accepted point is bitwise the last probe whenever the search returns the probe it just made:
            // that probe's (f, grad) are the next outer iteration's evaluation
            const bool reuse = same_bits(hi, cand);
            {   // rotate gradient buffers: gprev <- g, and g <- gt when the accepted point is the last probe
                T* old = gprev;
                gprev = g;
                if (reuse) { g = gt; gt = old; }
                else { g = old; }
            }
            ++k;
            __syncwarp();
            const bool stop_step = !(nrm > p.min_step);                         // :203-207 (strict >)
            if (stop_step || k >= p.max_iters) {                                // :118
                reason = stop_step ? DAVO_REASON_STEP : DAVO_REASON_CAP;
                if (reuse) {
                    f = cand_f;
                    break;
                }
                mode = kEvalFinal;
                continue;
            }
            if (Rec::kActive && rec.drop(b, k)) {                                 // :122-125, top of iteration k
                reason = DAVO_REASON_DROPPED;
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
            start_iteration = true;
        }
        if (start_iteration) {
            // ---- top of an outer iteration with (f, g) at x: bfgs_solver.py:136-190 ----
            ++fevals;  // the reference evaluates here even when the probe is reused
            if (!(f > p.thr)) {                                                 // :143 (strict >; NaN retires)
                reason = (f <= p.thr) ? DAVO_REASON_THRESHOLD : DAVO_REASON_NAN;
                break;
            }
            __syncwarp();
            if (k == 0) {
#pragma unroll
                for (int h = 0; h < kCols; ++h)                                  // :152-155
                    if (lane + 32 * h < n) d[lane + 32 * h] = mul_rn(T(-1), g[lane + 32 * h]);
            } else {
#pragma unroll
                for (int h = 0; h < kCols; ++h)                                     // :157
                    if (lane + 32 * h < n) y[lane + 32 * h] = sub_rn(g[lane + 32 * h], gprev[lane + 32 * h]);
                __syncwarp();
                const T sy = wide_dot<kCols>(s, y, n, lane);
                if (k == 1) {                                                      // :159-167, :217-233
                    T den = wide_dot<kCols>(y, y, n, lane);
                    den = (den < T(1e-5)) ? T(1e-5) : den;
                    T sc = div_rn(sy, den);
                    sc = (sc < T(1e-4)) ? T(1e-4) : sc;
                    for (int i = 0; i < n; ++i)
                        for (int j = lane; j < n; j += 32) H[i * ld + j] = mul_rn(sc, H[i * ld + j]);
                    __syncwarp();
                }
                T rho = div_rn(T(1), sy);                                          // func_inverse_curvature.py:8-11
                if (sy <= T(0)) rho = T(0);
#if DAVO_WIDE_FAITHFUL
                // Literal restatement: every product rounded before the next operation, one element at a time.
                for (int c = lane; c < n; c += 32) {
                    T a = T(0), bsum = T(0);
                    for (int i = 0; i < n; ++i) {
                        a = add_rn(a, mul_rn(y[i], H[i * ld + c]));                // (y^T H)_c, :268-270
                        bsum = add_rn(bsum, mul_rn(H[c * ld + i], y[i]));          // (H y)_c,   :293-295
                    }
                    yH[c] = a;
                    Hy[c] = bsum;
                }
                __syncwarp();
                T qp = T(0);                                                       // :271-274
                for (int c = lane; c < n; c += 32) qp = add_rn(qp, mul_rn(yH[c], mul_rn(y[c], rho)));
                const T onepq = add_rn(T(1), warp_allreduce(qp));
                for (int i = 0; i < n; ++i) {                                      // :278-303
                    const T sri = mul_rn(s[i], rho), Hyi = Hy[i];
                    for (int j = lane; j < n; j += 32) {
                        const T sop = mul_rn(mul_rn(sri, s[j]), onepq);
                        const T sgp = mul_rn(sri, yH[j]);
                        const T gsp = mul_rn(Hyi, mul_rn(s[j], rho));
                        H[i * ld + j] = sub_rn(sub_rn(add_rn(H[i * ld + j], sop), sgp), gsp);
                    }
                }
                __syncwarp();
                for (int c = lane; c < n; c += 32) {                               // :173-176
                    T a = T(0);
                    for (int j = 0; j < n; ++j) a = fma_t(H[c * ld + j], g[j], a);
                    d[c] = mul_rn(T(-1), a);
                }
#else
                // H' = H + (s rho) s^T (1+q) - (s rho)(y^T H) - (H y)(s rho)^T with the old H on the right (:263-303),
                // same operands as the literal form, products allowed to fuse.  Lane L owns columns/rows L and L+32.
                // Sweep 1 walks H once and forms, per owned index c, (y^T H)_c from COLUMN c (never replaced by
                // (H y)^T: H is symmetric only up to rounding, see solver_warp.cuh), (H y)_c and (H g)_c from ROW c:
                // six independent FMA chains per lane, no cross-lane traffic.  ld is odd, so both walks are bank
                // conflict free.  The new direction -H' g (:173-176) follows from
                //   H' g = H g + (s rho) [ (1+q) s.g - (y^T H).g ] - (H y) rho s.g
                // without reading H' back.  Sweep 2 rewrites the owned columns.
                if constexpr (kVectorRows) {
                    // Compile-time n: rows are 16-byte aligned (wide_ld_aligned) and the n-vectors are zero beyond n, so
                    // everything that walks a ROW moves four values per shared-memory instruction and pairs of products
                    // go through the packed FMA; only (y^T H)_c still reads a COLUMN one word at a time.  Sweep 2
                    // rewrites H by rows: H'[c][:] = H[c][:] + (s_c rho) inner[:] - (H y)_c rho s[:].
                    using V4 = typename Vec4<T>::type;
                    using P2 = typename Vec2<T>::type;
                    constexpr int kN = Obj::kParams, kChunks = (kN + 3) / 4;
                    const V4* Y4 = reinterpret_cast<const V4*>(y);
                    const V4* G4 = reinterpret_cast<const V4*>(g);
                    T yHc[kCols], Hyc[kCols], Hgc[kCols], yc[kCols], sc_[kCols], gc[kCols];
                    bool inc[kCols];
                    int kc[kCols];
#pragma unroll
                    for (int h = 0; h < kCols; ++h) {
                        inc[h] = lane + 32 * h < kN;
                        kc[h] = inc[h] ? lane + 32 * h : 0;
                        yHc[h] = T(0);
                    }
                    P2 hy[kCols], hg[kCols];
#pragma unroll
                    for (int h = 0; h < kCols; ++h) hy[h] = hg[h] = pk(T(0));
#pragma unroll
                    for (int ch = 0; ch < kChunks; ++ch) {
                        const V4 y4 = Y4[ch], g4 = G4[ch];
                        const T ye[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                        for (int h = 0; h < kCols; ++h) {
                            const V4 r4 = *reinterpret_cast<const V4*>(H + kc[h] * ld + 4 * ch);
                            P2 ra, rb, ya, yb, ga, gb;
                            ra.x = r4.x; ra.y = r4.y; rb.x = r4.z; rb.y = r4.w;
                            ya.x = y4.x; ya.y = y4.y; yb.x = y4.z; yb.y = y4.w;
                            ga.x = g4.x; ga.y = g4.y; gb.x = g4.z; gb.y = g4.w;
                            hy[h] = pfma(rb, yb, pfma(ra, ya, hy[h]));
                            hg[h] = pfma(rb, gb, pfma(ra, ga, hg[h]));
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (4 * ch + e < kN) yHc[h] = fma_t(ye[e], H[(4 * ch + e) * ld + kc[h]], yHc[h]);
                        }
                    }
                    T a0 = T(0), a1 = T(0), a2 = T(0);
#pragma unroll
                    for (int h = 0; h < kCols; ++h) {
                        Hyc[h] = inc[h] ? hy[h].x + hy[h].y : T(0);
                        Hgc[h] = inc[h] ? hg[h].x + hg[h].y : T(0);
                        if (!inc[h]) yHc[h] = T(0);
                        yc[h] = inc[h] ? y[kc[h]] : T(0);
                        sc_[h] = inc[h] ? s[kc[h]] : T(0);
                        gc[h] = inc[h] ? g[kc[h]] : T(0);
                        a0 = fma_t(yHc[h], yc[h], a0);
                        a1 = fma_t(sc_[h], gc[h], a1);
                        a2 = fma_t(yHc[h], gc[h], a2);
                    }
                    const T q = mul_rn(warp_allreduce(a0), rho);                   // y^T H y / (y^T s), :271-274
                    const T sg = warp_allreduce(a1);
                    const T yhg = warp_allreduce(a2);
                    const T onepq = add_rn(T(1), q);
                    const T dir = fma_t(onepq, sg, -yhg);
                    __syncwarp();  // sweep 1's reads of y are done before it is overwritten with `inner`
                    T sri[kCols], nhr[kCols];
#pragma unroll
                    for (int h = 0; h < kCols; ++h) {
                        sri[h] = mul_rn(sc_[h], rho);
                        nhr[h] = -mul_rn(Hyc[h], rho);
                        if (inc[h]) {
                            const int c = lane + 32 * h;
                            d[c] = mul_rn(T(-1), fma_t(nhr[h], sg, fma_t(sri[h], dir, Hgc[h])));
                            y[c] = fma_t(sc_[h], onepq, -yHc[h]);   // inner_j = s_j (1+q) - (y^T H)_j
                        }
                    }
                    __syncwarp();
                    const V4* S4 = reinterpret_cast<const V4*>(s);
#pragma unroll
                    for (int h = 0; h < kCols; ++h) {
                        if (inc[h]) {
                            V4* row = reinterpret_cast<V4*>(H + kc[h] * ld);
                            const P2 a2_ = pk(sri[h]), b2_ = pk(nhr[h]);
#pragma unroll
                            for (int ch = 0; ch < kChunks; ++ch) {
                                const V4 in4 = Y4[ch], s4 = S4[ch];
                                V4 r4 = row[ch];
                                P2 ra, rb, ia, ib, sa, sb;
                                ra.x = r4.x; ra.y = r4.y; rb.x = r4.z; rb.y = r4.w;
                                ia.x = in4.x; ia.y = in4.y; ib.x = in4.z; ib.y = in4.w;
                                sa.x = s4.x; sa.y = s4.y; sb.x = s4.z; sb.y = s4.w;
                                ra = pfma(b2_, sa, pfma(a2_, ia, ra));
                                rb = pfma(b2_, sb, pfma(a2_, ib, rb));
                                r4.x = ra.x; r4.y = ra.y; r4.z = rb.x; r4.w = rb.y;
                                row[ch] = r4;
                            }
                        }
                    }
                } else {
                    T yHc[kCols], Hyc[kCols], Hgc[kCols], yc[kCols], sc_[kCols], gc[kCols];
                    int kc[kCols];
                    bool inc[kCols];
#pragma unroll
                    for (int h = 0; h < kCols; ++h) {
                        inc[h] = lane + 32 * h < n;
                        kc[h] = inc[h] ? lane + 32 * h : 0;   // safe index for lanes beyond n (results discarded)
                        yHc[h] = Hyc[h] = Hgc[h] = T(0);
                    }
                    {
                        const T* col[kCols];
                        const T* row[kCols];
#pragma unroll
                        for (int h = 0; h < kCols; ++h) {
                            col[h] = H + kc[h];
                            row[h] = H + kc[h] * ld;
                        }
#pragma unroll 9
                        for (int i = 0; i < n; ++i) {
                            const T yi = y[i], gi = g[i];
#pragma unroll
                            for (int h = 0; h < kCols; ++h) {
                                const T hc = *col[h], hr = row[h][i];
                                col[h] += ld;
                                yHc[h] = fma_t(yi, hc, yHc[h]);
                                Hyc[h] = fma_t(hr, yi, Hyc[h]);
                                Hgc[h] = fma_t(hr, gi, Hgc[h]);
                            }
                        }
                    }
                    T a0 = T(0), a1 = T(0), a2 = T(0);
#pragma unroll
                    for (int h = 0; h < kCols; ++h) {
                        yc[h] = inc[h] ? y[kc[h]] : T(0);
                        sc_[h] = inc[h] ? s[kc[h]] : T(0);
                        gc[h] = inc[h] ? g[kc[h]] : T(0);
                        if (!inc[h]) { yHc[h] = T(0); Hyc[h] = T(0); Hgc[h] = T(0); }
                        a0 = fma_t(yHc[h], yc[h], a0);
                        a1 = fma_t(sc_[h], gc[h], a1);
                        a2 = fma_t(yHc[h], gc[h], a2);
                    }
                    const T q = mul_rn(warp_allreduce(a0), rho);                       // y^T H y / (y^T s), :271-274
                    const T sg = warp_allreduce(a1);
                    const T yhg = warp_allreduce(a2);
                    const T onepq = add_rn(T(1), q);
                    const T dir = fma_t(onepq, sg, -yhg);
                    __syncwarp();  // sweep 1's reads of y are done before it is overwritten with the row constants
#pragma unroll
                    for (int h = 0; h < kCols; ++h)
                        if (inc[h]) {
                            const int c = lane + 32 * h;
                            d[c] = mul_rn(T(-1), fma_t(-Hyc[h] * rho, sg, fma_t(sc_[h] * rho, dir, Hgc[h])));
                            y[c] = mul_rn(sc_[h], rho);       // row constants of sweep 2: s_i rho ...
                            Hy[c] = -mul_rn(Hyc[h], rho);     // ... and -(H y)_i rho
                        }
                    __syncwarp();
                    {
                        T inner[kCols];
                        T* col[kCols];
                        int st[kCols];
#pragma unroll
                        for (int h = 0; h < kCols; ++h) {
                            inner[h] = fma_t(sc_[h], onepq, -yHc[h]);                  // s_j (1+q) - (y^T H)_j
                            // a lane without column h updates a scratch word of xt instead of predicating every store
                            col[h] = inc[h] ? H + kc[h] : ws.xt + (32 * h + lane) % wide_vec(n);
                            st[h] = inc[h] ? ld : 0;
                        }
#pragma unroll 9
                        for (int i = 0; i < n; ++i) {
                            const T sri = y[i], nhr = Hy[i];
#pragma unroll
                            for (int h = 0; h < kCols; ++h) {
                                *col[h] = fma_t(nhr, sc_[h], fma_t(sri, inner[h], *col[h]));
                                col[h] += st[h];
                            }
                        }
                    }
                }
#endif
            }
            __syncwarp();
            // ---- line-search set-up, wolfe_conditions.py:77-114 ----
            f0 = f;
            g0 = wide_dot<kCols>(d, g, n, lane);                                       // :77
            neg_c2_g0 = mul_rn(T(-1) * p.c2, g0);
            widening = true; zooming = false;                                   // :80-82
            lo = T(0); hi = T(0); cand = T(1);                                  // :97-108
            lo_f = f0; hi_f = f0; cand_f = f0;                                  // :109-111
            if (Rec::kActive) lo_d = hi_d = cand_d = g0;
            ls_i = 0;
            mode = kEvalProbe;
        }
    }
    __syncwarp();
    for (int c = lane; c < n; c += 32) p.x_out[(size_t)b * n + c] = x[c];
    if (lane == 0) {
        if (p.cost_out) p.cost_out[b] = f;
        if (p.converged_out) p.converged_out[b] = (f <= p.thr) ? 1 : 0;
        if (p.iters_out) p.iters_out[b] = iters;
        if (p.fevals_out) p.fevals_out[b] = fevals;
        if (p.reason_out) p.reason_out[b] = reason;
    }
    rec.finish(b, iters, lane);
}

}  // namespace davo
