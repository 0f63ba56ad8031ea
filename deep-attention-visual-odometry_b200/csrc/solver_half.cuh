// solver_half.cuh — TWO problems per warp: each half-warp (16 lanes) solves one DISTORT10 problem (n = 10).
//
// Same state machine as solver_warp.cuh, restated from
//   autograd_solvers/bfgs_solver.py:80-303, utils/func_inverse_curvature.py:8-11,
//   autograd_solvers/line_search/wolfe_conditions.py:23-253,
// and the same objective (objectives.cuh match_pair_cost_grad: camera_model/distorted_camera_model.py:59-86,
// solvers/least_squares_utils.py:4-48).
//
// Why halves.  The warp-per-problem kernel is bound by instruction issue, and of its ~570 issue slots per
// objective evaluation only ~212 are the packed match arithmetic: the rest — the cross-lane sum of the 11
// partial sums, the parameter broadcast, the line-search bookkeeping, the BFGS update — costs the same number of
// warp instructions whether 32 or 16 lanes take part.  With two problems side by side every one of those
// instructions serves two problems, while the match loop takes twice the trips for twice the problems.
//
// Layout.  Lane l = lane & 15 of a half owns component l of x, g, d, s (zero for l >= 10) and row l of the
// inverse Hessian in registers.  Scalars of the state machine are replicated within a half.  Both halves run
// ONE loop with ONE evaluation site; what differs between them (mode of the evaluation, line-search state,
// whether a BFGS update is due, retirement) is per-half data, so the two halves only diverge in the short
// bookkeeping branches.  Shuffles and warp syncs inside those branches use the half's own lane mask.  A half
// whose problem retires pulls the next index from the work queue, starts the bulk TMA copy of that problem's
// matches and keeps looping (idle) until its mbarrier completes — it never blocks its partner.
#pragma once
#include "davo_common.cuh"
#include "objectives.cuh"
#include "solver_warp.cuh"  // same_bits

namespace davo {

constexpr int kHalfLanes = 16;
#ifndef DAVO_HALF_MIN_BLOCKS
#define DAVO_HALF_MIN_BLOCKS 4
#endif
#ifndef DAVO_HALF_UNROLL
#define DAVO_HALF_UNROLL 2
#endif
#define DAVO_PRAGMA_(x) _Pragma(#x)
#define DAVO_PRAGMA_UNROLL(n) DAVO_PRAGMA_(unroll n)
// transpose scratch of one half: 16 rows x kRedPitch words, + 16 words so that the two halves of a warp start
// 16 banks apart (their simultaneous column reads then cover all 32 banks once)
constexpr int kHalfScratch = kHalfLanes * kRedPitch + 16;

template <typename T>
__host__ __device__ inline size_t half_slab_bytes(int N) {
    return sizeof(typename Vec4<T>::type) * (size_t)(N + 32);  // pair layout: up to the next multiple of 32 entries
}
// shared memory of one half: match slab | trial-point line | broadcast line | transpose scratch | mbarrier
template <typename T>
__host__ __device__ inline size_t half_stride(int N) {
    size_t b = half_slab_bytes<T>(N) + (2 * kSlots + kHalfScratch) * sizeof(T) + 16;
    return (b + 127) & ~size_t(127);
}

template <typename T>
__device__ __forceinline__ T half_shfl_xor(unsigned mask, T v, int m) { return __shfl_xor_sync(mask, v, m); }

// sum over the 16 lanes of a half; every lane of the half ends with the bitwise-identical total
template <typename T>
__device__ __forceinline__ T half_allreduce(unsigned mask, T v) {
    v += half_shfl_xor(mask, v, 1);
    v += half_shfl_xor(mask, v, 2);
    v += half_shfl_xor(mask, v, 4);
    v += half_shfl_xor(mask, v, 8);
    return v;
}

// two distributed vectors -> all lanes of the half (one store + vector loads each)
template <typename T, int NP>
__device__ __forceinline__ void half_gather2(unsigned mask, T own_a, T own_b, T* line_a, T* line_b, int l,
                                             T (&out_a)[NP], T (&out_b)[NP]) {
    using V4 = typename Vec4<T>::type;
    __syncwarp(mask);  // earlier readers of the lines are done
    line_a[l] = own_a;
    line_b[l] = own_b;
    __syncwarp(mask);
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

// kN > 0 fixes the number of matches per problem at compile time (256 in BASELINE configs 2, 4, 5: shared-memory
// offsets become immediates, the match loop has a constant trip count); 0 = run time.
template <typename T, bool kRagged, int kN = 0>
__global__ void __launch_bounds__(128, sizeof(T) == 4 ? DAVO_HALF_MIN_BLOCKS : 2) half_problem_kernel(const SolveParams<T> p) {
    using V4 = typename Vec4<T>::type;
    using V2 = typename Vec2<T>::type;
    using P = typename Vec2<T>::type;
    constexpr int NP = 10;
    enum { kEvalOuter, kEvalProbe, kEvalFinal };
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane & 15, half = lane >> 4;
    const unsigned hmask = 0xffffu << (half * 16);
    const int N = kN > 0 ? kN : p.N, n = p.n;
    unsigned char* mine = smem + (size_t)(warp * 2 + half) * half_stride<T>(N);
    V4* matches = reinterpret_cast<V4*>(mine);
    T* xt_line = reinterpret_cast<T*>(mine + half_slab_bytes<T>(N));
    T* bc_line = xt_line + kSlots;
    T* scratch = bc_line + kSlots;
    uint64_t* bar = reinterpret_cast<uint64_t*>(scratch + kHalfScratch);
    if (l == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    // The straggler launch behind this one is a programmatic dependent launch: its CTAs may take the slots this grid
    // frees as its warps run out of problems, and start on the problems handed off so far (solve_kernels.cu).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#if DAVO_TIMELINE
    if (blockIdx.x == 0 && threadIdx.x == 0) printf("T primary start at %llu\n", global_ns());
#endif

    const bool own = l < n;
    // ---- per-half scheduling state ----
    bool have = false, waiting = false, done = false;
    unsigned parity = 0;
    int b = 0;
    // ---- per-problem solver state (solver_warp.cuh) ----
    T x = T(0), g = T(0), gprev = T(0), d = T(0), s = T(0), f = T(0);
    T H[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) H[j] = T(0);
    int iters = 0, fevals = 0, reason = DAVO_REASON_CAP, k = 0;
    bool widening = false, zooming = false;
    T lo = T(0), hi = T(0), cand = T(1), lo_f = T(0), hi_f = T(0), cand_f = T(0), f0 = T(0), g0 = T(0),
      neg_c2_g0 = T(0), gt = T(0);
    int ls_i = 0;
    int mode = kEvalFinal;

    for (;;) {
        // ---- a half without a problem pulls the next one and starts staging it (bulk TMA) ----
        if (!have && !done) {
            unsigned nb = 0;
            if (l == 0) nb = atomicAdd(p.queue, 1u);
            nb = __shfl_sync(hmask, nb, half * 16);
            if (nb >= (unsigned)p.B) {
                done = true;
            } else {
                b = (int)nb;
                fence_proxy_async();  // this lane's generic-proxy writes to the slab (pair layout) precede the async write
                __syncwarp(hmask);    // every lane of the half is done with the previous problem's slab
                if (l == 0) {
                    const unsigned bytes = (unsigned)(sizeof(V4) * (size_t)N);
                    mbar_expect_tx(bar, bytes);
                    tma_load_1d(matches, p.data0 + (size_t)b * N * 4, bytes, bar);
                }
                have = true;
                waiting = true;
            }
        }
        if (__ballot_sync(kFull, have) == 0u) {  // both halves are out of work
#if DAVO_TIMELINE
            if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
                printf("T primary block %u warp %d exits at %llu\n", blockIdx.x, warp, global_ns());
#endif
            break;
        }
        unsigned landed = 0;
        if (waiting) {  // the half's leader polls once per trip; the answer is made uniform within the half
            if (l == 0) landed = mbar_try_wait(bar, parity) ? 1u : 0u;
            landed = __shfl_sync(hmask, landed, half * 16);
        }
        if (landed) {
            mbar_wait(bar, parity);  // completes at once; every lane observes the phase itself (acquire)
            parity ^= 1u;
            // Pair layout, once per problem: this lane's matches i and i + 16 become entry i = {a_i, a_i+16, b_i, b_i+16}
            // and entry i + 16 = -{u*_i, u*_i+16, v*_i, v*_i+16}: one LDS.128 yields two aligned register pairs for the
            // packed arithmetic.  A lane rewrites only its own two entries; a missing second match is zero-filled.
            for (int i = l; i < N; i += 32) {
                const V4 m0 = matches[i];
                V4 m1;
                m1.x = m1.y = m1.z = m1.w = T(0);
                if (i + 16 < N) m1 = matches[i + 16];
                V4 ab, ob;
                ab.x = m0.x; ab.y = m1.x; ab.z = m0.y; ab.w = m1.y;
                ob.x = -m0.z; ob.y = -m1.z; ob.z = -m0.w; ob.w = -m1.w;
                matches[i] = ab;
                matches[i + 16] = ob;
            }
            x = own ? p.x0[(size_t)b * n + l] : T(0);
            g = gprev = d = s = f = T(0);
#pragma unroll
            for (int j = 0; j < NP; ++j) H[j] = (j == l) ? T(1) : T(0);  // bfgs_solver.py:112-117
            iters = 0; fevals = 0; reason = DAVO_REASON_CAP; k = 0;
            mode = (p.max_iters > 0) ? kEvalOuter : kEvalFinal;
            waiting = false;
        }
        const bool active = have && !waiting;

        // ---- the one evaluation site: both halves, each at its own point of its own problem ----
        const T pt = (mode == kEvalProbe) ? add_rn(x, mul_rn(cand, d)) : x;  // wolfe_conditions.py:139
        __syncwarp();
        xt_line[l] = pt;
        __syncwarp();
        T fe, ge;
        {
            Intrinsics<T> I;
            I.load(xt_line);
            P acc2[kPairAcc];
#pragma unroll
            for (int q = 0; q < kPairAcc; ++q) acc2[q] = pk(T(0));
            P gu, gv;
            // compile-time N: the whole loop is unrolled (8 pairs per lane at N = 256: 3.87 ms against 4.05 ms
            // unrolled by 2 for 64K x 256)
            constexpr int kUnroll = kN > 0 ? (kN + 31) / 32 : DAVO_HALF_UNROLL;
#pragma unroll kUnroll
            for (int i = l; i < N; i += 32) {
                const V4 ab = matches[i];
                const V4 ob = matches[i + 16];
                P a, bb, nus, nvs;
                a.x = ab.x; a.y = ab.y; bb.x = ab.z; bb.y = ab.w;
                nus.x = ob.x; nus.y = ob.y; nvs.x = ob.z; nvs.y = ob.w;
                const bool second_valid = !kRagged || i + 16 < N;
                match_pair_cost_grad<T, false>(I, a, bb, nus, nvs, pk(T(1)), second_valid, acc2, gu, gv);
            }
            T acc[kSlots];
#pragma unroll
            for (int q = 0; q < kPairAcc; ++q) acc[q] = acc2[q].x + acc2[q].y;
            fold_uv_terms(acc);  // acc[11], acc[12] are folded into acc[5], acc[6]: 11 sums remain
            // An overflowed evaluation is redone in the reference's association (objectives.cuh
            // match_cost_grad_unfused): whether the cost is inf or NaN steers the reference's line search.  Rare
            // (BASELINE config 4's points at z -> 0+); both halves take the branch together, each keeps its own sums.
            if (__any_sync(kFull, !isfinite(acc[10]))) {
                const bool mine_bad = __any_sync(hmask, !isfinite(acc[10]));
                T ref[kSlots];
#pragma unroll
                for (int q = 0; q < kSlots; ++q) ref[q] = T(0);
                T gu1, gv1;
#pragma unroll 1
                for (int i = l; i < N; i += 32) {
                    const V4 ab = matches[i];
                    const V4 ob = matches[i + 16];
                    match_cost_grad_unfused<T, false>(I, ab.x, ab.z, -ob.x, -ob.z, T(1), ref, gu1, gv1);
                    if (i + 16 < N) match_cost_grad_unfused<T, false>(I, ab.y, ab.w, -ob.y, -ob.w, T(1), ref, gu1, gv1);
                }
                if (mine_bad) {
#pragma unroll
                    for (int q = 0; q < 11; ++q) acc[q] = ref[q];
                }
            }
            // Sum over the 16 lanes as a transpose through shared memory: lane l stores its 11 partial sums as row l,
            // lane c adds up column c.
            V4* row = reinterpret_cast<V4*>(scratch + l * kRedPitch);
            V4 q0, q1, q2;
            q0.x = acc[0]; q0.y = acc[1]; q0.z = acc[2]; q0.w = acc[3];
            q1.x = acc[4]; q1.y = acc[5]; q1.z = acc[6]; q1.w = acc[7];
            q2.x = acc[8]; q2.y = acc[9]; q2.z = acc[10]; q2.w = T(0);
            row[0] = q0; row[1] = q1; row[2] = q2;
            __syncwarp();
            const T* col = scratch + min(l, 10);
            T part[kHalfLanes];
#pragma unroll
            for (int i = 0; i < kHalfLanes; ++i) part[i] = col[i * kRedPitch];
#pragma unroll
            for (int w = kHalfLanes / 2; w > 0; w >>= 1)
#pragma unroll
                for (int i = 0; i < w; ++i) part[i] += part[i + w];
            fe = __shfl_sync(kFull, part[0], 10, kHalfLanes);      // slot 10 = cost
            ge = (l < 10) ? T(2) * part[0] : T(0);                 // least_squares_utils.py:43 (factor 2)
        }

        if (active) {
            bool start_iteration = false, finished = false;
            if (mode == kEvalFinal) {                                         // networks/calibration_network.py:71
                f = fe;
                finished = true;
            } else if (mode == kEvalOuter) {                                  // bfgs_solver.py:128-135
                f = fe;
                g = ge;
                start_iteration = true;
            } else {
                // ---- one line-search probe has been evaluated: wolfe_conditions.py:143-237 ----
                cand_f = fe;
                gt = ge;
                ++fevals;
                const T dphi = half_allreduce(hmask, mul_rn(d, gt));          // d/d alpha f(x + alpha d), :141
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
                } else {
                    // ---- line search finished, alpha = upper_alpha (:239): take the step, bfgs_solver.py:191-207 ----
                    ++iters;
                    s = mul_rn(hi, d);                                        // :191
                    x = add_rn(x, s);                                         // :192
                    const T nrm = sqrt_rn(half_allreduce(hmask, mul_rn(s, s)));  // :203-205
                    gprev = g;
                    // The accepted point is bitwise x + alpha d: when the search returns the probe it has just made,
                    // that probe's (f, grad) ARE the next outer iteration's evaluation (solver_warp.cuh).
                    const bool reuse = same_bits(hi, cand);
                    ++k;
                    const bool stop_step = !(nrm > p.min_step);               // :203-207 (strict >)
                    if (stop_step || k >= p.max_iters) {                      // :118
                        reason = stop_step ? DAVO_REASON_STEP : DAVO_REASON_CAP;
                        if (reuse) {
                            f = cand_f;
                            finished = true;
                        } else {
                            mode = kEvalFinal;
                        }
                    } else if (!reuse) {
                        mode = kEvalOuter;
                    } else {
                        f = cand_f;
                        g = gt;
                        start_iteration = true;
                    }
                }
            }
            if (start_iteration) {
                // ---- top of an outer iteration with (f, g) at x: bfgs_solver.py:136-190 ----
                ++fevals;  // the reference evaluates here even when the probe is reused
                if (!(f > p.thr)) {                                           // :143 (strict >; NaN retires)
                    reason = (f <= p.thr) ? DAVO_REASON_THRESHOLD : DAVO_REASON_NAN;
                    finished = true;
                } else if (p.eval_cap > 0 && fevals > p.eval_cap) {
                    // A straggler (a problem that bisects every line search down to lo == hi for hundreds of
                    // iterations): its latency, not the batch's throughput, would set the launch time.  Hand it
                    // to the second launch, which gives it a whole warp.  The test depends on the problem's own
                    // trajectory only, so which problems are handed off is deterministic.
                    if (l == 0) {
                        p.reason_out[b] = kReasonHandoff;
                        __threadfence();
                        const unsigned slot = atomicAdd(p.queue + kWsReserved, 1u);   // publish to the second launch
                        if (slot < (unsigned)kHandoffList) atomicExch(p.queue + kWsList + slot, (unsigned)b + 1u);
#if DAVO_TIMELINE
                        printf("T handoff slot %u problem %d fevals %d at %llu\n", slot, b, fevals, global_ns());
#endif
                    }
                    have = false;
                    mode = kEvalFinal;
                } else {
                    if (k == 0) {
                        d = mul_rn(T(-1), g);                                 // :152-155
                    } else {
                        const T y = sub_rn(g, gprev);                         // :157
                        const T sy = half_allreduce(hmask, mul_rn(s, y));     // y^T s
                        if (k == 1) {                                         // :159-167, :217-233 (eq. 6.20)
                            T den = half_allreduce(hmask, mul_rn(y, y));
                            den = (den < T(1e-5)) ? T(1e-5) : den;
                            T sc = div_rn(sy, den);
                            sc = (sc < T(1e-4)) ? T(1e-4) : sc;
#pragma unroll
                            for (int j = 0; j < NP; ++j) H[j] = mul_rn(sc, H[j]);
                        }
                        T rho = div_rn(T(1), sy);                             // func_inverse_curvature.py:8-11
                        if (sy <= T(0)) rho = T(0);
                        // H + (s rho) s^T (1+q) - (s rho)(y^T H) - (H y)(s rho)^T with the old H on the right
                        // (:263-303); y^T H is formed from the COLUMNS of H (see solver_warp.cuh): the rows go
                        // through the half's scratch area and come back as columns.
                        T yv[NP], sv[NP], gvv[NP], yHv[NP];
                        half_gather2<T, NP>(hmask, y, s, bc_line, xt_line, l, yv, sv);
                        if (l < NP) {
                            T* rowp = scratch + l * kRedPitch;
                            V4 r0, r1;
                            V2 r2;
                            r0.x = H[0]; r0.y = H[1]; r0.z = H[2]; r0.w = H[3];
                            r1.x = H[4]; r1.y = H[5]; r1.z = H[6]; r1.w = H[7];
                            r2.x = H[8]; r2.y = H[9];
                            reinterpret_cast<V4*>(rowp)[0] = r0;
                            reinterpret_cast<V4*>(rowp)[1] = r1;
                            *reinterpret_cast<V2*>(rowp + 8) = r2;
                        }
                        __syncwarp(hmask);
                        T Hy = T(0), yH = T(0);                               // (H y)_c :293-295, (y^T H)_c :268-270
                        {
                            const T* colp = scratch + (own ? l : 0);
#pragma unroll
                            for (int j = 0; j < NP; ++j) {
                                Hy = fma_t(H[j], yv[j], Hy);
                                yH = fma_t(yv[j], colp[j * kRedPitch], yH);
                            }
                            if (!own) yH = T(0);
                        }
                        const T q = mul_rn(half_allreduce(hmask, yH * y), rho);  // y^T H y / (y^T s), :271-274
                        const T onepq = add_rn(T(1), q);
                        const T sr = mul_rn(s, rho);                          // :277
                        const T nHyrho = -mul_rn(Hy, rho);
                        half_gather2<T, NP>(hmask, yH, g, bc_line, xt_line, l, yHv, gvv);
                        T Hg = T(0);
#pragma unroll
                        for (int j = 0; j < NP; ++j) {
                            const T inner = fma_t(sv[j], onepq, -yHv[j]);     // s_j (1+q) - (y^T H)_j
                            H[j] = fma_t(nHyrho, sv[j], fma_t(sr, inner, H[j]));
                            Hg = fma_t(H[j], gvv[j], Hg);
                        }
                        d = own ? mul_rn(T(-1), Hg) : T(0);                   // :173-176
                    }
                    // ---- line-search set-up, wolfe_conditions.py:77-114 ----
                    f0 = f;
                    g0 = half_allreduce(hmask, mul_rn(d, g));                 // :77
                    neg_c2_g0 = mul_rn(T(-1) * p.c2, g0);
                    widening = true; zooming = false;                         // :80-82
                    lo = T(0); hi = T(0); cand = T(1);                        // :97-108
                    lo_f = f0; hi_f = f0; cand_f = f0;                        // :109-111
                    ls_i = 0;
                    mode = kEvalProbe;
                }
            }
            if (finished) {
                if (own) p.x_out[(size_t)b * n + l] = x;
                if (l == 0) {
                    if (p.cost_out) p.cost_out[b] = f;
                    if (p.converged_out) p.converged_out[b] = (f <= p.thr) ? 1 : 0;
                    if (p.iters_out) p.iters_out[b] = iters;
                    if (p.fevals_out) p.fevals_out[b] = fevals;
                    if (p.reason_out) p.reason_out[b] = reason;
                }
                have = false;
                mode = kEvalFinal;
            }
        }
    }
    if (lane == 0) {   // this warp is out of work: every hand-off it made is published before the count moves
        __threadfence();
        atomicAdd(p.queue + kWsExited, 1u);
    }
}

}  // namespace davo
