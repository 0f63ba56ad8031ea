// objectives.cuh — warp-cooperative objective evaluators for the warp-per-problem solver.
//
// Interface (one warp, one problem):
//   bind(b)                  make problem b current (stage its matches into shared memory)
//   eval(th, f, g_own)       th: >= 16 parameters in shared memory (slot order, zero padded);
//                            f: objective value, bitwise identical in all 32 lanes;
//                            g_own: d f / d theta_c for the slot c = lane >> 1 this lane owns (0 if c >= n)
//
// The camera model follows /root/reference/deep_attention_visual_odometry/camera_model/
// distorted_camera_model.py:59-86 (intrinsics + Brown-Conrady distortion) and
// solvers/least_squares_utils.py:4-48 (residual, squared error, gradient = sum 2 w r J).  The
// gradient is J^T r of THAT forward model accumulated on the fly; the
// reference's 2N x 16 Jacobian is never materialised (and is wrong in 8 columns at HEAD).
#pragma once
#include "davo_common.cuh"

#ifndef DAVO_ORACLE_ARITH
#define DAVO_ORACLE_ARITH 0
#endif

namespace davo {

// Per-evaluation constants derived from the 10 intrinsics, hoisted out of the match loop.
template <typename T>
struct Intrinsics {
    T cx, cy, k1, k2, k3, p1, p2, fx, s, fy;
    T k1x2, k2x4, k3x6, p1x2, p2x2, p1x6, p2x6;
    __device__ __forceinline__ void load(const T* th) {
        using V4 = typename Vec4<T>::type;
        const V4* t4 = reinterpret_cast<const V4*>(th);  // the parameter line is 16-byte aligned, >= 12 entries
        const V4 q0 = t4[0], q1 = t4[1], q2 = t4[2];
        cx = q0.x; cy = q0.y; k1 = q0.z; k2 = q0.w; k3 = q1.x;
        p1 = q1.y; p2 = q1.z; fx = q1.w; s = q2.x; fy = q2.y;
        k1x2 = T(2) * k1; k2x4 = T(4) * k2; k3x6 = T(6) * k3;
        p1x2 = T(2) * p1; p2x2 = T(2) * p2; p1x6 = T(6) * p1; p2x6 = T(6) * p2;
    }
};

// One match in the REFERENCE's association, operation by operation, without FMA: the expressions of
// camera_model/distorted_camera_model.py:59-86 exactly as written (rad = 1 + k1 r2 + k2 r2 r2 + k3 r2 r2 r2, not
// Horner), as restated by oracle/calib_oracle_impl.h match_intrinsics.  Used (a) when an evaluation overflows: whether
// a cost is +inf or NaN decides how the reference's line search continues (inf: "sufficient decrease failed", bisect
// back; NaN: every comparison false, keep doubling until the probe cap and retire on a NaN cost), and the Horner / FMA
// form of the fast path turns many of the reference's inf - inf = NaN into +-inf (BASELINE config 4's points at
// z -> 0+); (b) by DAVO_ORACLE_ARITH=1 debug builds for every evaluation.
template <typename T, bool kWeighted>
__device__ __forceinline__ void match_cost_grad_unfused(const Intrinsics<T>& I, T a, T b, T us, T vs, T w,
                                                        T (&acc)[kSlots], T& gu, T& gv) {
    const T u = add_rn(mul_rn(I.fx, a), mul_rn(I.s, b));
    const T v = mul_rn(I.fy, b);
    const T r2 = add_rn(mul_rn(u, u), mul_rn(v, v));
    const T uv = mul_rn(u, v);
    const T rad = add_rn(add_rn(add_rn(T(1), mul_rn(I.k1, r2)), mul_rn(mul_rn(I.k2, r2), r2)),
                         mul_rn(mul_rn(mul_rn(I.k3, r2), r2), r2));
    const T A = add_rn(r2, mul_rn(mul_rn(T(2), u), u)), Bv = add_rn(r2, mul_rn(mul_rn(T(2), v), v));
    const T up = add_rn(add_rn(add_rn(mul_rn(u, rad), mul_rn(mul_rn(T(2), I.p1), uv)), mul_rn(I.p2, A)), I.cx);
    const T vp = add_rn(add_rn(add_rn(mul_rn(v, rad), mul_rn(mul_rn(T(2), I.p2), uv)), mul_rn(I.p1, Bv)), I.cy);
    T ru = sub_rn(up, us), rv = sub_rn(vp, vs);
    T cost = add_rn(mul_rn(ru, ru), mul_rn(rv, rv));
    if (kWeighted) { cost = mul_rn(w, cost); ru = mul_rn(w, ru); rv = mul_rn(w, rv); }
    acc[10] = add_rn(acc[10], cost);
    const T r4 = mul_rn(r2, r2), r6 = mul_rn(r4, r2);
    const T radp = add_rn(add_rn(I.k1, mul_rn(mul_rn(T(2), I.k2), r2)), mul_rn(mul_rn(T(3), I.k3), r4));
    const T uv2 = mul_rn(T(2), uv);
    const T Duu = add_rn(add_rn(add_rn(rad, mul_rn(mul_rn(mul_rn(T(2), u), u), radp)), mul_rn(mul_rn(T(2), I.p1), v)),
                         mul_rn(mul_rn(T(6), I.p2), u));
    const T Dvv = add_rn(add_rn(add_rn(rad, mul_rn(mul_rn(mul_rn(T(2), v), v), radp)), mul_rn(mul_rn(T(6), I.p1), v)),
                         mul_rn(mul_rn(T(2), I.p2), u));
    const T Duv = add_rn(add_rn(mul_rn(uv2, radp), mul_rn(mul_rn(T(2), I.p1), u)), mul_rn(mul_rn(T(2), I.p2), v));
    gu = add_rn(mul_rn(ru, Duu), mul_rn(rv, Duv));
    gv = add_rn(mul_rn(ru, Duv), mul_rn(rv, Dvv));
    const T t = add_rn(mul_rn(ru, u), mul_rn(rv, v));
    acc[0] = add_rn(acc[0], ru);
    acc[1] = add_rn(acc[1], rv);
    acc[2] = add_rn(acc[2], mul_rn(t, r2));
    acc[3] = add_rn(acc[3], mul_rn(t, r4));
    acc[4] = add_rn(acc[4], mul_rn(t, r6));
    acc[5] = add_rn(acc[5], add_rn(mul_rn(ru, uv2), mul_rn(rv, Bv)));
    acc[6] = add_rn(acc[6], add_rn(mul_rn(ru, A), mul_rn(rv, uv2)));
    acc[7] = add_rn(acc[7], mul_rn(gu, a));
    acc[8] = add_rn(acc[8], mul_rn(gu, b));
    acc[9] = add_rn(acc[9], mul_rn(gv, b));
}

// One match's weighted residual w * (u' - u*, v' - v*) alone (distorted_camera_model.py:59-86): d cost / d (u*, v*) is
// -2 times it, which is what the backward pass of the differentiable solve differences (solver_train.cuh).
template <typename T>
__device__ __forceinline__ void match_residual(const Intrinsics<T>& I, T a, T b, T us, T vs, T w, T& ru, T& rv) {
    const T u = fma_t(I.fx, a, I.s * b), v = I.fy * b;
    const T uu = u * u, vv = v * v, uv = u * v, r2 = uu + vv;
    const T rad = fma_t(r2, fma_t(r2, fma_t(r2, I.k3, I.k2), I.k1), T(1));
    ru = w * (u * rad + T(2) * I.p1 * uv + I.p2 * (r2 + T(2) * uu) + I.cx - us);
    rv = w * (v * rad + T(2) * I.p2 * uv + I.p1 * (r2 + T(2) * vv) + I.cy - vs);
}

// One match: residuals, squared error and the intrinsic-gradient accumulators.
//   acc[0..9]  += d(cost/2)/d(cx,cy,k1,k2,k3,p1,p2,fx,s,fy)   (acc[5], acc[6] without their 2uv terms)
//   acc[10]    += cost
//   acc[11]    += ru * uv,  acc[12] += rv * uv                 (folded into acc[5], acc[6] by fold_uv_terms)
// (gu, gv) = d(cost/2)/d(u, v) are returned for the pose chain rule of the JOINT model.
// 53 FP32 instructions; counted as 93 flop by SURVEY.md 8(d).
template <typename T, bool kWeighted>
__device__ __forceinline__ void match_cost_grad(const Intrinsics<T>& I, T a, T b, T us, T vs, T w,
                                                T (&acc)[kSlots], T& gu, T& gv) {
#if DAVO_ORACLE_ARITH
    match_cost_grad_unfused<T, kWeighted>(I, a, b, us, vs, w, acc, gu, gv);
#else
    const T u = fma_t(I.fx, a, I.s * b);            // distorted_camera_model.py:59-61
    const T v = I.fy * b;                           // :62
    const T uu = u * u, vv = v * v, uv = u * v;     // :65
    const T r2 = uu + vv;                           // :64
    const T rad = fma_t(r2, fma_t(r2, fma_t(r2, I.k3, I.k2), I.k1), T(1));  // :66-74 (Horner)
    const T A = fma_t(T(2), uu, r2);                // r2 + 2u^2
    const T Bv = fma_t(T(2), vv, r2);               // r2 + 2v^2
    T ru = fma_t(u, rad, I.cx - us);                // :75-80 minus the observation
    ru = fma_t(I.p1x2, uv, ru);
    ru = fma_t(I.p2, A, ru);
    T rv = fma_t(v, rad, I.cy - vs);                // :81-86
    rv = fma_t(I.p2x2, uv, rv);
    rv = fma_t(I.p1, Bv, rv);
    if (kWeighted) {
        acc[10] = fma_t(w, fma_t(ru, ru, rv * rv), acc[10]);  // least_squares_utils.py:24-28
        ru *= w;                                               // :43-45
        rv *= w;
    } else {
        acc[10] = fma_t(ru, ru, acc[10]);
        acc[10] = fma_t(rv, rv, acc[10]);
    }
    const T radp2 = fma_t(r2, fma_t(r2, I.k3x6, I.k2x4), I.k1x2);  // 2 d rad / d r2
    T Duu = fma_t(uu, radp2, rad);                  // d u'/d u = rad + 2u^2 rad' + 2 p1 v + 6 p2 u
    Duu = fma_t(I.p1x2, v, Duu);
    Duu = fma_t(I.p2x6, u, Duu);
    T Dvv = fma_t(vv, radp2, rad);                  // d v'/d v = rad + 2v^2 rad' + 6 p1 v + 2 p2 u
    Dvv = fma_t(I.p1x6, v, Dvv);
    Dvv = fma_t(I.p2x2, u, Dvv);
    T Duv = uv * radp2;                             // d u'/d v = d v'/d u = 2uv rad' + 2 p1 u + 2 p2 v
    Duv = fma_t(I.p1x2, u, Duv);
    Duv = fma_t(I.p2x2, v, Duv);
    gu = fma_t(ru, Duu, rv * Duv);
    gv = fma_t(ru, Duv, rv * Dvv);
    const T t = fma_t(ru, u, rv * v);
    const T r4 = r2 * r2, r6 = r4 * r2;
    acc[0] += ru;
    acc[1] += rv;
    acc[2] = fma_t(t, r2, acc[2]);
    acc[3] = fma_t(t, r4, acc[3]);
    acc[4] = fma_t(t, r6, acc[4]);
    acc[5] = fma_t(rv, Bv, acc[5]);
    acc[6] = fma_t(ru, A, acc[6]);
    acc[11] = fma_t(ru, uv, acc[11]);
    acc[12] = fma_t(rv, uv, acc[12]);
    acc[7] = fma_t(gu, a, acc[7]);
    acc[8] = fma_t(gu, b, acc[8]);
    acc[9] = fma_t(gv, b, acc[9]);
#endif
}

// ---- packed pairs: two matches per instruction ---------------------------------------------------------
// sm_100 has fma/mul/add.rn.f32x2 (SASS FFMA2 / FMUL2 / FADD2): both components are rounded exactly like the
// scalar instruction, the FMA pipe spends the same cycles as for two scalar instructions, but the pair takes ONE
// issue slot, and a per-evaluation constant can be read as a broadcast 32-bit operand (no duplicated registers).
// The solve kernel is issue bound (tools/ffma2_probe.cu: 16 FFMA + 12 integer instructions per round run 1.39x
// faster with the FFMAs paired), so the match loop works on PAIRS of matches.  The float64 instantiation uses the
// same code with component-wise arithmetic.
__device__ __forceinline__ float2 pk(float s) { return make_float2(s, s); }
__device__ __forceinline__ double2 pk(double s) { return make_double2(s, s); }
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 pmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ double2 pfma(double2 a, double2 b, double2 c) {
    return make_double2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y));
}
__device__ __forceinline__ double2 pmul(double2 a, double2 b) { return make_double2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ double2 padd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

constexpr int kPairAcc = 13;  // acc[0..9] gradient sums, acc[10] cost, acc[11], acc[12] the uv terms of p1, p2

// match_cost_grad for two matches at once.  nus, nvs are the NEGATED observations (negated once, when the problem
// is staged into shared memory), second_valid = false zeroes the residuals of the pair's second match (ragged N):
// every accumulated quantity is linear in (ru, rv), so that removes it from all sums.
template <typename T, bool kWeighted>
__device__ __forceinline__ void match_pair_cost_grad(const Intrinsics<T>& I, typename Vec2<T>::type a,
                                                     typename Vec2<T>::type b, typename Vec2<T>::type nus,
                                                     typename Vec2<T>::type nvs, typename Vec2<T>::type w,
                                                     bool second_valid, typename Vec2<T>::type (&acc)[kPairAcc],
                                                     typename Vec2<T>::type& gu, typename Vec2<T>::type& gv) {
    using P = typename Vec2<T>::type;
    const P u = pfma(pk(I.fx), a, pmul(pk(I.s), b));                // distorted_camera_model.py:59-61
    const P v = pmul(pk(I.fy), b);                                  // :62
    const P uu = pmul(u, u), vv = pmul(v, v), uv = pmul(u, v);      // :65
    const P r2 = padd(uu, vv);                                      // :64
    const P rad = pfma(r2, pfma(r2, pfma(r2, pk(I.k3), pk(I.k2)), pk(I.k1)), pk(T(1)));  // :66-74 (Horner)
    const P A = pfma(pk(T(2)), uu, r2);                             // r2 + 2u^2
    const P Bv = pfma(pk(T(2)), vv, r2);                            // r2 + 2v^2
    P ru = pfma(u, rad, padd(pk(I.cx), nus));                       // :75-80 minus the observation
    ru = pfma(pk(I.p1x2), uv, ru);
    ru = pfma(pk(I.p2), A, ru);
    P rv = pfma(v, rad, padd(pk(I.cy), nvs));                       // :81-86
    rv = pfma(pk(I.p2x2), uv, rv);
    rv = pfma(pk(I.p1), Bv, rv);
    if (!second_valid) {
        ru.y = T(0);
        rv.y = T(0);
    }
    if (kWeighted) {
        acc[10] = pfma(w, pfma(ru, ru, pmul(rv, rv)), acc[10]);     // least_squares_utils.py:24-28
        ru = pmul(ru, w);                                           // :43-45
        rv = pmul(rv, w);
    } else {
        acc[10] = pfma(ru, ru, acc[10]);
        acc[10] = pfma(rv, rv, acc[10]);
    }
    const P radp2 = pfma(r2, pfma(r2, pk(I.k3x6), pk(I.k2x4)), pk(I.k1x2));  // 2 d rad / d r2
    P Duu = pfma(uu, radp2, rad);                   // d u'/d u = rad + 2u^2 rad' + 2 p1 v + 6 p2 u
    Duu = pfma(pk(I.p1x2), v, Duu);
    Duu = pfma(pk(I.p2x6), u, Duu);
    P Dvv = pfma(vv, radp2, rad);                   // d v'/d v = rad + 2v^2 rad' + 6 p1 v + 2 p2 u
    Dvv = pfma(pk(I.p1x6), v, Dvv);
    Dvv = pfma(pk(I.p2x2), u, Dvv);
    P Duv = pmul(uv, radp2);                        // d u'/d v = d v'/d u = 2uv rad' + 2 p1 u + 2 p2 v
    Duv = pfma(pk(I.p1x2), u, Duv);
    Duv = pfma(pk(I.p2x2), v, Duv);
    gu = pfma(ru, Duu, pmul(rv, Duv));
    gv = pfma(ru, Duv, pmul(rv, Dvv));
    const P t = pfma(ru, u, pmul(rv, v));
    const P r4 = pmul(r2, r2), r6 = pmul(r4, r2);
    acc[0] = padd(acc[0], ru);
    acc[1] = padd(acc[1], rv);
    acc[2] = pfma(t, r2, acc[2]);
    acc[3] = pfma(t, r4, acc[3]);
    acc[4] = pfma(t, r6, acc[4]);
    acc[5] = pfma(rv, Bv, acc[5]);
    acc[6] = pfma(ru, A, acc[6]);
    acc[11] = pfma(ru, uv, acc[11]);
    acc[12] = pfma(rv, uv, acc[12]);
    acc[7] = pfma(gu, a, acc[7]);
    acc[8] = pfma(gu, b, acc[8]);
    acc[9] = pfma(gv, b, acc[9]);
}

// d/dp1 = sum (2uv ru + Bv rv), d/dp2 = sum (A ru + 2uv rv): add the 2uv parts once per evaluation.
template <typename T>
__device__ __forceinline__ void fold_uv_terms(T (&acc)[kSlots]) {
    acc[5] = fma_t(T(2), acc[11], acc[5]);
    acc[6] = fma_t(T(2), acc[12], acc[6]);
    acc[11] = T(0);
    acc[12] = T(0);
}

// ---- DISTORT10: matches {a, b, u*, v*} staged once per problem into shared memory by bulk TMA ----
// kWeighted is a compile-time switch (the launcher picks the instantiation) so that only one copy of the
// match loop is compiled into a kernel: the solve kernel has to stay inside the instruction cache.
template <typename T, bool kWeighted = false>
struct Distort10Objective {
    using V4 = typename Vec4<T>::type;
    static constexpr bool kUsesSmemMatches = true;

    using P = typename Vec2<T>::type;
    const SolveParams<T>& p;
    V4* matches;      // this warp's shared-memory slab, N + 32 entries (pair layout, see bind)
    T* weights;       // N entries (only if p.has_w)
    T* scratch;       // this warp's kScratch-word transpose area (davo_common.cuh)
    uint64_t* bar;    // this warp's mbarrier
    unsigned parity;
    int lane;

    __device__ Distort10Objective(const SolveParams<T>& p_, unsigned char* slab, T* scratch_, uint64_t* bar_,
                                  int lane_)
        : p(p_), matches(reinterpret_cast<V4*>(slab)),
          weights(reinterpret_cast<T*>(slab + sizeof(V4) * (size_t)(p_.N + 32))), scratch(scratch_), bar(bar_),
          parity(0), lane(lane_) {}

    // bytes of shared memory one warp needs for its slab
    __host__ __device__ static size_t slab_bytes(int N, bool has_w) {
        size_t b = sizeof(V4) * (size_t)(N + 32) + (has_w ? sizeof(T) * (size_t)N : 0);
        return (b + 127) & ~size_t(127);
    }

    __device__ __forceinline__ void bind(int b) {
        // All lanes have finished reading the previous problem's slab (callers sync the warp).
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async();  // order our earlier generic-proxy accesses before the async-proxy write
            const unsigned bytes_m = (unsigned)(sizeof(V4) * (size_t)p.N);
            mbar_expect_tx(bar, bytes_m);
            tma_load_1d(matches, p.data0 + (size_t)b * p.N * 4, bytes_m, bar);
        }
        if (kWeighted) {  // [B,N] rows are not 16-byte aligned for every N: plain coalesced loads
            for (int i = lane; i < p.N; i += 32) weights[i] = p.w[(size_t)b * p.N + i];
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        // Pair layout, once per problem (the slab is then read by every evaluation): this lane's matches i and
        // i + 32 become entry i = {a_i, a_i+32, b_i, b_i+32} and entry i + 32 = -{u*_i, u*_i+32, v*_i, v*_i+32},
        // so one LDS.128 yields two aligned register pairs for the packed arithmetic.  A lane only rewrites its
        // own two entries; a missing second match (ragged N) is zero-filled and masked in eval.
        for (int i = lane; i < p.N; i += 64) {
            const V4 m0 = matches[i];
            V4 m1;
            m1.x = m1.y = m1.z = m1.w = T(0);
            if (i + 32 < p.N) m1 = matches[i + 32];
            V4 ab, ob;
            ab.x = m0.x; ab.y = m1.x; ab.z = m0.y; ab.w = m1.y;
            ob.x = -m0.z; ob.y = -m1.z; ob.z = -m0.w; ob.w = -m1.w;
            matches[i] = ab;
            matches[i + 32] = ob;
        }
        __syncwarp();
    }

    // The match loop.  kRagged = false (N a multiple of 64: every lane has only complete pairs) drops the
    // per-pair validity test and the selects that zero the second match's residuals.
    template <bool kRagged>
    __device__ __forceinline__ void accumulate(const Intrinsics<T>& I, P (&acc2)[kPairAcc]) {
        P gu, gv;
        const int N = p.N;
#pragma unroll 2
        for (int i = lane; i < N; i += 64) {
            const V4 ab = matches[i];
            const V4 ob = matches[i + 32];
            P a, b, nus, nvs, w = pk(T(1));
            a.x = ab.x; a.y = ab.y; b.x = ab.z; b.y = ab.w;
            nus.x = ob.x; nus.y = ob.y; nvs.x = ob.z; nvs.y = ob.w;
            const bool second_valid = !kRagged || i + 32 < N;
            if (kWeighted) {
                w.x = weights[i];
                w.y = second_valid ? weights[i + 32] : T(0);
            }
            match_pair_cost_grad<T, kWeighted>(I, a, b, nus, nvs, w, second_valid, acc2, gu, gv);
        }
    }

    // The match loop again in the reference's association (match_cost_grad_unfused), for an evaluation that
    // overflowed: rolled, rarely executed.
    __device__ __noinline__ void accumulate_unfused(const Intrinsics<T>& I, T (&acc)[kSlots]) {
#pragma unroll
        for (int k = 0; k < kSlots; ++k) acc[k] = T(0);
        T gu, gv;
        const int N = p.N;
#pragma unroll 1
        for (int i = lane; i < N; i += 64) {
            const V4 ab = matches[i];
            const V4 ob = matches[i + 32];
            match_cost_grad_unfused<T, kWeighted>(I, ab.x, ab.z, -ob.x, -ob.z, kWeighted ? weights[i] : T(1), acc, gu, gv);
            if (i + 32 < N)
                match_cost_grad_unfused<T, kWeighted>(I, ab.y, ab.w, -ob.y, -ob.w, kWeighted ? weights[i + 32] : T(1),
                                                      acc, gu, gv);
        }
    }

    // One evaluation.  Inlined at its call site (solve_one_warp has a single one): a __noinline__ version ran
    // 12 % slower (call + stack).
    __device__ __forceinline__ void eval(const T* th, T& f, T& g_own) {
        Intrinsics<T> I;
        I.load(th);
        P acc2[kPairAcc];
#pragma unroll
        for (int k = 0; k < kPairAcc; ++k) acc2[k] = pk(T(0));
        if ((p.N & 63) == 0) accumulate<false>(I, acc2);
        else                 accumulate<true>(I, acc2);
        T acc[kSlots];
#pragma unroll
        for (int k = 0; k < kPairAcc; ++k) acc[k] = acc2[k].x + acc2[k].y;
        fold_uv_terms(acc);  // acc[11], acc[12] are folded into acc[5], acc[6]: 11 sums remain
        // An overflowed evaluation is redone in the reference's association: whether the cost is inf or NaN steers
        // the reference's line search (see match_cost_grad_unfused).
        if (__any_sync(kFull, !isfinite(acc[10]))) accumulate_unfused(I, acc);
        // Sum over the 32 lanes as a transpose through shared memory: lane L stores its 11 partial sums as row
        // L, lane pair c adds up column c (lane 2c the even rows, lane 2c+1 the odd rows) and one shuffle joins
        // the two halves, so lanes 2c and 2c+1 both end with total c.  (The caller's eval_at has synchronised
        // the warp since the previous evaluation read the scratch area.)
        V4* row = reinterpret_cast<V4*>(scratch + lane * kRedPitch);
        V4 q0, q1, q2;
        q0.x = acc[0]; q0.y = acc[1]; q0.z = acc[2]; q0.w = acc[3];
        q1.x = acc[4]; q1.y = acc[5]; q1.z = acc[6]; q1.w = acc[7];
        q2.x = acc[8]; q2.y = acc[9]; q2.z = acc[10]; q2.w = T(0);
        row[0] = q0; row[1] = q1; row[2] = q2;
        __syncwarp();
        const int c = min(lane >> 1, 10), h = lane & 1;
        const T* col = scratch + h * kRedPitch + c;
        T part[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) part[i] = col[2 * i * kRedPitch];
#pragma unroll
        for (int w = 8; w > 0; w >>= 1)
#pragma unroll
            for (int i = 0; i < w; ++i) part[i] += part[i + w];
        const T mine = part[0] + shfl_xor(part[0], 1);
        f = shfl_idx(mine, 20);                            // slot 10 = cost
        g_own = (lane < 20) ? T(2) * mine : T(0);          // least_squares_utils.py:43 (factor 2)
    }
};

// ---- analytic objectives (reference test functions), n <= 16, evaluated redundantly per lane ----
// tests/autograd_solvers/reference_functions.py:20-62, tests/autograd_solvers/test_bfgs_solver.py:33-46,
// tests/autograd_solvers/line_search/test_wolffe_conditions.py:214-305.
template <typename T>
struct AnalyticObjective {
    static constexpr bool kUsesSmemMatches = false;
    const SolveParams<T>& p;
    const T* target;
    int lane;

    __device__ AnalyticObjective(const SolveParams<T>& p_, unsigned char*, T*, uint64_t*, int lane_)
        : p(p_), target(nullptr), lane(lane_) {}
    __host__ __device__ static size_t slab_bytes(int, bool) { return 0; }
    __device__ __forceinline__ void bind(int b) { target = p.data0 ? p.data0 + (size_t)b * p.n : nullptr; }

    __device__ __forceinline__ void eval(const T* th, T& f, T& g_own) {
        const int n = p.n;
        const int c = lane >> 1;
        const bool own = c < n;
        const T xc = own ? th[c] : T(0);
        T ss = T(0);
        for (int j = 0; j < n; ++j) {
            T d = th[j];
            if (p.model == DAVO_MODEL_DISTANCE) d -= target[j];
            ss = add_rn(ss, mul_rn(d, d));
        }
        T g = T(0);
        switch (p.model) {
            case DAVO_MODEL_SPHERE:
                f = ss; g = T(2) * xc; break;
            case DAVO_MODEL_SPHERE_OFFSET:
                f = add_rn(ss, T(10)); g = T(2) * xc; break;
            case DAVO_MODEL_LOG_SPHERE: {
                const T d = add_rn(ss, T(1));
                f = log(d); g = div_rn(T(2) * xc, d); break;
            }
            case DAVO_MODEL_ROSENBROCK: {
                const T x = th[0], y = th[1];
                const T a = sub_rn(T(1), x), bb = sub_rn(y, mul_rn(x, x));
                f = add_rn(mul_rn(a, a), mul_rn(T(100), mul_rn(bb, bb)));
                g = (c == 0) ? sub_rn(mul_rn(T(-2), a), mul_rn(mul_rn(T(400), x), bb)) : mul_rn(T(200), bb);
                break;
            }
            case DAVO_MODEL_COSINE: {
                const T nrm = sqrt_rn(ss);
                const T d = sub_rn(T(1), nrm);
                f = add_rn(sub_rn(T(1), div_rn(th[0], nrm)), mul_rn(d, d));
                const T t = sub_rn(div_rn(mul_rn(th[0], xc), mul_rn(mul_rn(nrm, nrm), nrm)),
                                   (c == 0) ? div_rn(T(1), nrm) : T(0));
                g = sub_rn(t, div_rn(mul_rn(mul_rn(T(2), d), xc), nrm));
                break;
            }
            case DAVO_MODEL_X2_SINE: {
                const T nrm = sqrt_rn(ss);
                const T sn = sin(nrm), cs = cos(nrm);
                f = mul_rn(mul_rn(nrm, nrm), add_rn(sn, T(2)));
                g = mul_rn(add_rn(mul_rn(T(2), add_rn(sn, T(2))), mul_rn(nrm, cs)), xc);
                break;
            }
            case DAVO_MODEL_DISTANCE: {
                const T nrm = sqrt_rn(ss);
                f = nrm;
                // torch's vector_norm backward uses the zero subgradient at the origin
                g = (own && nrm != T(0)) ? div_rn(sub_rn(xc, target[c]), nrm) : T(0);
                break;
            }
            default:
                f = T(NAN); g = T(NAN); break;
        }
        g_own = own ? g : T(0);
    }
};

}  // namespace davo
