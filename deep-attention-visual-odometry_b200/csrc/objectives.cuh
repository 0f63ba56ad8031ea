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

// One match: residuals, squared error and the intrinsic-gradient accumulators.
//   acc[0..9]  += d(cost/2)/d(cx,cy,k1,k2,k3,p1,p2,fx,s,fy)   (acc[5], acc[6] without their 2uv terms)
//   acc[10]    += cost
//   acc[11]    += ru * uv,  acc[12] += rv * uv                 (folded into acc[5], acc[6] by fold_uv_terms)
// (gu, gv) = d(cost/2)/d(u, v) are returned for the pose chain rule of the JOINT model.
// 53 FP32 instructions; counted as 93 flop by SURVEY.md 8(d).
template <typename T, bool kWeighted>
__device__ __forceinline__ void match_cost_grad(const Intrinsics<T>& I, T a, T b, T us, T vs, T w,
                                                T (&acc)[kSlots], T& gu, T& gv) {
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

    const SolveParams<T>& p;
    V4* matches;      // this warp's shared-memory slab, N entries
    T* weights;       // N entries (only if p.has_w)
    uint64_t* bar;    // this warp's mbarrier
    unsigned parity;
    int lane;

    __device__ Distort10Objective(const SolveParams<T>& p_, unsigned char* slab, uint64_t* bar_, int lane_)
        : p(p_), matches(reinterpret_cast<V4*>(slab)),
          weights(reinterpret_cast<T*>(slab + sizeof(V4) * (size_t)p_.N)), bar(bar_), parity(0), lane(lane_) {}

    // bytes of shared memory one warp needs for its slab
    __host__ __device__ static size_t slab_bytes(int N, bool has_w) {
        size_t b = sizeof(V4) * (size_t)N + (has_w ? sizeof(T) * (size_t)N : 0);
        return (b + 127) & ~size_t(127);
    }

    __device__ __forceinline__ void bind(int b) {
        // All lanes have finished reading the previous problem's slab (callers sync the warp).
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async();  // order our earlier generic-proxy reads before the async-proxy write
            const unsigned bytes_m = (unsigned)(sizeof(V4) * (size_t)p.N);
            mbar_expect_tx(bar, bytes_m);
            tma_load_1d(matches, p.data0 + (size_t)b * p.N * 4, bytes_m, bar);
        }
        if (kWeighted) {  // [B,N] rows are not 16-byte aligned for every N: plain coalesced loads
            for (int i = lane; i < p.N; i += 32) weights[i] = p.w[(size_t)b * p.N + i];
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        __syncwarp();
    }

    // One evaluation.  Inlined at its three call sites (outer iteration, line-search probe, final cost): a
    // __noinline__ version shrank the kernel from 57 KB to 27 KB of SASS but ran 12 % slower (call + stack).
    __device__ __forceinline__ void eval(const T* th, T& f, T& g_own) {
        Intrinsics<T> I;
        I.load(th);
        T acc[kSlots];
#pragma unroll
        for (int k = 0; k < kSlots; ++k) acc[k] = T(0);
        T gu, gv;
        const int N = p.N;
#pragma unroll 4
        for (int i = lane; i < N; i += 32) {
            const V4 m = matches[i];
            match_cost_grad<T, kWeighted>(I, m.x, m.y, m.z, m.w, kWeighted ? weights[i] : T(1), acc, gu, gv);
        }
        fold_uv_terms(acc);
        const T mine = reduce_scatter16<true>(acc, lane);  // slot c total in lanes 2c, 2c+1
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

    __device__ AnalyticObjective(const SolveParams<T>& p_, unsigned char*, uint64_t*, int lane_)
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
