// objectives_wide.cuh — the DISTORT10 model and the analytic test objectives behind the interface of the generic
// warp-per-problem kernels (wide_kernel.cuh: parameters and gradient as plain vectors in shared memory).  The eval-mode
// solve of these models runs on the specialised kernels (solver_half.cuh / solver_warp.cuh); the training-mode solve
// and its backward pass (train_kernels.cu) use ONE solver for every model, so they need these adapters.
//
// Forward model and gradient: camera_model/distorted_camera_model.py:59-86 + solvers/least_squares_utils.py:4-48 as in
// objectives.cuh (match_cost_grad); analytic functions: tests/autograd_solvers/reference_functions.py:20-62.
#pragma once
#include "davo_common.cuh"
#include "objectives.cuh"

namespace davo {

// kGlobal = true: the matches are read from global memory at every evaluation instead of being staged into shared
// memory — the route for problems whose N does not fit a warp's slab (N > ~14 000 in float32; the specialised kernels
// stop at ~3 500): slower per evaluation (L2 instead of shared memory), but no N is refused.
template <typename T, bool kGlobal = false>
struct Distort10WideObjective {
    using V4 = typename Vec4<T>::type;
    static constexpr int kParams = 0;
    const SolveParams<T>& p;
    const V4* matches;   // [N] staged {a, b, u*, v*}
    const T* weights;    // [N] (only if p.has_w)
    V4* slab_matches;
    T* slab_weights;
    uint64_t* bar;
    unsigned parity;
    int lane;

    __host__ __device__ static size_t data_bytes(int N, bool has_w) {
        if (kGlobal) return 0;
        size_t b = sizeof(V4) * (size_t)N + (has_w ? sizeof(T) * (size_t)N : 0);
        return (b + 127) & ~size_t(127);
    }
    __host__ __device__ static size_t slab_bytes(int N, int, bool has_w) { return data_bytes(N, has_w) + 16; }

    __device__ Distort10WideObjective(const SolveParams<T>& p_, unsigned char* slab, int lane_)
        : p(p_), matches(nullptr), weights(nullptr), slab_matches(reinterpret_cast<V4*>(slab)),
          slab_weights(reinterpret_cast<T*>(slab + sizeof(V4) * (size_t)(kGlobal ? 0 : p_.N))),
          bar(reinterpret_cast<uint64_t*>(slab + data_bytes(p_.N, p_.has_w != 0))), parity(0), lane(lane_) {}

    __device__ __forceinline__ void init() {
        if (!kGlobal && lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
    }

    __device__ __forceinline__ void bind(int b) {
        __syncwarp();
        if (kGlobal) {
            matches = reinterpret_cast<const V4*>(p.data0) + (size_t)b * p.N;
            weights = p.has_w ? p.w + (size_t)b * p.N : nullptr;
            return;
        }
        if (p.N > 0) {
            if (lane == 0) {
                fence_proxy_async();
                const unsigned bytes = (unsigned)(sizeof(V4) * (size_t)p.N);
                mbar_expect_tx(bar, bytes);
                tma_load_1d(slab_matches, p.data0 + (size_t)b * p.N * 4, bytes, bar);
            }
            if (p.has_w)
                for (int i = lane; i < p.N; i += 32) slab_weights[i] = p.w[(size_t)b * p.N + i];
            mbar_wait(bar, parity);
            parity ^= 1u;
        }
        matches = slab_matches;
        weights = slab_weights;
        __syncwarp();
    }

    __device__ __forceinline__ T eval(const T* th, T* gout) {
        Intrinsics<T> I;
        I.load(th);
        T acc[kSlots];
#pragma unroll
        for (int k = 0; k < kSlots; ++k) acc[k] = T(0);
        T gu, gv;
        if (p.has_w) {
            for (int i = lane; i < p.N; i += 32) {
                const V4 m = matches[i];
                match_cost_grad<T, true>(I, m.x, m.y, m.z, m.w, weights[i], acc, gu, gv);
            }
        } else {
            for (int i = lane; i < p.N; i += 32) {
                const V4 m = matches[i];
                match_cost_grad<T, false>(I, m.x, m.y, m.z, m.w, T(1), acc, gu, gv);
            }
        }
        fold_uv_terms(acc);
        const T mine = reduce_scatter16<true>(acc, lane);
        const T f = shfl_idx(mine, 20);
        __syncwarp();
        if (!(lane & 1) && lane < 20) gout[lane >> 1] = T(2) * mine;  // least_squares_utils.py:43 (factor 2)
        __syncwarp();
        return f;
    }
};

// d loss / d (problem data) of the differentiable solve: the observations of the camera objectives.  x_out depends on
// the data only through the gradients g_k = grad f(x_k; data), so d loss / d data = sum_k d/d data [ g_k . gbar_k ]
// with gbar_k the adjoint of g_k in the reverse sweep (solver_train.cuh) — by symmetry of the mixed second derivative,
// the directional derivative of d f / d data along gbar_k.  Two ways to get it:
//   analytic(): written out (DISTORT10);
//   arm():      the objective's evaluator adds coef * d f / d data to `out` while armed; the four evaluations of the
//               fourth-order difference that already produces the Hessian-vector product then give the directional
//               derivative at no extra evaluation (JOINT, ANGLE_BA).
// Generic objectives: not provided.
template <typename Obj, typename T>
struct DataGradient {
    static constexpr bool kSupported = false;
    __device__ static int elements(const SolveParams<T>&) { return 0; }   // scalars per problem
    __device__ static void analytic(Obj&, const T*, const T*, T*, int) {}
    __device__ static void arm(Obj&, T*, T) {}
};

// DISTORT10, observations: f = sum w |proj_i(x) - obs_i|^2 gives d (g . v) / d obs_i = -2 w_i J_i v with J_i the
// 2 x 10 Jacobian of (u', v') at match i (camera_model/distorted_camera_model.py:59-86 differentiated; SURVEY.md
// Appendix C).  out: this problem's [N, 2] row block, owned by the calling warp.
template <typename T>
struct DataGradient<Distort10WideObjective<T, false>, T> {
    static constexpr bool kSupported = true;
    __device__ static int elements(const SolveParams<T>& p) { return 2 * p.N; }
    __device__ static void arm(Distort10WideObjective<T, false>&, T*, T) {}
    __device__ static void analytic(Distort10WideObjective<T, false>& obj, const T* th, const T* v, T* out, int lane) {
        Intrinsics<T> I;
        I.load(th);
        const T vcx = v[DAVO_CX], vcy = v[DAVO_CY], vk1 = v[DAVO_K1], vk2 = v[DAVO_K2], vk3 = v[DAVO_K3],
                vp1 = v[DAVO_P1], vp2 = v[DAVO_P2], vfx = v[DAVO_FX], vs = v[DAVO_S], vfy = v[DAVO_FY];
        for (int i = lane; i < obj.p.N; i += 32) {
            const typename Vec4<T>::type m = obj.matches[i];
            const T a = m.x, b = m.y;
            const T u = I.fx * a + I.s * b, w_ = I.fy * b;
            const T uu = u * u, vv = w_ * w_, uv = u * w_, r2 = uu + vv, r4 = r2 * r2, r6 = r4 * r2;
            const T rad = T(1) + I.k1 * r2 + I.k2 * r4 + I.k3 * r6;
            const T radp2 = T(2) * I.k1 + T(4) * I.k2 * r2 + T(6) * I.k3 * r4;  // 2 d rad / d r2
            const T Duu = rad + uu * radp2 + T(2) * I.p1 * w_ + T(6) * I.p2 * u;
            const T Dvv = rad + vv * radp2 + T(6) * I.p1 * w_ + T(2) * I.p2 * u;
            const T Duv = uv * radp2 + T(2) * I.p1 * u + T(2) * I.p2 * w_;
            const T du = vfx * a + vs * b, dv = vfy * b;  // d(u, v) along v
            const T dk = vk1 * r2 + vk2 * r4 + vk3 * r6;
            const T Ju = vcx + u * dk + T(2) * uv * vp1 + (r2 + T(2) * uu) * vp2 + Duu * du + Duv * dv;
            const T Jv = vcy + w_ * dk + (r2 + T(2) * vv) * vp1 + T(2) * uv * vp2 + Duv * du + Dvv * dv;
            const T wt = obj.p.has_w ? obj.weights[i] : T(1);
            out[2 * i] -= T(2) * wt * Ju;
            out[2 * i + 1] -= T(2) * wt * Jv;
        }
    }
};

// The analytic objectives (n <= 16): AnalyticObjective's arithmetic, gradient written out as a vector.
template <typename T>
struct AnalyticWideObjective {
    static constexpr int kParams = 0;
    AnalyticObjective<T> inner;
    int lane;
    __host__ __device__ static size_t slab_bytes(int, int, bool) { return 0; }
    __device__ AnalyticWideObjective(const SolveParams<T>& p_, unsigned char*, int lane_)
        : inner(p_, nullptr, nullptr, nullptr, lane_), lane(lane_) {}
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ void bind(int b) { inner.bind(b); }
    __device__ __forceinline__ T eval(const T* th, T* gout) {
        T f, g_own;
        inner.eval(th, f, g_own);
        __syncwarp();
        if (!(lane & 1) && (lane >> 1) < inner.p.n) gout[lane >> 1] = g_own;
        __syncwarp();
        return f;
    }
};

}  // namespace davo
