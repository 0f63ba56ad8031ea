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

template <typename T>
struct Distort10WideObjective {
    using V4 = typename Vec4<T>::type;
    static constexpr int kParams = 0;
    const SolveParams<T>& p;
    V4* matches;   // [N] staged {a, b, u*, v*}
    T* weights;    // [N] (only if p.has_w)
    uint64_t* bar;
    unsigned parity;
    int lane;

    __host__ __device__ static size_t data_bytes(int N, bool has_w) {
        size_t b = sizeof(V4) * (size_t)N + (has_w ? sizeof(T) * (size_t)N : 0);
        return (b + 127) & ~size_t(127);
    }
    __host__ __device__ static size_t slab_bytes(int N, int, bool has_w) { return data_bytes(N, has_w) + 16; }

    __device__ Distort10WideObjective(const SolveParams<T>& p_, unsigned char* slab, int lane_)
        : p(p_), matches(reinterpret_cast<V4*>(slab)),
          weights(reinterpret_cast<T*>(slab + sizeof(V4) * (size_t)p_.N)),
          bar(reinterpret_cast<uint64_t*>(slab + data_bytes(p_.N, p_.has_w != 0))), parity(0), lane(lane_) {}

    __device__ __forceinline__ void init() {
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
    }

    __device__ __forceinline__ void bind(int b) {
        __syncwarp();
        if (p.N > 0) {
            if (lane == 0) {
                fence_proxy_async();
                const unsigned bytes = (unsigned)(sizeof(V4) * (size_t)p.N);
                mbar_expect_tx(bar, bytes);
                tma_load_1d(matches, p.data0 + (size_t)b * p.N * 4, bytes, bar);
            }
            if (p.has_w)
                for (int i = lane; i < p.N; i += 32) weights[i] = p.w[(size_t)b * p.N + i];
            mbar_wait(bar, parity);
            parity ^= 1u;
        }
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
