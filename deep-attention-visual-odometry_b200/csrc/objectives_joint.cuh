// objectives_joint.cuh — the JOINT objective: 10 intrinsics + (rx,ry,rz,tx,ty,tz) per view, V views
// of N shared world points.  One warp evaluates one problem; world points [N,3] and observations
// [V,N,2] are staged once per problem into this warp's shared memory by bulk TMA and reused by every
// evaluation of every line search.
//
// Forward model: camera_model/distorted_camera_model.py:24-103 evaluated per view (Euler R = Rz Ry Rx,
// X' = R X + t, z' == 0 -> += 1e-8, perspective divide, intrinsics + distortion); cost: squared
// residuals summed over views (solvers/least_squares_utils.py:4-28).  Gradient:
// intrinsics as in DISTORT10; per view gX' = d(cost/2)/dX' accumulated into g_t = sum gX' and
// M = sum gX' (x) X, then d/dr_k = <dR/dr_k, M>_F after the warp reduction.
#pragma once
#include "davo_common.cuh"
#include "objectives.cuh"

namespace davo {

constexpr int kMaxViews = 19;  // n = 10 + 6 V <= 124 (the wide solver's 128); the CTA-per-problem solve takes V <= 9

// kDataGrad (the backward pass of the differentiable solve only, solver_train.cuh): while `dgrad` is set every
// evaluation also adds dcoef * d cost / d obs to dgrad[V,N,2] — pair (view, i) is always visited by the same lane.
template <typename T, bool kDataGrad = false>
struct JointObjective {
    static constexpr int kParams = 0;  // n is a run-time value (wide_kernel.cuh)
    T* dgrad = nullptr;
    T dcoef = T(0);
    const SolveParams<T>& p;
    T* world;   // [N,3]
    T* obs;     // [V,N,2]
    T* wts;     // [V,N] (only if p.has_w)
    T* red;     // 16-entry reduction line
    T* sc;      // sin/cos of the 3V Euler angles: [3V][2]
    uint64_t* bar;
    unsigned parity;
    int lane;

    __host__ __device__ static size_t data_bytes(int N, int V, bool has_w) {
        size_t b = sizeof(T) * ((size_t)N * 3 + (size_t)V * N * 2 + (has_w ? (size_t)V * N : 0));
        return (b + 127) & ~size_t(127);
    }
    __host__ __device__ static size_t slab_bytes(int N, int V, bool has_w) {
        return data_bytes(N, V, has_w) + sizeof(T) * (16 + 128) + 16;
    }

    __device__ JointObjective(const SolveParams<T>& p_, unsigned char* slab, int lane_) : p(p_), parity(0), lane(lane_) {
        world = reinterpret_cast<T*>(slab);
        obs = world + (size_t)p.N * 3;
        wts = obs + (size_t)p.V * p.N * 2;
        unsigned char* tail = slab + data_bytes(p.N, p.V, p.has_w != 0);
        red = reinterpret_cast<T*>(tail);
        sc = red + 16;
        bar = reinterpret_cast<uint64_t*>(sc + 128);
    }

    __device__ __forceinline__ void init() {
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
    }

    __device__ __forceinline__ void bind(int b) {
        __syncwarp();
        const size_t wb = sizeof(T) * (size_t)p.N * 3, ob = sizeof(T) * (size_t)p.V * p.N * 2;
        const T* gw = p.data0 + (size_t)b * p.N * 3;
        const T* go = p.data1 + (size_t)b * p.V * p.N * 2;
        const bool bulk = (wb % 16 == 0) && (ob % 16 == 0);  // rows stay 16-byte aligned for every b
        if (bulk) {
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(bar, (unsigned)(wb + ob));
                tma_load_1d(world, gw, (unsigned)wb, bar);
                tma_load_1d(obs, go, (unsigned)ob, bar);
            }
        } else {
            for (int i = lane; i < p.N * 3; i += 32) world[i] = gw[i];
            for (int i = lane; i < p.V * p.N * 2; i += 32) obs[i] = go[i];
        }
        if (p.has_w)
            for (int i = lane; i < p.V * p.N; i += 32) wts[i] = p.w[(size_t)b * p.V * p.N + i];
        if (bulk) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        }
        __syncwarp();
    }

    // th: parameters in shared memory; gout: gradient written to shared memory; returns the cost.
    __device__ __forceinline__ T eval(const T* th, T* gout) {
        const int N = p.N, V = p.V;
        Intrinsics<T> I;
        I.load(th);
        __syncwarp();
        for (int a = lane; a < 3 * V; a += 32) {
            const T ang = th[10 + 6 * (a / 3) + (a % 3)];
            sc[2 * a] = sin(ang);
            sc[2 * a + 1] = cos(ang);
        }
        __syncwarp();
        T acc[kSlots];
#pragma unroll
        for (int k = 0; k < kSlots; ++k) acc[k] = T(0);
        for (int v = 0; v < V; ++v) {
            const T sx = sc[6 * v], cx = sc[6 * v + 1], sy = sc[6 * v + 2], cy = sc[6 * v + 3], sz = sc[6 * v + 4],
                    cz = sc[6 * v + 5];
            const T* pose = th + 10 + 6 * v;
            const T t0 = pose[3], t1 = pose[4], t2 = pose[5];
            // R = Rz Ry Rx, distorted_camera_model.py:38-55
            const T r00 = cy * cz, r01 = sx * sy * cz - cx * sz, r02 = cx * sy * cz + sx * sz;
            const T r10 = cy * sz, r11 = sx * sy * sz + cx * cz, r12 = cx * sy * sz - sx * cz;
            const T r20 = -sy, r21 = sx * cy, r22 = cx * cy;
            T part[kSlots];  // 0..8: M = sum gX' (x) X (row major), 9..11: g_t = sum gX'
#pragma unroll
            for (int k = 0; k < kSlots; ++k) part[k] = T(0);
            const T* ob = obs + (size_t)v * N * 2;
            const T* wv = wts + (size_t)v * N;
#pragma unroll 2
            for (int i = lane; i < N; i += 32) {
                const T X = world[3 * i], Y = world[3 * i + 1], Z = world[3 * i + 2];
                const T us = ob[2 * i], vs = ob[2 * i + 1];
                const T xp = fma_t(X, r00, fma_t(Y, r01, fma_t(Z, r02, t0)));
                const T yp = fma_t(X, r10, fma_t(Y, r11, fma_t(Z, r12, t1)));
                T zp = fma_t(X, r20, fma_t(Y, r21, fma_t(Z, r22, t2)));
                if (zp == T(0)) zp += T(1e-8);  // :57
                const T iz = div_rn(T(1), zp);
                const T a = xp * iz, b = yp * iz;
                T gu, gv;
                if (p.has_w) match_cost_grad<T, true>(I, a, b, us, vs, wv[i], acc, gu, gv);
                else         match_cost_grad<T, false>(I, a, b, us, vs, T(1), acc, gu, gv);
                if (kDataGrad && dgrad) {  // d cost / d (u*, v*) = -2 w (residual)
                    T ru, rv;
                    match_residual(I, a, b, us, vs, p.has_w ? wv[i] : T(1), ru, rv);
                    T* o = dgrad + 2 * ((size_t)v * N + i);
                    o[0] -= T(2) * dcoef * ru;
                    o[1] -= T(2) * dcoef * rv;
                }
                const T gA = gu * I.fx;
                const T gB = fma_t(gu, I.s, gv * I.fy);
                const T gx = gA * iz, gy = gB * iz;
                const T gz = -fma_t(gA, a, gB * b) * iz;
                part[0] = fma_t(gx, X, part[0]); part[1] = fma_t(gx, Y, part[1]); part[2] = fma_t(gx, Z, part[2]);
                part[3] = fma_t(gy, X, part[3]); part[4] = fma_t(gy, Y, part[4]); part[5] = fma_t(gy, Z, part[5]);
                part[6] = fma_t(gz, X, part[6]); part[7] = fma_t(gz, Y, part[7]); part[8] = fma_t(gz, Z, part[8]);
                part[9] += gx; part[10] += gy; part[11] += gz;
            }
            const T mine = reduce_scatter16<true>(part, lane);
            __syncwarp();
            if (!(lane & 1)) red[lane >> 1] = mine;
            __syncwarp();
            if (lane < 6) {
                T out;
                if (lane >= 3) {
                    out = red[9 + (lane - 3)];  // d/dt = sum gX'
                } else {
                    // <dR/dr_k, M>_F with dR/dr_k of R = Rz Ry Rx
                    T d0, d1, d2, d3, d4, d5, d6, d7, d8;
                    if (lane == 0) {
                        d0 = T(0); d1 = cx * sy * cz + sx * sz;  d2 = -sx * sy * cz + cx * sz;
                        d3 = T(0); d4 = cx * sy * sz - sx * cz;  d5 = -sx * sy * sz - cx * cz;
                        d6 = T(0); d7 = cx * cy;                 d8 = -sx * cy;
                    } else if (lane == 1) {
                        d0 = -sy * cz; d1 = sx * cy * cz; d2 = cx * cy * cz;
                        d3 = -sy * sz; d4 = sx * cy * sz; d5 = cx * cy * sz;
                        d6 = -cy;      d7 = -sx * sy;     d8 = -cx * sy;
                    } else {
                        d0 = -cy * sz; d1 = -sx * sy * sz - cx * cz; d2 = -cx * sy * sz + sx * cz;
                        d3 = cy * cz;  d4 = sx * sy * cz - cx * sz;  d5 = cx * sy * cz + sx * sz;
                        d6 = T(0); d7 = T(0); d8 = T(0);
                    }
                    out = d0 * red[0] + d1 * red[1] + d2 * red[2] + d3 * red[3] + d4 * red[4] + d5 * red[5] +
                          d6 * red[6] + d7 * red[7] + d8 * red[8];
                }
                gout[10 + 6 * v + lane] = T(2) * out;  // least_squares_utils.py:43
            }
        }
        fold_uv_terms(acc);
        const T mine = reduce_scatter16<true>(acc, lane);
        const T f = shfl_idx(mine, 20);
        if (!(lane & 1) && lane < 20) gout[lane >> 1] = T(2) * mine;
        __syncwarp();
        return f;
    }
};

}  // namespace davo
