// elementwise_kernels.cu — the HBM-bound pieces either side of the solve: staging of matches, the
// 16-parameter forward model (+ Jacobian), explicit least-squares reductions and the stand-alone
// BFGS update.  One CTA per problem, threads over matches (coalesced), per-problem constants in
// shared memory.
//
// Reference lines (relative to /root/reference/deep_attention_visual_odometry/):
//   camera_model/distorted_camera_model.py:24-103,106-111   forward model
//   camera_model/distorted_camera_model.py:114-385          Jacobian layout [B,2N,16] (values re-derived)
//   solvers/least_squares_utils.py:16-48                    find_error / find_error_gradient
//   autograd_solvers/bfgs_solver.py:217-233, 235-303        eq. 6.20 / eq. 6.17
#include "davo_common.cuh"
#include "launch.h"

namespace davo {

// R = Rz Ry Rx (distorted_camera_model.py:29-55) and optionally dR/drx, dR/dry, dR/drz.
template <typename T>
__device__ void euler_matrices(T rx, T ry, T rz, T* Rm, T* dRx, T* dRy, T* dRz) {
    const T sx = sin(rx), cx = cos(rx), sy = sin(ry), cy = cos(ry), sz = sin(rz), cz = cos(rz);
    Rm[0] = cy * cz; Rm[1] = sx * sy * cz - cx * sz; Rm[2] = cx * sy * cz + sx * sz;
    Rm[3] = cy * sz; Rm[4] = sx * sy * sz + cx * cz; Rm[5] = cx * sy * sz - sx * cz;
    Rm[6] = -sy;     Rm[7] = sx * cy;                Rm[8] = cx * cy;
    if (dRx) {
        dRx[0] = 0; dRx[1] = cx * sy * cz + sx * sz;  dRx[2] = -sx * sy * cz + cx * sz;
        dRx[3] = 0; dRx[4] = cx * sy * sz - sx * cz;  dRx[5] = -sx * sy * sz - cx * cz;
        dRx[6] = 0; dRx[7] = cx * cy;                 dRx[8] = -sx * cy;
        dRy[0] = -sy * cz; dRy[1] = sx * cy * cz; dRy[2] = cx * cy * cz;
        dRy[3] = -sy * sz; dRy[4] = sx * cy * sz; dRy[5] = cx * cy * sz;
        dRy[6] = -cy;      dRy[7] = -sx * sy;     dRy[8] = -cx * sy;
        dRz[0] = -cy * sz; dRz[1] = -sx * sy * sz - cx * cz; dRz[2] = -cx * sy * sz + sx * cz;
        dRz[3] = cy * cz;  dRz[4] = sx * sy * cz - cx * sz;  dRz[5] = cx * sy * cz + sx * sz;
        dRz[6] = 0; dRz[7] = 0; dRz[8] = 0;
    }
}

// X' = R X + t as un-fused left-to-right sums of separately rounded products; z' == 0 -> += 1e-8 (:38-57).
// The entries of R are rounded first and then multiplied by the coordinate.  The reference multiplies the
// coordinate by each trigonometric factor in turn in four places ((X cy) cz, (X cy) sz, (Y sx) cy, (Z cx) cy,
// :39,:45,:52-53), so a float32 X' can differ from the reference's in the last bit: parity of the staged matches
// and of the forward model is a TOLERANCE (tests: rtol 3e-6 / 2e-5 in float32, 1e-12 in float64), not bitwise,
// except for the identity pose, where every product is exact.
template <typename T>
__device__ __forceinline__ void transform_point(const T* Rm, const T* t, T X, T Y, T Z, T& xp, T& yp, T& zp) {
    xp = add_rn(add_rn(add_rn(mul_rn(X, Rm[0]), mul_rn(Y, Rm[1])), mul_rn(Z, Rm[2])), t[0]);
    yp = add_rn(add_rn(add_rn(mul_rn(X, Rm[3]), mul_rn(Y, Rm[4])), mul_rn(Z, Rm[5])), t[1]);
    zp = add_rn(add_rn(add_rn(mul_rn(X, Rm[6]), mul_rn(Y, Rm[7])), mul_rn(Z, Rm[8])), t[2]);
    if (zp == T(0)) zp += T(1e-8);
}

// ---- staging: {x'/z', y'/z', u*, v*} per match ---------------------------------------------------
// General pose.  A CTA takes kStageGroup problems per trip: kStageGroup threads build the rotation matrices (one
// problem each, in parallel — a single thread doing the trigonometry for 255 waiting ones kept this kernel at 41 %
// of the HBM rate), then all threads stream the group's matches, kStageUnroll per thread with every load issued
// before the first use.  20 B read + 16 B written per match in float32.
constexpr int kStageGroup = 8;
constexpr int kStageUnroll = 4;

template <typename T>
__global__ void __launch_bounds__(256) stage_kernel(int B, int N, const T* __restrict__ pts,
                                                    const T* __restrict__ obs, const T* __restrict__ pose,
                                                    T* __restrict__ staged) {
    using V4 = typename Vec4<T>::type;
    __shared__ T Rm[kStageGroup][12];  // 9 rotation entries + translation
    for (long long g0 = (long long)blockIdx.x * kStageGroup; g0 < B; g0 += (long long)gridDim.x * kStageGroup) {
        const int nb = (int)((B - g0) < kStageGroup ? (B - g0) : kStageGroup);
        __syncthreads();
        if ((int)threadIdx.x < nb) {
            const T* ps = pose + 6 * (size_t)(g0 + threadIdx.x);
            T* R = Rm[threadIdx.x];
            euler_matrices<T>(ps[0], ps[1], ps[2], R, nullptr, nullptr, nullptr);
            R[9] = ps[3]; R[10] = ps[4]; R[11] = ps[5];
        }
        __syncthreads();
        const long long first = g0 * N;                  // first match of the group in the flat [B*N] arrays
        const int total = nb * N;
        for (int base = threadIdx.x; base < total; base += 256 * kStageUnroll) {
            T X[kStageUnroll], Y[kStageUnroll], Z[kStageUnroll], U[kStageUnroll], Vv[kStageUnroll];
#pragma unroll
            for (int k = 0; k < kStageUnroll; ++k) {
                const int i = base + 256 * k;
                if (i < total) {
                    const long long m = first + i;
                    X[k] = __ldg(pts + 3 * m); Y[k] = __ldg(pts + 3 * m + 1); Z[k] = __ldg(pts + 3 * m + 2);
                    U[k] = __ldg(obs + 2 * m); Vv[k] = __ldg(obs + 2 * m + 1);
                }
            }
#pragma unroll
            for (int k = 0; k < kStageUnroll; ++k) {
                const int i = base + 256 * k;
                if (i < total) {
                    const T* R = Rm[i / N];
                    T xp, yp, zp;
                    transform_point<T>(R, R + 9, X[k], Y[k], Z[k], xp, yp, zp);
                    V4 out;
                    out.x = div_rn(xp, zp);
                    out.y = div_rn(yp, zp);
                    out.z = U[k];
                    out.w = Vv[k];
                    reinterpret_cast<V4*>(staged)[first + i] = out;
                }
            }
        }
    }
}

// Identity pose (BASELINE configs 2, 4, 5): no per-problem constants, so the [B*N] matches are one flat
// array.  Each thread stages kStageUnroll matches a grid-stride apart and issues all of their loads
// before the first use (20 B x 4 in flight per thread) — the kernel is pure HBM streaming:
// 20 B read + 16 B written per match in float32.

template <typename T>
__global__ void __launch_bounds__(256) stage_identity_kernel(long long total, const T* __restrict__ pts,
                                                             const T* __restrict__ obs, T* __restrict__ staged) {
    using V4 = typename Vec4<T>::type;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < total;
         base += stride * kStageUnroll) {
        T X[kStageUnroll], Y[kStageUnroll], Z[kStageUnroll], U[kStageUnroll], Vv[kStageUnroll];
#pragma unroll
        for (int k = 0; k < kStageUnroll; ++k) {
            const long long i = base + k * stride;
            if (i < total) {
                X[k] = __ldg(pts + 3 * i); Y[k] = __ldg(pts + 3 * i + 1); Z[k] = __ldg(pts + 3 * i + 2);
                U[k] = __ldg(obs + 2 * i); Vv[k] = __ldg(obs + 2 * i + 1);
            }
        }
#pragma unroll
        for (int k = 0; k < kStageUnroll; ++k) {
            const long long i = base + k * stride;
            if (i < total) {
                // R = I, t = 0 evaluated with the reference's sums (x*1 + y*0 + z*0 + 0 is exact); z' == 0 -> 1e-8
                T zp = Z[k];
                if (zp == T(0)) zp += T(1e-8);
                V4 out;
                out.x = div_rn(X[k], zp);
                out.y = div_rn(Y[k], zp);
                out.z = U[k];
                out.w = Vv[k];
                reinterpret_cast<V4*>(staged)[i] = out;
            }
        }
    }
}

template <typename T>
int launch_stage(int B, int N, const T* pts, const T* obs, const T* pose, T* staged, cudaStream_t s) {
    if (B == 0 || N == 0) return DAVO_OK;
    if (!pose) {
        const long long total = (long long)B * N;
        long long blocks = (total + 256LL * kStageUnroll - 1) / (256LL * kStageUnroll);
        if (blocks > 148LL * 64) blocks = 148LL * 64;
        stage_identity_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(total, pts, obs, staged);
    } else {
        const long long groups = ((long long)B + kStageGroup - 1) / kStageGroup;
        const int grid = (int)(groups < 148LL * 16 ? groups : 148LL * 16);
        stage_kernel<T><<<grid, 256, 0, s>>>(B, N, pts, obs, pose, staged);
    }
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}
template int launch_stage<float>(int, int, const float*, const float*, const float*, float*, cudaStream_t);
template int launch_stage<double>(int, int, const double*, const double*, const double*, double*, cudaStream_t);

// ---- 16-parameter forward model and Jacobian ------------------------------------------------------
// A CTA takes kProjGroup problems per trip: kProjGroup threads load the 16 parameters and build R (and dR/dr for
// the Jacobian) of one problem each, in parallel, then all threads stream the group's matches flat, kProjUnroll
// per thread with the point loads issued before the first use (one thread doing the trigonometry for a whole
// CTA, one match per thread, kept this kernel at a third of the HBM rate).
// (The Jacobian variant writes 136 B per match and does ~10x the arithmetic: there one problem per trip with two
// matches per thread is faster — 81 us against 103 us for 8K x 256 — so it keeps a group of 1.)
template <bool kJac>
struct ProjGroup { static constexpr int value = kJac ? 1 : 8; };

template <typename T, bool kJac>
__global__ void __launch_bounds__(128) project_kernel(int B, int N, const T* __restrict__ pts,
                                                      const T* __restrict__ th16, T* __restrict__ u_out,
                                                      T* __restrict__ v_out, T* __restrict__ J) {
    constexpr int kProjUnroll = kJac ? 1 : 4;
    constexpr int kProjGroup = ProjGroup<kJac>::value;
    __shared__ T ths[kProjGroup][16];
    __shared__ T Rms[kProjGroup][9], dRs[kProjGroup][3][9];
    for (long long g0 = (long long)blockIdx.x * kProjGroup; g0 < B; g0 += (long long)gridDim.x * kProjGroup) {
        const int nb = (int)((B - g0) < kProjGroup ? (B - g0) : kProjGroup);
        __syncthreads();
        if ((int)threadIdx.x < 16 * nb) ths[threadIdx.x >> 4][threadIdx.x & 15] = th16[16 * (size_t)g0 + threadIdx.x];
        __syncthreads();
        if ((int)threadIdx.x < nb) {
            const int q = threadIdx.x;
            euler_matrices<T>(ths[q][DAVO_RX], ths[q][DAVO_RY], ths[q][DAVO_RZ], Rms[q], kJac ? dRs[q][0] : nullptr,
                              dRs[q][1], dRs[q][2]);
        }
        __syncthreads();
        const long long first = g0 * N;
        const int total = nb * N;
        for (int base = threadIdx.x; base < total; base += 128 * kProjUnroll) {
            T Xs[kProjUnroll][3];
#pragma unroll
            for (int k = 0; k < kProjUnroll; ++k) {
                const int i = base + 128 * k;
                if (i < total) {
                    const T* X = pts + 3 * (first + i);
                    Xs[k][0] = __ldg(X); Xs[k][1] = __ldg(X + 1); Xs[k][2] = __ldg(X + 2);
                }
            }
#pragma unroll
            for (int k = 0; k < kProjUnroll; ++k) {
                const int i = base + 128 * k;
                if (i >= total) continue;
                const int q = i / N, m = i - q * N;
                const long long b = g0 + q, mi = first + i;
                const T* th = ths[q];
                const T* Rm = Rms[q];
                const T (*dR)[9] = dRs[q];
                const T cx = th[0], cy = th[1], k1 = th[2], k2 = th[3], k3 = th[4], p1 = th[5], p2 = th[6], fx = th[7],
                        sk = th[8], fy = th[9];
                const T X0 = Xs[k][0], X1 = Xs[k][1], X2 = Xs[k][2];
                T xp, yp, zp;
                transform_point<T>(Rm, th + DAVO_TX, X0, X1, X2, xp, yp, zp);
                // un-fused, in the reference's order, so the float32 forward agrees to the last bit or two
                const T a = div_rn(xp, zp), bb = div_rn(yp, zp);
                const T u = add_rn(mul_rn(fx, a), mul_rn(sk, bb));                   // :59-61
                const T v = mul_rn(fy, bb);                                          // :62
                const T r2 = add_rn(mul_rn(u, u), mul_rn(v, v));                     // :64
                const T uv = mul_rn(u, v);                                           // :65
                const T rad = add_rn(add_rn(add_rn(T(1), mul_rn(k1, r2)), mul_rn(mul_rn(k2, r2), r2)),
                                     mul_rn(mul_rn(mul_rn(k3, r2), r2), r2));        // :66-74
                const T A = add_rn(r2, mul_rn(mul_rn(T(2), u), u));
                const T Bv = add_rn(r2, mul_rn(mul_rn(T(2), v), v));
                const T up = add_rn(add_rn(add_rn(mul_rn(u, rad), mul_rn(mul_rn(T(2), p1), uv)), mul_rn(p2, A)), cx);
                const T vp = add_rn(add_rn(add_rn(mul_rn(v, rad), mul_rn(mul_rn(T(2), p2), uv)), mul_rn(p1, Bv)), cy);
                u_out[mi] = up;
                v_out[mi] = vp;
                if (kJac) {
                    const T r4 = r2 * r2, r6 = r4 * r2;
                    const T radp = k1 + T(2) * k2 * r2 + T(3) * k3 * r4;
                    const T Duu = rad + T(2) * u * u * radp + T(2) * p1 * v + T(6) * p2 * u;
                    const T Dvv = rad + T(2) * v * v * radp + T(6) * p1 * v + T(2) * p2 * u;
                    const T Duv = T(2) * uv * radp + T(2) * p1 * u + T(2) * p2 * v;
                    const T iz = T(1) / zp;
                    T ju[16], jv[16];
                    ju[0] = 1; jv[0] = 0;
                    ju[1] = 0; jv[1] = 1;
                    ju[2] = u * r2; jv[2] = v * r2;
                    ju[3] = u * r4; jv[3] = v * r4;
                    ju[4] = u * r6; jv[4] = v * r6;
                    ju[5] = T(2) * uv; jv[5] = Bv;
                    ju[6] = A; jv[6] = T(2) * uv;
                    ju[7] = Duu * a;  jv[7] = Duv * a;
                    ju[8] = Duu * bb; jv[8] = Duv * bb;
                    ju[9] = Duv * bb; jv[9] = Dvv * bb;
                    const T ux = Duu * fx * iz, uy = (Duu * sk + Duv * fy) * iz, uz = -(Duu * u + Duv * v) * iz;
                    const T vx = Duv * fx * iz, vy = (Duv * sk + Dvv * fy) * iz, vz = -(Duv * u + Dvv * v) * iz;
    #pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const T dX = dR[k][0] * X0 + dR[k][1] * X1 + dR[k][2] * X2;
                        const T dY = dR[k][3] * X0 + dR[k][4] * X1 + dR[k][5] * X2;
                        const T dZ = dR[k][6] * X0 + dR[k][7] * X1 + dR[k][8] * X2;
                        ju[10 + k] = ux * dX + uy * dY + uz * dZ;
                        jv[10 + k] = vx * dX + vy * dY + vz * dZ;
                    }
                    ju[13] = ux; ju[14] = uy; ju[15] = uz;
                    jv[13] = vx; jv[14] = vy; jv[15] = vz;
                    using V4 = typename Vec4<T>::type;
                    V4* Ju = reinterpret_cast<V4*>(J + 16 * ((size_t)b * 2 * N + m));          // rows 0..N-1: u'
                    V4* Jv = reinterpret_cast<V4*>(J + 16 * ((size_t)b * 2 * N + N + m));      // rows N..2N-1: v'
    #pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        V4 a4, b4;
                        a4.x = ju[4 * q]; a4.y = ju[4 * q + 1]; a4.z = ju[4 * q + 2]; a4.w = ju[4 * q + 3];
                        b4.x = jv[4 * q]; b4.y = jv[4 * q + 1]; b4.z = jv[4 * q + 2]; b4.w = jv[4 * q + 3];
                        Ju[q] = a4;
                        Jv[q] = b4;
                    }
                }
            }
        }
    }
}

template <typename T>
int launch_project(int B, int N, const T* pts, const T* th16, T* u, T* v, T* J, cudaStream_t s) {
    if (B == 0 || N == 0) return DAVO_OK;
    const int group = J ? ProjGroup<true>::value : ProjGroup<false>::value;
    const long long groups = ((long long)B + group - 1) / group;
    const int grid = (int)(groups < 148LL * 64 ? groups : 148LL * 64);
    if (J) project_kernel<T, true><<<grid, 128, 0, s>>>(B, N, pts, th16, u, v, J);
    else   project_kernel<T, false><<<grid, 128, 0, s>>>(B, N, pts, th16, u, v, nullptr);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}
template int launch_project<float>(int, int, const float*, const float*, float*, float*, float*, cudaStream_t);
template int launch_project<double>(int, int, const double*, const double*, double*, double*, double*, cudaStream_t);

// ---- explicit least squares: error[B] = sum w r^2, gradient[B,P] = sum 2 w r J --------------------
// One WARP per problem (8 per CTA, grid-stride): lanes stride over the R residuals with several loads in flight,
// per-lane partial sums, one butterfly per output.  (A 256-thread CTA per problem with an 8-stage block reduction
// ran at a fifth of the HBM rate at R = 512.)
constexpr int kLsqThreads = 256;
constexpr int kLsqCols = 16;  // gradient columns accumulated per pass over the residuals

template <typename T>
__device__ __forceinline__ T lsq_warp_sum(T v) {
    v += shfl_xor(v, 16); v += shfl_xor(v, 8); v += shfl_xor(v, 4); v += shfl_xor(v, 2); v += shfl_xor(v, 1);
    return v;
}

template <typename T>
__global__ void __launch_bounds__(kLsqThreads) least_squares_kernel(int B, int R, int P, const T* __restrict__ res,
                                                                    const T* __restrict__ jac,
                                                                    const T* __restrict__ w, T* __restrict__ err,
                                                                    T* __restrict__ grad) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = kLsqThreads / 32;
    for (long long b = (long long)blockIdx.x * warps + warp; b < B; b += (long long)gridDim.x * warps) {
        const T* r = res + (size_t)b * R;
        const T* ww = w ? w + (size_t)b * R : nullptr;
        if (err) {
            T e0 = T(0), e1 = T(0), e2 = T(0), e3 = T(0);
            int i = lane;
            for (; i + 96 < R; i += 128) {                      // four independent loads per trip
                const T r0 = r[i], r1 = r[i + 32], r2 = r[i + 64], r3 = r[i + 96];
                if (ww) {
                    e0 = fma_t(ww[i] * r0, r0, e0); e1 = fma_t(ww[i + 32] * r1, r1, e1);
                    e2 = fma_t(ww[i + 64] * r2, r2, e2); e3 = fma_t(ww[i + 96] * r3, r3, e3);
                } else {
                    e0 = fma_t(r0, r0, e0); e1 = fma_t(r1, r1, e1); e2 = fma_t(r2, r2, e2); e3 = fma_t(r3, r3, e3);
                }
            }
            for (; i < R; i += 32) {
                const T r0 = r[i];
                e0 = ww ? fma_t(ww[i] * r0, r0, e0) : fma_t(r0, r0, e0);   // least_squares_utils.py:24-28
            }
            const T e = lsq_warp_sum((e0 + e1) + (e2 + e3));
            if (lane == 0) err[b] = e;
        }
        if (grad && jac) {
            const T* Jb = jac + (size_t)b * R * P;
            for (int p0 = 0; p0 < P; p0 += kLsqCols) {
                T acc[kLsqCols];
#pragma unroll
                for (int k = 0; k < kLsqCols; ++k) acc[k] = T(0);
                for (int i = lane; i < R; i += 32) {
                    T gr = T(2) * r[i];                         // :43
                    if (ww) gr = ww[i] * gr;                    // :44-45
                    const T* row = Jb + (size_t)i * P + p0;
#pragma unroll
                    for (int k = 0; k < kLsqCols; ++k)
                        if (p0 + k < P) acc[k] = fma_t(gr, row[k], acc[k]);
                }
#pragma unroll
                for (int k = 0; k < kLsqCols; ++k) {
                    const T sum = lsq_warp_sum(acc[k]);
                    if (lane == 0 && p0 + k < P) grad[(size_t)b * P + p0 + k] = sum;
                }
            }
        }
    }
}

template <typename T>
int launch_least_squares(int B, int R, int P, const T* res, const T* jac, const T* w, T* err, T* grad,
                         cudaStream_t s) {
    if (B == 0) return DAVO_OK;
    const long long ctas = ((long long)B + kLsqThreads / 32 - 1) / (kLsqThreads / 32);
    const int grid = (int)(ctas < 148LL * 8 ? ctas : 148LL * 8);
    least_squares_kernel<T><<<grid, kLsqThreads, 0, s>>>(B, R, P, res, jac, w, err, grad);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}
template int launch_least_squares<float>(int, int, int, const float*, const float*, const float*, float*, float*, cudaStream_t);
template int launch_least_squares<double>(int, int, int, const double*, const double*, const double*, double*, double*, cudaStream_t);

// ---- stand-alone BFGS update: one warp per problem, H staged in shared memory ---------------------
template <typename T>
__global__ void __launch_bounds__(128) bfgs_update_kernel(int K, int n, T* __restrict__ H, const T* __restrict__ s_,
                                                          const T* __restrict__ y_) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int ld = n + 1;  // padded row stride: a lane walking a row does not collide with its neighbours
    T* base = reinterpret_cast<T*>(raw) + (size_t)warp * ((size_t)n * ld + 5 * (size_t)n);
    T* Hs = base;
    T* s = Hs + (size_t)n * ld;
    T* y = s + n;
    T* yH = y + n;
    T* Hy = yH + n;
    T* sr = Hy + n;
    for (int b = blockIdx.x * warps + warp; b < K; b += gridDim.x * warps) {
        T* Hg = H + (size_t)b * n * n;
        __syncwarp();
        for (int i = lane; i < n * n; i += 32) Hs[(i / n) * ld + (i % n)] = Hg[i];
        for (int i = lane; i < n; i += 32) {
            s[i] = s_[(size_t)b * n + i];
            y[i] = y_[(size_t)b * n + i];
        }
        __syncwarp();
        T c = T(0);                                           // func_inverse_curvature.py:8-11
        for (int j = 0; j < n; ++j) c = add_rn(c, mul_rn(s[j], y[j]));
        T rho = div_rn(T(1), c);
        if (c <= T(0)) rho = T(0);
        for (int j = lane; j < n; j += 32) {                  // y^T H, bfgs_solver.py:268-270
            T a = T(0);
            for (int i = 0; i < n; ++i) a = add_rn(a, mul_rn(y[i], Hs[i * ld + j]));
            yH[j] = a;
        }
        for (int i = lane; i < n; i += 32) {                  // H y, :293-295
            T a = T(0);
            for (int j = 0; j < n; ++j) a = add_rn(a, mul_rn(Hs[i * ld + j], y[j]));
            Hy[i] = a;
            sr[i] = mul_rn(s[i], rho);                        // :277
        }
        __syncwarp();
        T q = T(0);                                           // :271-274
        for (int j = 0; j < n; ++j) q = add_rn(q, mul_rn(yH[j], mul_rn(y[j], rho)));
        const T onepq = add_rn(T(1), q);
        for (int e = lane; e < n * n; e += 32) {              // :278-303
            const int i = e / n, j = e % n;
            const T sop = mul_rn(mul_rn(sr[i], s[j]), onepq);
            const T sgp = mul_rn(sr[i], yH[j]);
            const T gsp = mul_rn(Hy[i], sr[j]);
            Hg[e] = sub_rn(sub_rn(add_rn(Hs[i * ld + j], sop), sgp), gsp);
        }
    }
}

// n <= 16 (the calibration fits: n = 10): a GROUP of n lanes per problem, 32 / n problems per warp.  Lane i of a
// group keeps row i of H in registers; columns are reached through a padded shared-memory copy.  Every sum runs in
// the same order with the same separately rounded operations as the kernel above and the oracle (bit-identical
// results), but all lanes stream memory: the one-warp-per-problem kernel reached 5 % of the HBM rate at n = 10.
constexpr int kSmallN = 16;

__host__ __device__ inline int bfgs_small_group_words(int n) { return ((n * (n + 1) + 3) & ~3) + 4 * kSmallN; }

// read a 16-entry, 16-byte aligned shared vector into registers with vector loads
template <typename T>
__device__ __forceinline__ void load16(const T* v, int n, T (&out)[kSmallN]) {
    using V4 = typename Vec4<T>::type;
    const V4* v4 = reinterpret_cast<const V4*>(v);
#pragma unroll
    for (int k = 0; k < kSmallN / 4; ++k)
        if (4 * k < n) {
            const V4 t = v4[k];
            out[4 * k] = t.x; out[4 * k + 1] = t.y; out[4 * k + 2] = t.z; out[4 * k + 3] = t.w;
        }
}

template <typename T>
__global__ void __launch_bounds__(128) bfgs_update_small_kernel(int K, int n, T* __restrict__ H,
                                                                const T* __restrict__ s_, const T* __restrict__ y_) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int G = 32 / n;                       // problems per warp
    const int q = lane / n, i = lane - q * n;   // group, row
    const bool member = q < G;
    const int ld = n + 1;
    T* base = reinterpret_cast<T*>(raw) + ((size_t)warp * G + (member ? q : 0)) * bfgs_small_group_words(n);
    T* Hs = base;
    T* ss = Hs + ((n * ld + 3) & ~3);           // four 16-entry vectors, 16-byte aligned
    T* ys = ss + kSmallN;
    T* yHs = ys + kSmallN;
    T* srs = yHs + kSmallN;
    const long long groups = (long long)gridDim.x * warps * G;
    for (long long b0 = ((long long)blockIdx.x * warps + warp) * G; b0 < K; b0 += groups) {
        const long long b = b0 + q;
        const bool on = member && b < K;
        T row[kSmallN];
        T si = T(0);
        T* Hg = H + (size_t)(on ? b : 0) * n * n + (size_t)i * n;
        __syncwarp();
        if (on) {
#pragma unroll
            for (int j = 0; j < kSmallN; ++j)
                if (j < n) {
                    row[j] = Hg[j];
                    Hs[i * ld + j] = row[j];
                }
            si = s_[(size_t)b * n + i];
            ss[i] = si;
            ys[i] = y_[(size_t)b * n + i];
        }
        __syncwarp();
        T rho = T(0), Hy = T(0), sr = T(0);
        T sv[kSmallN], yv[kSmallN];
        if (on) {
            load16(ss, n, sv);
            load16(ys, n, yv);
            T c = T(0);                                       // func_inverse_curvature.py:8-11
#pragma unroll
            for (int j = 0; j < kSmallN; ++j)
                if (j < n) c = add_rn(c, mul_rn(sv[j], yv[j]));
            rho = div_rn(T(1), c);
            if (c <= T(0)) rho = T(0);
            T a = T(0);                                       // (y^T H)_i from column i, bfgs_solver.py:268-270
#pragma unroll
            for (int k = 0; k < kSmallN; ++k)
                if (k < n) a = add_rn(a, mul_rn(yv[k], Hs[k * ld + i]));
            yHs[i] = a;
#pragma unroll
            for (int j = 0; j < kSmallN; ++j)                 // (H y)_i from row i, :293-295
                if (j < n) Hy = add_rn(Hy, mul_rn(row[j], yv[j]));
            sr = mul_rn(si, rho);                             // :277
            srs[i] = sr;
        }
        __syncwarp();
        if (on) {
            T yHv[kSmallN], srv[kSmallN];
            load16(yHs, n, yHv);
            load16(srs, n, srv);
            T qq = T(0);                                      // :271-274
#pragma unroll
            for (int j = 0; j < kSmallN; ++j)
                if (j < n) qq = add_rn(qq, mul_rn(yHv[j], mul_rn(yv[j], rho)));
            const T onepq = add_rn(T(1), qq);
#pragma unroll
            for (int j = 0; j < kSmallN; ++j)                 // :278-303
                if (j < n) {
                    const T sop = mul_rn(mul_rn(sr, sv[j]), onepq);
                    const T sgp = mul_rn(sr, yHv[j]);
                    const T gsp = mul_rn(Hy, srv[j]);
                    Hg[j] = sub_rn(sub_rn(add_rn(row[j], sop), sgp), gsp);
                }
        }
    }
}

template <typename T>
int launch_bfgs_update(int k, int n, T* H, const T* s_, const T* y, cudaStream_t s) {
    if (k == 0) return DAVO_OK;
    if (n < 1 || n > 96) return DAVO_ERR_UNSUPPORTED;
    if (n <= kSmallN) {
        const int G = 32 / n, warps = 4;
        const size_t smem = (size_t)warps * G * bfgs_small_group_words(n) * sizeof(T);
        long long grid = ((long long)k + warps * G - 1) / (warps * G);
        if (grid > 148LL * 16) grid = 148LL * 16;
        bfgs_update_small_kernel<T><<<(unsigned)grid, warps * 32, smem, s>>>(k, n, H, s_, y);
        count_launch();
        return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
    }
    const size_t per_warp = ((size_t)n * (n + 1) + 5 * (size_t)n) * sizeof(T);
    const int warps = (4 * per_warp <= 160 * 1024) ? 4 : 1;
    const size_t smem = warps * per_warp;
    auto kernel = bfgs_update_kernel<T>;
    if (!ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem))
        return DAVO_ERR_CUDA;
    int grid = (k + warps - 1) / warps;
    if (grid > 148 * 8) grid = 148 * 8;
    kernel<<<grid, warps * 32, smem, s>>>(k, n, H, s_, y);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}
template int launch_bfgs_update<float>(int, int, float*, const float*, const float*, cudaStream_t);
template int launch_bfgs_update<double>(int, int, double*, const double*, const double*, cudaStream_t);

template <typename T>
__global__ void bfgs_initial_scale_kernel(int K, int n, const T* __restrict__ s, const T* __restrict__ y,
                                          T* __restrict__ scale) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= K) return;
    T den = T(0), num = T(0);                                  // bfgs_solver.py:229-232
    for (int j = 0; j < n; ++j) den = add_rn(den, mul_rn(y[(size_t)b * n + j], y[(size_t)b * n + j]));
    den = (den < T(1e-5)) ? T(1e-5) : den;
    for (int j = 0; j < n; ++j) num = add_rn(num, mul_rn(s[(size_t)b * n + j], y[(size_t)b * n + j]));
    T sc = div_rn(num, den);
    scale[b] = (sc < T(1e-4)) ? T(1e-4) : sc;
}

template <typename T>
int launch_bfgs_initial_scale(int k, int n, const T* s_, const T* y, T* scale, cudaStream_t s) {
    if (k == 0) return DAVO_OK;
    bfgs_initial_scale_kernel<T><<<(k + 127) / 128, 128, 0, s>>>(k, n, s_, y, scale);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}
template int launch_bfgs_initial_scale<float>(int, int, const float*, const float*, float*, cudaStream_t);
template int launch_bfgs_initial_scale<double>(int, int, const double*, const double*, double*, cudaStream_t);

// ---- interpolate_alpha (utils/func_interpolate_alpha.py) -----------------------------------------------------
// Forward :15-33 (davo_common.cuh interpolate_alpha_value) or the custom backward :42-79, one element per thread.
template <typename T>
__global__ void interpolate_alpha_kernel(long long k, const T* __restrict__ a1, const T* __restrict__ a2,
                                         const T* __restrict__ v1, const T* __restrict__ v2, T* __restrict__ out,
                                         const T* __restrict__ grad_out, T* __restrict__ g_a1, T* __restrict__ g_a2,
                                         T* __restrict__ g_v1, T* __restrict__ g_v2) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < k; i += (long long)gridDim.x * blockDim.x) {
        bool nl;
        const T x1 = a1[i], x2 = a2[i], y1 = v1[i], y2 = v2[i];
        const T c = interpolate_alpha_value(x1, x2, y1, y2, &nl);
        if (out) out[i] = c;
        if (grad_out) {
            const T go = grad_out[i];
            const T diff = y2 - y1;
            const T inv_gradient = (x2 - x1) / diff;            // :19 (saved for backward, :35-37)
            const T one_on_diff = nl ? T(0) : T(1) / diff;      // :34-35
            if (g_a1) g_a1[i] = nl ? T(0.5) * go : one_on_diff * y2 * go;                           // :56-61
            if (g_a2) g_a2[i] = nl ? T(0.5) * go : T(-1) * one_on_diff * y1 * go;                   // :62-67
            if (g_v1) g_v1[i] = nl ? T(0) : T(-1) * y2 * inv_gradient * one_on_diff * go;           // :68-73
            if (g_v2) g_v2[i] = nl ? T(0) : y1 * inv_gradient * one_on_diff * go;                   // :74-79
        }
    }
}

template <typename T>
int launch_interpolate_alpha(long long k, const T* a1, const T* a2, const T* v1, const T* v2, T* out, const T* grad_out,
                             T* g_a1, T* g_a2, T* g_v1, T* g_v2, cudaStream_t s) {
    const int threads = 256;
    long long blocks = (k + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    interpolate_alpha_kernel<T><<<(unsigned)blocks, threads, 0, s>>>(k, a1, a2, v1, v2, out, grad_out, g_a1, g_a2, g_v1, g_v2);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}
template int launch_interpolate_alpha<float>(long long, const float*, const float*, const float*, const float*, float*,
                                             const float*, float*, float*, float*, float*, cudaStream_t);
template int launch_interpolate_alpha<double>(long long, const double*, const double*, const double*, const double*,
                                              double*, const double*, double*, double*, double*, double*, cudaStream_t);

}  // namespace davo
