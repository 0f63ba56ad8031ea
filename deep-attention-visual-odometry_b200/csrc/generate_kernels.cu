// generate_kernels.cu — synthetic oracle-match batches produced directly in HBM.
//
// What it replaces: the reference's host-side generator deep_attention_visual_odometry/data/
// camera_and_parameters_dataset.py:85-151 (world points, look-at extrinsics -> axis-angle, intrinsics, projection +
// visibility mask; that file does not parse at HEAD) with the batch layout of base_types/camera_views_and_points.py:
// 21-33, and this repo's numpy generators for the BASELINE configurations (synthetic.py, distributions of
// SURVEY.md 8(d)).  A 1M-problem batch is 5.4 GB of raw inputs: generated here it never crosses PCIe.
//
// Randomness is counter based (Philox4x32-10, Salmon et al. 2011): every draw is a pure function of
// (seed, GLOBAL problem index, stream, element), so a rank that generates rows [lo, hi) of a sharded batch gets
// exactly the rows a single GPU would have generated, whatever the grid.  oracle/gen_oracle.py restates the same
// generator in numpy; tests compare the two.  All arithmetic is float64 and rounded to the output type at the end
// (the float32 and float64 variants of one seed are the same problems).
#include "davo_common.cuh"
#include "launch.h"

namespace davo {

// ---- Philox4x32-10 ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint4 philox4x32(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
        uint4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k.x;
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k.y;
        n.w = (uint32_t)p0;
        c = n;
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// streams of one problem's draws (counter word 2)
enum : uint32_t { kStreamScalars = 0, kStreamPoints = 1, kStreamNoise = 2, kStreamViews = 3, kStreamStart = 4 };

struct Rng {
    uint2 key;
    uint32_t lo, hi;  // global problem index
    __device__ uint4 draw(uint32_t stream, uint32_t element) const {
        uint4 c;
        c.x = lo; c.y = hi; c.z = stream; c.w = element;
        return philox4x32(c, key);
    }
};

__device__ __forceinline__ double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }  // (0, 1)
// Box-Muller: two 32-bit words -> two independent N(0,1)
__device__ __forceinline__ void normal2(uint32_t a, uint32_t b, double& n0, double& n1) {
    const double r = sqrt(-2.0 * log(u01(a)));
    double s, c;
    sincospi(2.0 * u01(b), &s, &c);
    n0 = r * c;
    n1 = r * s;
}
__device__ __forceinline__ void normal4(const uint4 w, double (&n)[4]) {
    normal2(w.x, w.y, n[0], n[1]);
    normal2(w.z, w.w, n[2], n[3]);
}

struct GenParams {
    int B, N, V;
    uint64_t seed, first;  // first = global index of row 0 (sharded generation)
    double fov, noise, pathological, start_noise;
    int ill_conditioned, random_pose;
};

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// camera_model/distorted_camera_model.py:24-103 in float64 (observations of the generated truth)
__device__ __forceinline__ void forward16(const double* th, const double* R, double X, double Y, double Z, double& up,
                                          double& vp) {
    const double xp = R[0] * X + R[1] * Y + R[2] * Z + th[DAVO_TX];
    const double yp = R[3] * X + R[4] * Y + R[5] * Z + th[DAVO_TY];
    double zp = R[6] * X + R[7] * Y + R[8] * Z + th[DAVO_TZ];
    if (zp == 0.0) zp += 1e-8;
    const double a = xp / zp, b = yp / zp;
    const double u = th[DAVO_FX] * a + th[DAVO_S] * b, v = th[DAVO_FY] * b;
    const double r2 = u * u + v * v;
    const double rad = 1.0 + th[DAVO_K1] * r2 + th[DAVO_K2] * r2 * r2 + th[DAVO_K3] * r2 * r2 * r2;
    up = u * rad + 2.0 * th[DAVO_P1] * u * v + th[DAVO_P2] * (r2 + 2.0 * u * u) + th[DAVO_CX];
    vp = v * rad + 2.0 * th[DAVO_P2] * u * v + th[DAVO_P1] * (r2 + 2.0 * v * v) + th[DAVO_CY];
}

__device__ __forceinline__ void euler_rows(double rx, double ry, double rz, double* R) {  // :38-55, R = Rz Ry Rx
    double sx, cx, sy, cy, sz, cz;
    sincos(rx, &sx, &cx);
    sincos(ry, &sy, &cy);
    sincos(rz, &sz, &cz);
    R[0] = cy * cz; R[1] = sx * sy * cz - cx * sz; R[2] = cx * sy * cz + sx * sz;
    R[3] = cy * sz; R[4] = sx * sy * sz + cx * cz; R[5] = cx * sy * sz - sx * cz;
    R[6] = -sy;     R[7] = sx * cy;                R[8] = cx * cy;
}

// intrinsics + distortion of SURVEY.md 8(d) (synthetic._intrinsics) and the start of configs 2 / 4
__device__ __forceinline__ void draw_intrinsics(const Rng& rng, const GenParams& g, double* truth10, double* x0_10,
                                                double& u_path) {
    const uint4 s0 = rng.draw(kStreamScalars, 0);
    double n1[4], n2[4];
    normal4(rng.draw(kStreamScalars, 1), n1);
    normal4(rng.draw(kStreamScalars, 2), n2);
    const double k1s = g.ill_conditioned ? 0.5 : 0.05, k2s = g.ill_conditioned ? 0.2 : 0.005,
                 k3s = g.ill_conditioned ? 0.1 : 0.0005, ps = g.ill_conditioned ? 0.05 : 0.005;
    const double fx = 1.0 + 0.5 * u01(s0.x);
    truth10[DAVO_FX] = fx;
    truth10[DAVO_FY] = fx * (1.0 + 0.02 * n1[0]);
    truth10[DAVO_S] = 0.0;
    truth10[DAVO_CX] = clampd(0.1 * n1[1], -0.5, 0.5);
    truth10[DAVO_CY] = clampd(0.1 * n1[2], -0.5, 0.5);
    truth10[DAVO_K1] = k1s * n1[3];
    truth10[DAVO_K2] = k2s * n2[0];
    truth10[DAVO_K3] = k3s * n2[1];
    truth10[DAVO_P1] = ps * n2[2];
    truth10[DAVO_P2] = ps * n2[3];
    const double uf = u01(s0.y);
    const double f0 = g.ill_conditioned ? fx * exp(log(0.3) + uf * (log(3.0) - log(0.3)))
                                        : fx * (1.0 + 0.2 * (2.0 * uf - 1.0));
#pragma unroll
    for (int j = 0; j < 10; ++j) x0_10[j] = 0.0;
    x0_10[DAVO_FX] = f0;
    x0_10[DAVO_FY] = f0;
    u_path = u01(s0.z);
}

// camera-frame points of SURVEY.md 8(d) (synthetic._points): z = |4 + N(0,1)| + 1, xy = z * fov * U(-1,1)
__device__ __forceinline__ void draw_point(const Rng& rng, const GenParams& g, int i, double& X, double& Y, double& Z) {
    const uint4 w = rng.draw(kStreamPoints, (uint32_t)i);
    double n0, n1;
    normal2(w.x, w.y, n0, n1);
    Z = fabs(4.0 + n0) + 1.0;
    X = Z * g.fov * (2.0 * u01(w.z) - 1.0);
    Y = Z * g.fov * (2.0 * u01(w.w) - 1.0);
}

// ---- DISTORT10 batches (BASELINE configs 2, 4, 5): one warp per problem, lanes over the matches -----------------
template <typename T>
__global__ void __launch_bounds__(256) generate_distort10_kernel(GenParams g, T* __restrict__ pts, T* __restrict__ obs,
                                                                 T* __restrict__ pose, T* __restrict__ x0,
                                                                 T* __restrict__ truth) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < g.B; b += warps) {
        const uint64_t gb = g.first + (uint64_t)b;
        Rng rng{make_uint2((uint32_t)g.seed, (uint32_t)(g.seed >> 32)), (uint32_t)gb, (uint32_t)(gb >> 32)};
        double th[16], xs[10], u_path;
        draw_intrinsics(rng, g, th, xs, u_path);
#pragma unroll
        for (int j = 10; j < 16; ++j) th[j] = 0.0;
        if (g.random_pose) {
            double a[4], c[4];
            normal4(rng.draw(kStreamScalars, 3), a);
            normal4(rng.draw(kStreamScalars, 4), c);
            th[DAVO_RX] = 0.2 * a[0]; th[DAVO_RY] = 0.2 * a[1]; th[DAVO_RZ] = 0.2 * a[2];
            th[DAVO_TX] = 0.3 * a[3]; th[DAVO_TY] = 0.3 * c[0]; th[DAVO_TZ] = 0.3 * c[1];
        }
        // config 4's pathological rows: points a hair in front of the camera plane (first half of the fraction) or an
        // ascent-inducing start on the far side of a pole of the radial polynomial (second half)
        const bool near = u_path < 0.5 * g.pathological;
        const bool ascent = !near && u_path < g.pathological;
        if (ascent) {
            xs[DAVO_K1] = 5.0;
            xs[DAVO_FX] *= -1.0;
        }
        double R[9];
        euler_rows(th[DAVO_RX], th[DAVO_RY], th[DAVO_RZ], R);
        for (int i = lane; i < g.N; i += 32) {
            double X, Y, Z, up, vp;
            draw_point(rng, g, i, X, Y, Z);
            if (near && i < 4) Z = 1e-6;
            forward16(th, R, X, Y, Z, up, vp);
            if (g.noise > 0.0) {
                const uint4 w = rng.draw(kStreamNoise, (uint32_t)i);
                double n0, n1;
                normal2(w.x, w.y, n0, n1);
                up += g.noise * n0;
                vp += g.noise * n1;
            }
            const size_t m = (size_t)b * g.N + i;
            pts[3 * m + 0] = (T)X; pts[3 * m + 1] = (T)Y; pts[3 * m + 2] = (T)Z;
            obs[2 * m + 0] = (T)up; obs[2 * m + 1] = (T)vp;
        }
        if (lane < 10) {
            x0[(size_t)b * 10 + lane] = (T)xs[lane];
            if (truth) truth[(size_t)b * 10 + lane] = (T)th[lane];
        }
        if (pose && lane < 6) pose[(size_t)b * 6 + lane] = (T)th[10 + lane];
    }
}

// ---- JOINT batches (BASELINE config 3): V views of N shared world points, a 6-DoF pose per view ------------------
template <typename T>
__global__ void __launch_bounds__(256) generate_joint_kernel(GenParams g, T* __restrict__ pts, T* __restrict__ obs,
                                                             T* __restrict__ x0, T* __restrict__ truth) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int n = 10 + 6 * g.V;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < g.B; b += warps) {
        const uint64_t gb = g.first + (uint64_t)b;
        Rng rng{make_uint2((uint32_t)g.seed, (uint32_t)(g.seed >> 32)), (uint32_t)gb, (uint32_t)(gb >> 32)};
        double th[16], xs[10], u_path;
        draw_intrinsics(rng, g, th, xs, u_path);
        if (lane < 10) {
            x0[(size_t)b * n + lane] = (T)xs[lane];
            if (truth) truth[(size_t)b * n + lane] = (T)th[lane];
        }
        for (int v = 0; v < g.V; ++v) {
            // truth pose r ~ 0.2 N, t ~ 0.3 N; start = truth + 0.05 N (rotation) / 0.1 N (translation)
            double a[4], c[4], sa[4], sc[4];
            normal4(rng.draw(kStreamViews, 2 * v), a);
            normal4(rng.draw(kStreamViews, 2 * v + 1), c);
            normal4(rng.draw(kStreamStart, 2 * v), sa);
            normal4(rng.draw(kStreamStart, 2 * v + 1), sc);
            th[DAVO_RX] = 0.2 * a[0]; th[DAVO_RY] = 0.2 * a[1]; th[DAVO_RZ] = 0.2 * a[2];
            th[DAVO_TX] = 0.3 * a[3]; th[DAVO_TY] = 0.3 * c[0]; th[DAVO_TZ] = 0.3 * c[1];
            if (lane < 6) {
                const double st = lane < 3 ? 0.05 : 0.1;
                const double sn = lane == 0 ? sa[0] : lane == 1 ? sa[1] : lane == 2 ? sa[2] : lane == 3 ? sa[3]
                                  : lane == 4 ? sc[0] : sc[1];
                const size_t o = (size_t)b * n + 10 + 6 * v + lane;
                x0[o] = (T)(th[10 + lane] + st * sn);
                if (truth) truth[o] = (T)th[10 + lane];
            }
            double R[9];
            euler_rows(th[DAVO_RX], th[DAVO_RY], th[DAVO_RZ], R);
            for (int i = lane; i < g.N; i += 32) {
                double X, Y, Z, up, vp;
                draw_point(rng, g, i, X, Y, Z);
                forward16(th, R, X, Y, Z, up, vp);
                if (g.noise > 0.0) {
                    const uint4 w = rng.draw(kStreamNoise, (uint32_t)(v * g.N + i));
                    double n0, n1;
                    normal2(w.x, w.y, n0, n1);
                    up += g.noise * n0;
                    vp += g.noise * n1;
                }
                if (v == 0) {
                    const size_t m = (size_t)b * g.N + i;
                    pts[3 * m + 0] = (T)X; pts[3 * m + 1] = (T)Y; pts[3 * m + 2] = (T)Z;
                }
                const size_t m = ((size_t)b * g.V + v) * g.N + i;
                obs[2 * m + 0] = (T)up; obs[2 * m + 1] = (T)vp;
            }
        }
    }
}

// ---- CameraViewsAndPoints batches (the entry script's data: M views of N world points) ---------------------------
// data/camera_and_parameters_dataset.py:85-151.  World points relative to the first view: xy ~ 3 N(0,1),
// z = |20 + 5 N(0,1)| (:85-94).  Each further camera: location ~ 3 N(0,1), looking at a target near the points'
// centre of mass (centroid * (1 + U) + 1.5 N + 3 N, :107-116) with "up" near up_distance * (0,-1,0) + 3 N, the two
// directions orthonormalised (:119-124).  The reference stacks (-left, -up, forward) with left = forward x up, an
// improper matrix (determinant -1) that mat2axangle cannot represent; here the camera axes are x = y x z, y = -up,
// z = forward (a rotation, identity when the camera looks down +z with -y up), converted to axis-angle, and the
// camera is pushed back along its view direction until every point is min_camera_distance in front of it (:134-143).
// camera-relative point = R (X - location) = R X + t with t = -R location (the convention of
// get_camera_relative_points, camera_model/calibration_pinhole_camera_model.py:78-117).  Intrinsics (:147-151):
// f' = 1 / tan(U(30 deg, 120 deg) / 2), centre ~ clamp(0.2 N, +-0.5); u = f' x / max(z, 1e-8) + cx, visibility =
// |u| < 1 and |v| < 1 (:185-192).  One warp per problem; lane v-1 builds camera v.
constexpr int kMaxViews = 32;

template <typename T>
__global__ void __launch_bounds__(128) generate_views_kernel(GenParams g, double min_camera_distance,
                                                             T* __restrict__ projected, T* __restrict__ visibility,
                                                             T* __restrict__ intrinsics, T* __restrict__ orientations,
                                                             T* __restrict__ translations, T* __restrict__ world,
                                                             T* __restrict__ x0, T* __restrict__ truth) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = g.V, N = g.N, n = 3 + 3 * N + 6 * (M - 1);
    double* Xs = sm + (size_t)warp * (3 * N + 12 * kMaxViews);  // world points, then per-view R (9) + t (3)
    double* cam = Xs + 3 * N;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < g.B; b += warps) {
        const uint64_t gb = g.first + (uint64_t)b;
        Rng rng{make_uint2((uint32_t)g.seed, (uint32_t)(g.seed >> 32)), (uint32_t)gb, (uint32_t)(gb >> 32)};
        __syncwarp();
        double cxs = 0.0, cys = 0.0, czs = 0.0;
        for (int i = lane; i < N; i += 32) {
            double nn[4];
            normal4(rng.draw(kStreamPoints, (uint32_t)i), nn);
            const double X = 3.0 * nn[0], Y = 3.0 * nn[1], Z = fabs(20.0 + 5.0 * nn[2]);
            Xs[3 * i] = X; Xs[3 * i + 1] = Y; Xs[3 * i + 2] = Z;
            cxs += X; cys += Y; czs += Z;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cxs += __shfl_xor_sync(kFull, cxs, o);
            cys += __shfl_xor_sync(kFull, cys, o);
            czs += __shfl_xor_sync(kFull, czs, o);
        }
        // scalars shared by the cameras of this problem
        const uint4 s0 = rng.draw(kStreamScalars, 0);
        double ns[4], nt[4];
        normal4(rng.draw(kStreamScalars, 1), ns);
        normal4(rng.draw(kStreamScalars, 2), nt);
        const double fov = 3.0 * M_PI / 18.0 + (9.0 * M_PI / 18.0) * u01(s0.x);
        const double fp = 1.0 / tan(0.5 * fov);
        const double pcx = clampd(0.2 * ns[0], -0.5, 0.5), pcy = clampd(0.2 * ns[1], -0.5, 0.5);
        const double up_distance = fabs(20.0 + 5.0 * ns[2]);
        const double tscale = 1.0 + u01(s0.y);
        const double tbx = cxs / N * tscale + 1.5 * nt[0], tby = cys / N * tscale + 1.5 * nt[1],
                     tbz = czs / N * tscale + 1.5 * nt[2];
        __syncwarp();
        if (lane >= 1 && lane < M) {
            const int v = lane;
            double a[4], c[4], e[4];
            normal4(rng.draw(kStreamViews, 3 * v), a);
            normal4(rng.draw(kStreamViews, 3 * v + 1), c);
            normal4(rng.draw(kStreamViews, 3 * v + 2), e);
            double L[3] = {3.0 * a[0], 3.0 * a[1], 3.0 * a[2]};
            const double tg[3] = {tbx + 3.0 * a[3], tby + 3.0 * c[0], tbz + 3.0 * c[1]};
            const double ub[3] = {3.0 * c[2], -up_distance + 3.0 * c[3], 3.0 * e[0]};
            double f[3] = {tg[0] - L[0], tg[1] - L[1], tg[2] - L[2]};
            double u[3] = {ub[0] - L[0], ub[1] - L[1], ub[2] - L[2]};
            const double fn = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
            f[0] /= fn; f[1] /= fn; f[2] /= fn;
            const double fu = f[0] * u[0] + f[1] * u[1] + f[2] * u[2];
            u[0] -= f[0] * fu; u[1] -= f[1] * fu; u[2] -= f[2] * fu;
            const double un = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
            const double y[3] = {-u[0] / un, -u[1] / un, -u[2] / un};                 // image y points down
            const double x[3] = {y[1] * f[2] - y[2] * f[1], y[2] * f[0] - y[0] * f[2], y[0] * f[1] - y[1] * f[0]};
            // push the camera back until every point is at least min_camera_distance in front of it
            double zmin = 1e300;
            for (int i = 0; i < N; ++i) {
                const double z = f[0] * (Xs[3 * i] - L[0]) + f[1] * (Xs[3 * i + 1] - L[1]) + f[2] * (Xs[3 * i + 2] - L[2]);
                zmin = fmin(zmin, z - min_camera_distance);
            }
            if (zmin < 1e-3) { L[0] += zmin * f[0]; L[1] += zmin * f[1]; L[2] += zmin * f[2]; }  // zmin < 0: backwards
            double* R = cam + 12 * v;
            R[0] = x[0]; R[1] = x[1]; R[2] = x[2];
            R[3] = y[0]; R[4] = y[1]; R[5] = y[2];
            R[6] = f[0]; R[7] = f[1]; R[8] = f[2];
            R[9] = -(x[0] * L[0] + x[1] * L[1] + x[2] * L[2]);
            R[10] = -(y[0] * L[0] + y[1] * L[1] + y[2] * L[2]);
            R[11] = -(f[0] * L[0] + f[1] * L[1] + f[2] * L[2]);
            // rotation matrix -> axis-angle (the reference calls transforms3d.mat2axangle)
            const double wx = R[7] - R[5], wy = R[2] - R[6], wz = R[3] - R[1];        // 2 sin(angle) * axis
            const double s2 = sqrt(wx * wx + wy * wy + wz * wz);
            const double angle = atan2(0.5 * s2, 0.5 * (R[0] + R[4] + R[8] - 1.0));
            const double k = s2 > 1e-12 ? angle / s2 : 0.5;
            const double om[3] = {k * wx, k * wy, k * wz};
            const size_t o = ((size_t)b * (M - 1) + (v - 1)) * 3;
            double sn[4], sr[4];
            normal4(rng.draw(kStreamStart, 2 * v), sn);
            normal4(rng.draw(kStreamStart, 2 * v + 1), sr);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                orientations[o + j] = (T)om[j];
                translations[o + j] = (T)R[9 + j];
                const size_t pt = (size_t)b * n + 3 + 3 * N + 3 * (v - 1) + j;
                const size_t pr = pt + 3 * (M - 1);
                if (truth) { truth[pt] = (T)R[9 + j]; truth[pr] = (T)om[j]; }
                if (x0) { x0[pt] = (T)(R[9 + j] + g.start_noise * 0.3 * sn[j]); x0[pr] = (T)(om[j] + g.start_noise * 0.03 * sr[j]); }
            }
        }
        if (lane == 0) {  // view 0 is the world frame
            double* R = cam;
            R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
            R[9] = R[10] = R[11] = 0;
            // f' = elu(f) + 1 (geometry/homogeneous_projection.py:37): the parameter the objective optimises
            const double fpar = fp > 1.0 ? fp - 1.0 : log(fp);
            intrinsics[(size_t)b * 3] = (T)fp; intrinsics[(size_t)b * 3 + 1] = (T)pcx; intrinsics[(size_t)b * 3 + 2] = (T)pcy;
            double sn[4];
            normal4(rng.draw(kStreamStart, 0), sn);
            const double tv[3] = {fpar, pcx, pcy}, sg[3] = {0.1, 0.05, 0.05};
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (truth) truth[(size_t)b * n + j] = (T)tv[j];
                if (x0) x0[(size_t)b * n + j] = (T)(tv[j] + g.start_noise * sg[j] * sn[j]);
            }
        }
        __syncwarp();
        for (int i = lane; i < N; i += 32) {
            const double X = Xs[3 * i], Y = Xs[3 * i + 1], Z = Xs[3 * i + 2];
            double sn[4];
            normal4(rng.draw(kStreamStart, 64 + i), sn);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double c = Xs[3 * i + j];
                world[((size_t)b * N + i) * 3 + j] = (T)c;
                if (truth) truth[(size_t)b * n + 3 + 3 * i + j] = (T)c;
                if (x0) x0[(size_t)b * n + 3 + 3 * i + j] = (T)(c + g.start_noise * 0.3 * sn[j]);
            }
            for (int v = 0; v < M; ++v) {
                const double* R = cam + 12 * v;
                const double xr = R[0] * X + R[1] * Y + R[2] * Z + R[9], yr = R[3] * X + R[4] * Y + R[5] * Z + R[10],
                             zr = R[6] * X + R[7] * Y + R[8] * Z + R[11];
                const double zc = fmax(zr, 1e-8);
                const double pu = fp * xr / zc + pcx, pv = fp * yr / zc + pcy;
                const size_t m = ((size_t)b * M + v) * N + i;
                projected[2 * m] = (T)pu; projected[2 * m + 1] = (T)pv;
                visibility[m] = (pu > -1.0 && pu < 1.0 && pv > -1.0 && pv < 1.0 && zr > 0.0) ? T(1) : T(0);
            }
        }
    }
}

static GenParams to_params(const davo_generator_desc* d) {
    GenParams g{};
    g.B = d->B; g.N = d->N; g.V = d->V;
    g.seed = d->seed; g.first = d->first_problem;
    g.fov = d->fov; g.noise = d->noise; g.pathological = d->pathological; g.start_noise = d->start_noise;
    g.ill_conditioned = d->ill_conditioned; g.random_pose = d->random_pose;
    return g;
}

static int gen_grid(int B, int warps_per_cta) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long need = ((long long)B + warps_per_cta - 1) / warps_per_cta;
    long long cap = (long long)sms * 8;
    return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

template <typename T>
int launch_generate_distort10(const davo_generator_desc* d, T* pts, T* obs, T* pose, T* x0, T* truth, cudaStream_t s) {
    generate_distort10_kernel<T><<<gen_grid(d->B, 8), 256, 0, s>>>(to_params(d), pts, obs, pose, x0, truth);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}
template <typename T>
int launch_generate_joint(const davo_generator_desc* d, T* pts, T* obs, T* x0, T* truth, cudaStream_t s) {
    generate_joint_kernel<T><<<gen_grid(d->B, 8), 256, 0, s>>>(to_params(d), pts, obs, x0, truth);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}
template <typename T>
int launch_generate_views(const davo_generator_desc* d, T* projected, T* visibility, T* intrinsics, T* orientations,
                          T* translations, T* world, T* x0, T* truth, cudaStream_t s) {
    if (d->V > kMaxViews) return DAVO_ERR_UNSUPPORTED;
    const size_t smem = 4 * (3 * (size_t)d->N + 12 * kMaxViews) * sizeof(double);
    auto kernel = generate_views_kernel<T>;
    if (!ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem)) return DAVO_ERR_UNSUPPORTED;
    kernel<<<gen_grid(d->B, 4), 128, smem, s>>>(to_params(d), d->min_camera_distance, projected, visibility, intrinsics,
                                               orientations, translations, world, x0, truth);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}

#define DAVO_INSTANTIATE_GEN(T)                                                                                      \
    template int launch_generate_distort10<T>(const davo_generator_desc*, T*, T*, T*, T*, T*, cudaStream_t);         \
    template int launch_generate_joint<T>(const davo_generator_desc*, T*, T*, T*, T*, cudaStream_t);                 \
    template int launch_generate_views<T>(const davo_generator_desc*, T*, T*, T*, T*, T*, T*, T*, T*, cudaStream_t);
DAVO_INSTANTIATE_GEN(float)
DAVO_INSTANTIATE_GEN(double)

}  // namespace davo
