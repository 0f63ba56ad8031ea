// objectives_ba.cuh — the entry script's objective: bundle adjustment with an angular error
// (networks/calibration_network.py:58-67).  Parameters of one problem, n = 3 + 3N + 6(V-1):
//   (f, cx, cy | X[N][3] world points | t[V-1][3] translations | w[V-1][3] axis-angle rotations)
// as camera_model/calibration_pinhole_camera_model.py:33-75 unpacks them; view 0 is the origin.
//   sigma  = (N mean|X| + V mean|t|) / (N + V),  Xs = X / sigma,  ts = t / sigma             (:98-104)
//   P_0j   = Xs_j,   P_mj = R(w_m) Xs_j + ts_m  (Rodrigues, geometry/axis_angle_rotation.py:25-54, with the
//            Taylor-guarded sin(a)/a and (1-cos a)/a^2 of utils/func_sin_x_on_x.py, func_one_minus_cos_x_on_x_squared.py)
//   h_mj   = (u_mj - cx, v_mj - cy, elu(f) + 1)            (geometry/homogeneous_projection.py:21-44)
//   error  = sum_mj vis_mj * 2 atan2(|h^ - P^|, |h^ + P^|)  (geometry/projective_plane_angle_distance.py:20-64)
// The gradient is reverse mode written out by hand (oracle/calib_oracle_impl.h angle_ba is the scalar
// statement; it matches torch.autograd of the reference functions to 1e-15 in float64, zero sub-gradients
// included).
//
// One warp evaluates one problem: lane i handles (view, point) pairs i, i+32, ...; every pair writes its 9
// partial derivatives (d/dXs, d/dts, d/dw) as one row of a shared-memory table, and the lane that owns
// parameter c then adds up the rows that touch c in a fixed order (deterministic, no atomics).
#pragma once
#include "davo_common.cuh"
#include "solver_wide.cuh"

namespace davo {

// norm = max(sqrt(sq), floor), inv = 1 / norm.  (inv is unused garbage when norm == 0 and floor == 0: callers test.)
__device__ __forceinline__ void norm_and_inverse(double sq, double floor, double& norm, double& inv) {
    norm = sqrt_rn(sq);
    norm = norm < floor ? floor : norm;
    inv = rcp_rn(norm);
}
#ifndef DAVO_BA_EXACT_NORMS
#define DAVO_BA_EXACT_NORMS 0  // 1: A/B builds round the float32 norms and reciprocals correctly, like float64
#endif
__device__ __forceinline__ void norm_and_inverse(float sq, float floor, float& norm, float& inv) {
#if DAVO_BA_EXACT_NORMS
    norm = sqrt_rn(sq);
    norm = norm < floor ? floor : norm;
    inv = rcp_rn(norm);
    return;
#endif
    const float f2 = floor * floor;
    const float v = sq < f2 ? f2 : sq;
    inv = rsqrtf(v);
    norm = v * inv;
    if (v == 0.0f) { norm = 0.0f; inv = 0.0f; }
}

// kV, kN > 0 fix the number of views / points at compile time (the entry script's configuration is 4 views x 8
// points, camera_calibration_from_oracle_matches.py:34-35: loops unroll, index arithmetic folds); 0 = run time.
// kDataGrad (the backward pass of the differentiable solve only, solver_train.cuh): while `dgrad` is set every
// evaluation also adds dcoef * d error / d obs to dgrad[V,N,2] — pair i is always visited by the same lane.
template <typename T, int kV = 0, int kN = 0, bool kDataGrad = false>
struct AngleBAObjective {
    T* dgrad = nullptr;
    T dcoef = T(0);
    static constexpr int kRow = 9;  // odd pitch: consecutive rows start in different banks
    static constexpr int kParams = (kV > 0 && kN > 0) ? 3 + 3 * kN + 6 * (kV - 1) : 0;  // compile-time n (0: run time)
    const SolveParams<T>& p;
    T* obs;      // [V*N][2]
    T* vis;      // [V*N]
    T* contrib;  // [V*N][kRow]
    T* rot;      // [V-1][8]: cos a, sin a, sin a / a, (1 - cos a)/a^2, d(sin a / a)/da, d((1 - cos a)/a^2)/da, 1/a | 0
    int lane;
    int m0, j0;  // (view, point) of this lane's first pair
    T wX, wT;    // d sigma / d|X_jk| = 1 / (3 (N + V)),  d sigma / d|t_mk| = V / (3 (V - 1) (N + V))

    __host__ __device__ static size_t slab_bytes(int N, int V, bool) {
        size_t b = sizeof(T) * ((size_t)V * N * (3 + kRow) + 8 * (size_t)(V > 1 ? V - 1 : 1));
        return (b + 127) & ~size_t(127);
    }

    __device__ AngleBAObjective(const SolveParams<T>& p_, unsigned char* slab, int lane_) : p(p_), lane(lane_) {
        obs = reinterpret_cast<T*>(slab);
        vis = obs + (size_t)p.V * p.N * 2;
        contrib = vis + (size_t)p.V * p.N;
        rot = contrib + (size_t)p.V * p.N * kRow;
        m0 = lane / p.N;
        j0 = lane - m0 * p.N;
        // sigma = (N mean|X| + V mean|t|) / (N + V) with mean|X| over 3N entries, mean|t| over 3(V-1) entries
        wX = T(1) / T(3 * (p.N + p.V));
        wT = T(p.V) / T(3 * (p.V - 1) * (p.N + p.V));
    }

    __device__ __forceinline__ void init() {}

    __device__ __forceinline__ void bind(int b) {
        __syncwarp();
        const int MN = p.V * p.N;
        const T* go = p.data0 + (size_t)b * MN * 2;
        for (int i = lane; i < 2 * MN; i += 32) obs[i] = go[i];
        if (p.has_w) {
            const T* gv = p.w + (size_t)b * MN;
            for (int i = lane; i < MN; i += 32) vis[i] = gv[i];
        } else {
            for (int i = lane; i < MN; i += 32) vis[i] = T(1);
        }
        __syncwarp();
    }

    // th: parameters in shared memory; gout: gradient written to shared memory; returns the error.
    __device__ __forceinline__ T eval(const T* th, T* gout) {
        const int N = kN > 0 ? kN : p.N, V = kV > 0 ? kV : p.V, MN = V * N, n = kParams > 0 ? kParams : p.n;
        const T* X = th + 3;
        const T* t = X + 3 * N;
        const T* w = t + 3 * (V - 1);
        const T kEps = T(2.220446049250313e-16);  // projective_plane_angle_distance.py:48,51
        const int tb = 3 + 3 * N, wb = tb + 3 * (V - 1);   // first translation / rotation parameter
        // one scale for points and translations, calibration_pinhole_camera_model.py:98-104 (X and t are
        // contiguous in the parameter vector: one weighted sum of absolute values)
        T part = T(0);
        for (int i = 3 + lane; i < wb; i += 32) part = fma_t(fabs(th[i]), i < tb ? wX : wT, part);
        const T sig = warp_allreduce(part);
        const T isig = rcp_rn(sig);  // the divisions by sigma, |h|, |P| below are multiplications by one reciprocal each
        __syncwarp();  // the previous evaluation's readers of rot / contrib are done
        if (lane < V - 1) {
            const T w0 = w[3 * lane], w1 = w[3 * lane + 1], w2 = w[3 * lane + 2];
            const T a = sqrt_rn(w0 * w0 + w1 * w1 + w2 * w2);  // axis_angle_rotation.py:37
            const T a2 = a * a;
            T sn, cn;
            sincos(a, &sn, &cn);
            T s, k, oc;
            const T rec = (a == T(0)) ? T(0) : rcp_rn(a);   // 1/a once; the quotients below multiply by it
            const T rec2 = rec * rec;
            if (fabs(a) < T(0.01)) {  // func_sin_x_on_x.py:10-22, :45-66
                const T a4 = a2 * a2, a6 = a4 * a2;
                s = T(1) - a2 * T(1.0 / 6.0) + a4 * T(1.0 / 120.0) - a6 * T(1.0 / 5040.0);
                k = T(-1.0 / 3.0) + a2 * T(1.0 / 30.0) - a4 * T(1.0 / 840.0) + a6 * T(1.0 / 45360.0);
            } else {
                s = sn * rec;
                k = (cn - sn * rec) * rec2;              // cos a / a^2 - sin a / a^3
            }
            if (fabs(a) < T(0.05)) {  // func_one_minus_cos_x_on_x_squared.py:12-28
                const T a4 = a2 * a2, a6 = a4 * a2;
                oc = T(0.5) - a2 * T(1.0 / 24.0) + a4 * T(1.0 / 720.0) - a6 * T(1.0 / 40320.0);
            } else {
                oc = (T(1) - cn) * rec2;
            }
            T* r = rot + 8 * lane;
            r[0] = cn; r[1] = sn; r[2] = s; r[3] = oc;
            r[4] = a * k;                   // SinXonX.backward
            r[5] = rec * (s - T(2) * oc);   // OneMinusCosXonXsquared.backward
            r[6] = rec;
        }
        __syncwarp();
        const T f = th[0], cx = th[1], cy = th[2];
        const T ef = exp(f);
        const T fp = f > T(0) ? f + T(1) : ef;   // elu(f) + 1, homogeneous_projection.py:37
        const T dfp = f > T(0) ? T(1) : ef;
        T cost = T(0), gf = T(0), gcx = T(0), gcy = T(0);
        for (int i = lane; i < MN; i += 32) {
            int m = m0, j = j0;
            if (i != lane) {
                m = i / N;
                j = i - m * N;
            }
            const T x0 = X[3 * j] * isig, x1 = X[3 * j + 1] * isig, x2 = X[3 * j + 2] * isig;
            T P0 = x0, P1 = x1, P2 = x2;
            T o0 = T(0), o1 = T(0), o2 = T(0), c0 = T(0), c1 = T(0), c2 = T(0), dot = T(0);
            T cn = T(1), sn = T(0), s = T(1), oc = T(0.5), ds = T(0), doc = T(0), rec = T(0);
            T ts0 = T(0), ts1 = T(0), ts2 = T(0);
            if (m > 0) {  // axis_angle_rotation.py:38-48, then + translation (calibration_pinhole_camera_model.py:110)
                const T* r = rot + 8 * (m - 1);
                cn = r[0]; sn = r[1]; s = r[2]; oc = r[3]; ds = r[4]; doc = r[5]; rec = r[6];
                o0 = w[3 * (m - 1)]; o1 = w[3 * (m - 1) + 1]; o2 = w[3 * (m - 1) + 2];
                ts0 = t[3 * (m - 1)] * isig; ts1 = t[3 * (m - 1) + 1] * isig; ts2 = t[3 * (m - 1) + 2] * isig;
                dot = x0 * o0 + x1 * o1 + x2 * o2;
                c0 = o1 * x2 - o2 * x1; c1 = o2 * x0 - o0 * x2; c2 = o0 * x1 - o1 * x0;
                const T od = oc * dot;
                P0 = x0 * cn + od * o0 + c0 * s + ts0;
                P1 = x1 * cn + od * o1 + c1 * s + ts1;
                P2 = x2 * cn + od * o2 + c2 * s + ts2;
            }
            const T h0 = obs[2 * i] - cx, h1 = obs[2 * i + 1] - cy, h2 = fp;  // homogeneous_projection.py:38-44
            // |h|, |P|, |a + b|, |a - b| and their reciprocals: float64 rounds every one correctly (the parity gates of
            // this objective are float64); float32 takes norm and reciprocal from ONE reciprocal square root each
            // (MUFU.RSQ, ~1 ulp): the correctly rounded sqrt and 1/x sequences were a fifth of the evaluator
            T nh, nP, inh, inP;
            norm_and_inverse(h0 * h0 + h1 * h1 + h2 * h2, kEps, nh, inh);
            norm_and_inverse(P0 * P0 + P1 * P1 + P2 * P2, kEps, nP, inP);
            const T a0 = h0 * inh, a1 = h1 * inh, a2 = h2 * inh;
            const T b0 = P0 * inP, b1 = P1 * inP, b2 = P2 * inP;
            const T s0 = a0 + b0, s1 = a1 + b1, s2 = a2 + b2;
            const T d0 = a0 - b0, d1 = a1 - b1, d2 = a2 - b2;
            T S, D, rS, rD;
            norm_and_inverse(s0 * s0 + s1 * s1 + s2 * s2, T(0), S, rS);
            norm_and_inverse(d0 * d0 + d1 * d1 + d2 * d2, T(0), D, rD);
            const T vz = vis[i];
            cost += T(2) * atan2(D, S) * vz;  // projective_plane_angle_distance.py:53-60
            const T w2 = T(2) * vz * rcp_rn(S * S + D * D);  // atan2 backward: d/dD = S / (S^2 + D^2), d/dS = -D / (S^2 + D^2)
            const T iD = (D != T(0)) ? w2 * S * rD : T(0);   // vector_norm backward: zero sub-gradient at 0
            const T iS = (S != T(0)) ? -w2 * D * rS : T(0);
            const T ga0 = iD * d0 + iS * s0, ga1 = iD * d1 + iS * s1, ga2 = iD * d2 + iS * s2;
            const T gb0 = iS * s0 - iD * d0, gb1 = iS * s1 - iD * d1, gb2 = iS * s2 - iD * d2;
            const T gaa = ga0 * a0 + ga1 * a1 + ga2 * a2;
            const T gbb = gb0 * b0 + gb1 * b1 + gb2 * b2;
            const T gh0 = (ga0 - a0 * gaa) * inh, gh1 = (ga1 - a1 * gaa) * inh, gh2 = (ga2 - a2 * gaa) * inh;
            const T g0 = (gb0 - b0 * gbb) * inP, g1 = (gb1 - b1 * gbb) * inP, g2 = (gb2 - b2 * gbb) * inP;
            gcx -= gh0;
            gcy -= gh1;
            gf += gh2;
            if (kDataGrad && dgrad) {  // h = (u - cx, v - cy, .): d error / d (u, v) = (gh0, gh1)
                dgrad[2 * i] += dcoef * gh0;
                dgrad[2 * i + 1] += dcoef * gh1;
            }
            T* row = contrib + kRow * i;
            if (m == 0) {
                row[0] = g0; row[1] = g1; row[2] = g2;
            } else {
                const T wg = o0 * g0 + o1 * g1 + o2 * g2;
                const T xg = x0 * g0 + x1 * g1 + x2 * g2;
                const T crg = c0 * g0 + c1 * g1 + c2 * g2;
                const T ow = oc * wg;
                row[0] = cn * g0 + ow * o0 + s * (g1 * o2 - g2 * o1);  // d/dXs: c g + oc (w.g) w + s (g x w)
                row[1] = cn * g1 + ow * o1 + s * (g2 * o0 - g0 * o2);
                row[2] = cn * g2 + ow * o2 + s * (g0 * o1 - g1 * o0);
                row[3] = g0; row[4] = g1; row[5] = g2;                 // d/dts
                const T gang = (doc * dot * wg + ds * crg - sn * xg) * rec;  // d/d angle, times 1/angle (0 at 0)
                const T od = oc * dot;
                row[6] = od * g0 + ow * x0 + s * (x1 * g2 - x2 * g1) + gang * o0;  // d/dw
                row[7] = od * g1 + ow * x1 + s * (x2 * g0 - x0 * g2) + gang * o1;
                row[8] = od * g2 + ow * x2 + s * (x0 * g1 - x1 * g0) + gang * o2;
            }
        }
        cost = warp_allreduce(cost);
        gf = warp_allreduce(gf);
        gcx = warp_allreduce(gcx);
        gcy = warp_allreduce(gcy);
        __syncwarp();
        // gather: the lane that owns parameter c (c = lane, lane + 32) adds the rows that touch c
        constexpr int kOwn = kParams > 0 ? (kParams + 31) / 32 : kWideMax / 32;  // components per lane
        T own[kOwn];
        T dsig = T(0);  // sum d/dXs . Xs + sum d/dts . ts  (-> d/d sigma)
#pragma unroll
        for (int h = 0; h < kOwn; ++h) {
            const int c = lane + 32 * h;
            T a = T(0);
            if (c < 3) {
                a = (c == 0) ? gf * dfp : (c == 1 ? gcx : gcy);
            } else if (c < tb) {
                const int j = (c - 3) / 3, k = (c - 3) - 3 * j;
                const T* q = contrib + kRow * j + k;
                for (int m = 0; m < V; ++m, q += kRow * N) a += *q;
                dsig = fma_t(a, th[c] * isig, dsig);
            } else if (c < n) {
                const int r = (c < wb) ? c - tb : c - wb;
                const int m = r / 3 + 1, k = r - 3 * (m - 1) + ((c < wb) ? 3 : 6);
                const T* q = contrib + kRow * (m * N) + k;
                for (int j = 0; j < N; ++j, q += kRow) a += *q;
                if (c < wb) dsig = fma_t(a, th[c] * isig, dsig);
            }
            own[h] = a;
        }
        const T gsig = -warp_allreduce(dsig) * isig;
#pragma unroll
        for (int h = 0; h < kOwn; ++h) {
            const int c = lane + 32 * h;
            if (c < n) {
                T a = own[h];
                if (c >= 3 && c < wb) {
                    const T v = th[c];
                    const T sg = v > T(0) ? T(1) : (v < T(0) ? T(-1) : T(0));  // d|x|/dx with sign(0) = 0
                    a = fma_t(a, isig, gsig * (c < tb ? wX : wT) * sg);
                }
                gout[c] = a;
            }
        }
        __syncwarp();
        return cost;
    }
};

}  // namespace davo
