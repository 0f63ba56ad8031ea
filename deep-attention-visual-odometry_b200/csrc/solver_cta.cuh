// solver_cta.cuh — one CTA (W warps) solves one JOINT problem (intrinsics + 6 pose parameters per
// view, n = 10 + 6V <= 64).  Same state machine as solver_warp.cuh / solver_wide.cuh, restated from
//   autograd_solvers/bfgs_solver.py:80-303, utils/func_inverse_curvature.py:8-11,
//   autograd_solvers/line_search/wolfe_conditions.py:23-253.
//
// Why a CTA here: one JOINT evaluation is V x N matches (1024 at config 3) and a 16K-problem batch
// gives a warp-per-problem grid only ~9 problems per resident warp, so a quarter of the launch was tail.
// With W = 4 warps per problem the per-problem latency drops ~4x, the CTA's staged points / observations /
// H / vectors are shared by the 4 warps (more CTAs fit per SM), and there are ~22 problems per CTA to
// balance.  Warp w evaluates views w, w+W, ...: a view's pose gradient never leaves its warp; only the
// 11 intrinsic sums cross warps (shared memory + one barrier).
//
// Every thread carries the same scalar state (f, alpha, lo, hi, counters ...): dot products over the
// n-vectors are recomputed per warp in the same order from shared memory, so the scalars are bitwise
// identical across the CTA and all control flow is CTA-uniform (barriers are safe).
#pragma once
#include "davo_common.cuh"
#include "objectives.cuh"
#include "solver_warp.cuh"  // LineSearchResult, same_bits
#include "solver_wide.cuh"  // warp_allreduce, wide_dot, kWideMax

#ifndef DAVO_JOINT_PACKED
// 1: evaluate the JOINT matches two per instruction (FFMA2).  Measured on config 3 (16K x 1024) with every
// occupancy that fits: never faster than the scalar loop (W = 2: 25.4-26.5 ms against 25.0 ms; W = 4: 28.7-35.7 ms
// against 27.3 ms): the 12 pose sums double the accumulator registers.  Kept for A/B builds.
#define DAVO_JOINT_PACKED 0
#endif

namespace davo {

template <typename T>
struct CtaWorkspace {
    T *x, *g, *gprev, *d, *s, *y, *yH, *Hy, *xt, *gt, *H;
    T* red;  // [W][16] per-warp intrinsic sums
    int ld;
    __host__ __device__ static size_t bytes(int n, int W) {
        return sizeof(T) * (10 * (size_t)kWideMax + (size_t)n * (n + 1) + (size_t)W * kSlots);
    }
    __device__ void carve(unsigned char* base, int n, int W) {
        T* p = reinterpret_cast<T*>(base);
        x = p; g = x + kWideMax; gprev = g + kWideMax; d = gprev + kWideMax; s = d + kWideMax;
        y = s + kWideMax; yH = y + kWideMax; Hy = yH + kWideMax; xt = Hy + kWideMax; gt = xt + kWideMax;
        red = gt + kWideMax;
        H = red + (size_t)W * kSlots;
        ld = n + 1;
    }
};

// The JOINT objective evaluated by a whole CTA.  th and gout are shared-memory vectors; the caller has
// made th visible to the CTA (barrier) before the call; on return gout is visible to the CTA.
// kV, kN > 0 fix views / points per view at compile time (config 3: 4 x 256, n = 34): loop bounds, the row
// stride of H and every shared-memory offset of the solver become constants.  0 = run time.
template <typename T, int W, bool kWeighted, int kV = 0, int kN = 0>
struct JointCtaObjective {
    static constexpr int kParams = kV > 0 ? 10 + 6 * kV : 0;  // compile-time n (0: run time)
    static constexpr int kSpecProbes = 1;                      // no speculative line search (line_search_cta)
    static constexpr bool kSpeculative = false;
    const SolveParams<T>& p;
    T* world;  // [N,3]
    T* obs;    // [V,N,2]
    T* wts;    // [V,N] (only if p.has_w)
    T* red;    // [W][16]
    uint64_t* bar;
    unsigned parity;
    int lane, warp, tid;

    // world points are padded to a multiple of 4 elements so that the observations start 16-byte aligned
    // (bulk-TMA destination, float2 / double2 reads) for every N
    __host__ __device__ static size_t world_elems(int N) { return ((size_t)N * 3 + 3) & ~size_t(3); }
    __host__ __device__ static size_t obs_elems(int N, int V) { return ((size_t)V * N * 2 + 3) & ~size_t(3); }
    __host__ __device__ static size_t data_bytes(int N, int V, bool has_w) {
        size_t b = sizeof(T) * (world_elems(N) + obs_elems(N, V) + (has_w ? (size_t)V * N : 0));
        return (b + 127) & ~size_t(127);
    }

    __device__ JointCtaObjective(const SolveParams<T>& p_, unsigned char* slab, T* red_, uint64_t* bar_)
        : p(p_), red(red_), bar(bar_), parity(0) {
        tid = threadIdx.x; lane = tid & 31; warp = tid >> 5;
        world = reinterpret_cast<T*>(slab);
        obs = world + world_elems(p.N);
        wts = obs + obs_elems(p.N, p.V);
    }

    __device__ __forceinline__ void init() {
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
    }

    // Stage problem b: two bulk TMA copies (world points, observations) completing on one mbarrier.
    __device__ __forceinline__ void bind(int b) {
        __syncthreads();  // every warp is done with the previous problem's slab
        const size_t wb = sizeof(T) * (size_t)p.N * 3, ob = sizeof(T) * (size_t)p.V * p.N * 2;
        const T* gw = p.data0 + (size_t)b * p.N * 3;
        const T* go = p.data1 + (size_t)b * p.V * p.N * 2;
        const bool bulk = (wb % 16 == 0) && (ob % 16 == 0);
        if (bulk) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(bar, (unsigned)(wb + ob));
                tma_load_1d(world, gw, (unsigned)wb, bar);
                tma_load_1d(obs, go, (unsigned)ob, bar);
            }
        } else {
            for (int i = tid; i < p.N * 3; i += 32 * W) world[i] = gw[i];
            for (int i = tid; i < p.V * p.N * 2; i += 32 * W) obs[i] = go[i];
        }
        if (kWeighted)
            for (int i = tid; i < p.V * p.N; i += 32 * W) wts[i] = p.w[(size_t)b * p.V * p.N + i];
        if (bulk) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        }
        __syncthreads();
    }

    __device__ __forceinline__ T eval(const T* th, T* gout) {
        using V2 = typename Vec2<T>::type;
        const int N = kN > 0 ? kN : p.N, V = kV > 0 ? kV : p.V;
        Intrinsics<T> I;
        I.load(th);
#if DAVO_JOINT_PACKED
        // Two matches per instruction (FFMA2 / FMUL2 / FADD2, see objectives.cuh): the kernel is bound by
        // instruction issue, not by the FMA pipe, so lane L works on the pair (i, i + 32) of every block of 64.
        using P = V2;
        P acc2[kPairAcc];
#pragma unroll
        for (int k = 0; k < kPairAcc; ++k) acc2[k] = pk(T(0));
#else
        T acc[kSlots];
#pragma unroll
        for (int k = 0; k < kSlots; ++k) acc[k] = T(0);
#endif
        for (int v = warp; v < V; v += W) {
            const T* pose = th + 10 + 6 * v;
            // sin/cos of this view's three Euler angles: lanes 0..2 compute, everyone receives
            T sn = T(0), cs = T(0);
            if (lane < 3) sincos(pose[lane], &sn, &cs);  // one shared range reduction
            const T sx = shfl_idx(sn, 0), cx = shfl_idx(cs, 0), sy = shfl_idx(sn, 1), cy = shfl_idx(cs, 1),
                    sz = shfl_idx(sn, 2), cz = shfl_idx(cs, 2);
            const T t0 = pose[3], t1 = pose[4], t2 = pose[5];
            // R = Rz Ry Rx, distorted_camera_model.py:38-55
            const T r00 = cy * cz, r01 = sx * sy * cz - cx * sz, r02 = cx * sy * cz + sx * sz;
            const T r10 = cy * sz, r11 = sx * sy * sz + cx * cz, r12 = cx * sy * sz - sx * cz;
            const T r20 = -sy, r21 = sx * cy, r22 = cx * cy;
            T part[kSlots];  // 0..2: sum X' x gX', 3..5: g_t = sum gX'
#pragma unroll
            for (int k = 0; k < kSlots; ++k) part[k] = T(0);
            const V2* ob = reinterpret_cast<const V2*>(obs) + (size_t)v * N;
            const T* wv = wts + (size_t)v * N;
#if DAVO_JOINT_PACKED
            P part2[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) part2[k] = pk(T(0));
            for (int i = lane; i < N; i += 64) {
                const bool second_valid = i + 32 < N;
                const int i1 = second_valid ? i + 32 : i;  // a missing second match re-reads the first; its residuals are zeroed
                P X, Y, Z, nus, nvs, w2 = pk(T(1));
                X.x = world[3 * i]; Y.x = world[3 * i + 1]; Z.x = world[3 * i + 2];
                X.y = world[3 * i1]; Y.y = world[3 * i1 + 1]; Z.y = world[3 * i1 + 2];
                const V2 o0 = ob[i], o1 = ob[i1];
                nus.x = -o0.x; nus.y = -o1.x; nvs.x = -o0.y; nvs.y = -o1.y;
                if (kWeighted) { w2.x = wv[i]; w2.y = second_valid ? wv[i1] : T(0); }
                const P xp = pfma(X, pk(r00), pfma(Y, pk(r01), pfma(Z, pk(r02), pk(t0))));
                const P yp = pfma(X, pk(r10), pfma(Y, pk(r11), pfma(Z, pk(r12), pk(t1))));
                const P zr = pfma(X, pk(r20), pfma(Y, pk(r21), pfma(Z, pk(r22), pk(t2))));
                P zp = zr;
                if (zp.x == T(0)) zp.x += T(1e-8);  // :57
                if (zp.y == T(0)) zp.y += T(1e-8);
                P iz;
                iz.x = div_rn(T(1), zp.x);
                iz.y = div_rn(T(1), zp.y);
                const P a = pmul(xp, iz), b = pmul(yp, iz);
                P gu, gv;
                match_pair_cost_grad<T, kWeighted>(I, a, b, nus, nvs, w2, second_valid, acc2, gu, gv);
                const P gA = pmul(gu, pk(I.fx));
                const P gB = pfma(gu, pk(I.s), pmul(gv, pk(I.fy)));
                const P gx = pmul(gA, iz), gy = pmul(gB, iz);
                const P gz = pmul(pfma(gA, a, pmul(gB, b)), pmul(pk(T(-1)), iz));
                const P ngx = pmul(pk(T(-1)), gx), ngy = pmul(pk(T(-1)), gy), ngz = pmul(pk(T(-1)), gz);
                part2[0] = pfma(yp, gz, pfma(zr, ngy, part2[0]));   // X' x gX'
                part2[1] = pfma(zr, gx, pfma(xp, ngz, part2[1]));
                part2[2] = pfma(xp, gy, pfma(yp, ngx, part2[2]));
                part2[3] = padd(part2[3], gx); part2[4] = padd(part2[4], gy); part2[5] = padd(part2[5], gz);
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) part[k] = part2[k].x + part2[k].y;
#else
#pragma unroll 2
            for (int i = lane; i < N; i += 32) {
                const T X = world[3 * i], Y = world[3 * i + 1], Z = world[3 * i + 2];
                const V2 o = ob[i];
                const T xp = fma_t(X, r00, fma_t(Y, r01, fma_t(Z, r02, t0)));
                const T yp = fma_t(X, r10, fma_t(Y, r11, fma_t(Z, r12, t1)));
                const T zr = fma_t(X, r20, fma_t(Y, r21, fma_t(Z, r22, t2)));
                T zp = zr;
                if (zp == T(0)) zp += T(1e-8);  // :57
                const T iz = div_rn(T(1), zp);
                const T a = xp * iz, b = yp * iz;
                T gu, gv;
                match_cost_grad<T, kWeighted>(I, a, b, o.x, o.y, kWeighted ? wv[i] : T(1), acc, gu, gv);
                const T gA = gu * I.fx;
                const T gB = fma_t(gu, I.s, gv * I.fy);
                const T gx = gA * iz, gy = gB * iz;
                const T gz = -fma_t(gA, a, gB * b) * iz;
                part[0] = fma_t(yp, gz, fma_t(-zr, gy, part[0]));   // X' x gX'
                part[1] = fma_t(zr, gx, fma_t(-xp, gz, part[1]));
                part[2] = fma_t(xp, gy, fma_t(-yp, gx, part[2]));
                part[3] += gx; part[4] += gy; part[5] += gz;
            }
#endif
            // d(cost/2)/dr_k = sum gX' . dX'/dr_k with dX'/dr_k = a_k x (R X): the derivative of R = Rz Ry Rx with
            // respect to an Euler angle is the cross product with that angle's (rotated) axis, a_x = Rz Ry e_x,
            // a_y = Rz e_y, a_z = e_z.  So sum gX' . (a_k x RX) = a_k . sum (RX x gX'), and with RX = X' - t:
            // c = sum X' x gX' - t x sum gX'.  Three accumulators instead of the nine of sum gX' (x) X.
            const T mine = reduce_scatter16<true>(part, lane);  // total of slot lane >> 1
            const T c0 = shfl_idx(mine, 0), c1 = shfl_idx(mine, 2), c2 = shfl_idx(mine, 4), g0 = shfl_idx(mine, 6),
                    g1 = shfl_idx(mine, 8), g2 = shfl_idx(mine, 10);
            if (lane < 3) {
                const T cx_ = c0 - (t1 * g2 - t2 * g1), cy_ = c1 - (t2 * g0 - t0 * g2), cz_ = c2 - (t0 * g1 - t1 * g0);
                T out;
                if (lane == 0) out = cz * cy * cx_ + sz * cy * cy_ - sy * cz_;   // a_x = (cz cy, sz cy, -sy)
                else if (lane == 1) out = cz * cy_ - sz * cx_;                   // a_y = (-sz, cz, 0)
                else out = cz_;                                                  // a_z = (0, 0, 1)
                gout[10 + 6 * v + lane] = T(2) * out;  // least_squares_utils.py:43
            }
            if (lane >= 6 && lane < 12 && !(lane & 1)) gout[10 + 6 * v + 3 + ((lane - 6) >> 1)] = T(2) * mine;  // d/dt
        }
#if DAVO_JOINT_PACKED
        T acc[kSlots];
#pragma unroll
        for (int k = 0; k < kPairAcc; ++k) acc[k] = acc2[k].x + acc2[k].y;
#pragma unroll
        for (int k = kPairAcc; k < kSlots; ++k) acc[k] = T(0);
#endif
        fold_uv_terms(acc);
        const T mine = reduce_scatter16<true>(acc, lane);
        if (!(lane & 1)) red[warp * kSlots + (lane >> 1)] = mine;
        __syncthreads();
        T f = T(0);
#pragma unroll
        for (int w = 0; w < W; ++w) f += red[w * kSlots + 10];
        if (tid < 10) {
            T gsum = T(0);
#pragma unroll
            for (int w = 0; w < W; ++w) gsum += red[w * kSlots + tid];
            gout[tid] = T(2) * gsum;
        }
        __syncthreads();
        return f;
    }
};

// The DISTORT10 objective (objectives.cuh) evaluated by a CTA: the second launch of a DISTORT10 solve gives each
// straggler a whole CTA, because there the LATENCY of one evaluation — not the batch's throughput — sets the launch
// time.  The CTA is G groups of W warps.  One evaluation is the work of ONE group (thread t of the group handles the
// match pair (t, t + 32 W) of every block of 64 W matches; the 11 sums are reduced inside each warp and across the
// group's warps through `red` with one barrier); the G groups exist for the speculative line search below, where each
// group evaluates a different trial point.  Results do not depend on G.
#ifndef DAVO_SPEC_PROBES
#define DAVO_SPEC_PROBES 8  // G: groups of warps, each evaluating its own trial points (1 x 1: no speculation)
#endif
#ifndef DAVO_SPEC_PER_GROUP
#define DAVO_SPEC_PER_GROUP 4  // trial points each group evaluates per round, one after the other
#endif

template <typename T, int W, int G = 1>
struct Distort10CtaObjective {
    static constexpr int kParams = 10;
    static constexpr int kPerGroup = DAVO_SPEC_PER_GROUP;
    static constexpr int kSpecProbes = G * kPerGroup;   // trial points per round
    static constexpr bool kSpeculative = true;
    static constexpr int kKeep = 11;  // folded per-thread sums of one probe: 10 gradient sums + cost
    using V4 = typename Vec4<T>::type;
    using P = typename Vec2<T>::type;
    const SolveParams<T>& p;
    V4* matches;  // [N] {a, b, -u*, -v*} (observations negated once per problem)
    T* red;       // [W G][16]
    uint64_t* bar;
    unsigned parity;
    int lane, warp, tid, grp;  // warp, tid: within the group

    __host__ __device__ static size_t data_bytes(int N, int, bool) {
        return (sizeof(V4) * (size_t)N + 127) & ~size_t(127);
    }

    __device__ Distort10CtaObjective(const SolveParams<T>& p_, unsigned char* slab, T* red_, uint64_t* bar_)
        : p(p_), matches(reinterpret_cast<V4*>(slab)), red(red_), bar(bar_), parity(0) {
        grp = threadIdx.x / (32 * W);
        tid = threadIdx.x - grp * 32 * W; lane = tid & 31; warp = tid >> 5;
    }

    __device__ __forceinline__ void init() {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
    }

    __device__ __forceinline__ void bind(int b) {
        fence_proxy_async();
        __syncthreads();  // every warp is done with the previous problem's slab
        if (threadIdx.x == 0) {
            const unsigned bytes = (unsigned)(sizeof(V4) * (size_t)p.N);
            mbar_expect_tx(bar, bytes);
            tma_load_1d(matches, p.data0 + (size_t)b * p.N * 4, bytes, bar);
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        for (int i = threadIdx.x; i < p.N; i += 32 * W * G) {
            V4 m = matches[i];
            m.z = -m.z;
            m.w = -m.w;
            matches[i] = m;
        }
        __syncthreads();
    }

    // this thread's share of one evaluation at the parameter line `th`
    __device__ __forceinline__ void accumulate(const T* th, T (&acc)[kSlots]) {
        Intrinsics<T> I;
        I.load(th);
        P acc2[kPairAcc];
#pragma unroll
        for (int k = 0; k < kPairAcc; ++k) acc2[k] = pk(T(0));
        P gu, gv;
        const int N = p.N;
        for (int i = tid; i < N; i += 64 * W) {
            const V4 m0 = matches[i];
            const bool second_valid = i + 32 * W < N;
            V4 m1;
            m1.x = m1.y = m1.z = m1.w = T(0);
            if (second_valid) m1 = matches[i + 32 * W];
            P a, b, nus, nvs;
            a.x = m0.x; a.y = m1.x; b.x = m0.y; b.y = m1.y;
            nus.x = m0.z; nus.y = m1.z; nvs.x = m0.w; nvs.y = m1.w;
            match_pair_cost_grad<T, false>(I, a, b, nus, nvs, pk(T(1)), second_valid, acc2, gu, gv);
        }
#pragma unroll
        for (int k = 0; k < kPairAcc; ++k) acc[k] = acc2[k].x + acc2[k].y;
#pragma unroll
        for (int k = kPairAcc; k < kSlots; ++k) acc[k] = T(0);
        fold_uv_terms(acc);
    }

    // One evaluation (group 0 computes; every thread of the CTA gets f, gout is visible to the CTA on return).
    __device__ __forceinline__ T eval(const T* th, T* gout) {
        if (grp == 0) {
            T acc[kSlots];
            accumulate(th, acc);
            const T mine = reduce_scatter16<true>(acc, lane);
            if (!(lane & 1)) red[warp * kSlots + (lane >> 1)] = mine;
        }
        __syncthreads();
        T f = T(0);
#pragma unroll
        for (int w = 0; w < W; ++w) f += red[w * kSlots + 10];
        if (threadIdx.x < 10) {
            T gsum = T(0);
#pragma unroll
            for (int w = 0; w < W; ++w) gsum += red[w * kSlots + threadIdx.x];
            gout[threadIdx.x] = T(2) * gsum;  // least_squares_utils.py:43
        }
        __syncthreads();
        return f;
    }

    // ---- speculative line search (line_search_cta_spec) ----------------------------------------------------
    // G trial points per round, one per group.  A probe reduces only TWO scalars — the cost and d . grad — and keeps
    // its 11 folded sums in registers; the full gradient is reduced once, for the probe the search accepts
    // (finish_gradient).  xt: G parameter lines of 16 entries; d2[c] = 2 d_c (least_squares_utils.py:43's factor
    // folded in).  Leaves this group's per-warp partial (cost, d . grad) in red[(grp W + warp) 2 + {0, 1}].
    __device__ __forceinline__ void probe_partial(const T* xt, const T (&d2)[10], T (&keep)[kPerGroup][kKeep]) {
#pragma unroll
        for (int q = 0; q < kPerGroup; ++q) {
            const int j = q * G + grp;   // this group's q-th trial point
            T acc[kSlots];
            accumulate(xt + 16 * j, acc);
#pragma unroll
            for (int k = 0; k < kKeep; ++k) keep[q][k] = acc[k];
            T e = d2[0] * acc[0], o = d2[1] * acc[1];
#pragma unroll
            for (int c = 2; c < 10; c += 2) {
                e = fma_t(d2[c], acc[c], e);
                o = fma_t(d2[c + 1], acc[c + 1], o);
            }
            T pf = acc[10], pd = e + o;
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                pf += shfl_xor(pf, m);
                pd += shfl_xor(pd, m);
            }
            if (lane == 0) {
                red[(j * W + warp) * 2] = pf;
                red[(j * W + warp) * 2 + 1] = pd;
            }
        }
    }
    // (cost, d . grad) of probe j from the partials (every thread that calls it adds them in the same order)
    __device__ __forceinline__ void probe_result(int j, T& f, T& dphi) const {
        T fs = T(0), ds = T(0);
#pragma unroll
        for (int w = 0; w < W; ++w) {
            fs += red[(j * W + w) * 2];
            ds += red[(j * W + w) * 2 + 1];
        }
        f = fs;
        dphi = ds;
    }

    // The gradient of trial point `sel` of the last probe_partial round: the tail of eval().
    __device__ __forceinline__ void finish_gradient(const T (&keep)[kPerGroup][kKeep], int sel_probe, T* gout) {
        const int sel = sel_probe % G, q_sel = sel_probe / G;
        T acc[kSlots];
#pragma unroll
        for (int k = 0; k < kKeep; ++k) {
            T v = keep[0][k];
#pragma unroll
            for (int q = 1; q < kPerGroup; ++q) v = (q_sel == q) ? keep[q][k] : v;
            acc[k] = v;
        }
#pragma unroll
        for (int k = kKeep; k < kSlots; ++k) acc[k] = T(0);
        const T mine = reduce_scatter16<true>(acc, lane);
        __syncthreads();  // every thread has read the probe round's `red`
        if (grp == sel && !(lane & 1)) red[warp * kSlots + (lane >> 1)] = mine;
        __syncthreads();
        if (threadIdx.x < 10) {
            T gsum = T(0);
#pragma unroll
            for (int w = 0; w < W; ++w) gsum += red[w * kSlots + threadIdx.x];
            gout[threadIdx.x] = T(2) * gsum;  // least_squares_utils.py:43
        }
        __syncthreads();
    }
};

// out[c] = scale * sum_j H[c][j] v[j].  n is small (34 at config 3), so a whole warp per row would idle
// most lanes and pay a 5-step butterfly per row; instead FOUR threads share a row (j = q, q+4, ...) and
// combine with two shuffles, i.e. 8 rows per warp and 32 rows per pass of a 4-warp CTA.
template <typename T, int W>
__device__ __forceinline__ void cta_matvec(const T* H, int ld, const T* v, T* out, int n, T scale) {
    const int tid = threadIdx.x, q = tid & 3;
    constexpr int kRowsPerPass = 32 * W / 4;
    for (int c0 = 0; c0 < n; c0 += kRowsPerPass) {   // CTA-uniform trip count: the shuffles below are safe
        const int c = c0 + (tid >> 2);
        T a = T(0);
        if (c < n)
            for (int j = q; j < n; j += 4) a = fma_t(H[c * ld + j], v[j], a);
        a += shfl_xor(a, 1);
        a += shfl_xor(a, 2);
        if (c < n && q == 0) out[c] = scale * a;
    }
}

// One sweep over H: out_a = H a, out_b = H b (four threads share a ROW) and out_t = a^T H (the same four
// threads share the COLUMN with the row's index; H rows are padded to n+1 words, so both walks are at most
// 2-way bank conflicted).
template <typename T, int W>
__device__ __forceinline__ void cta_matvec3(const T* H, int ld, const T* a, const T* b, T* out_a, T* out_b,
                                            T* out_t, int n) {
    const int tid = threadIdx.x, q = tid & 3;
    constexpr int kRowsPerPass = 32 * W / 4;
    for (int c0 = 0; c0 < n; c0 += kRowsPerPass) {   // CTA-uniform trip count: the shuffles below are safe
        const int c = c0 + (tid >> 2);
        T sa = T(0), sb = T(0), st = T(0);
        if (c < n)
            for (int j = q; j < n; j += 4) {
                const T h = H[c * ld + j];
                sa = fma_t(h, a[j], sa);
                sb = fma_t(h, b[j], sb);
                st = fma_t(a[j], H[j * ld + c], st);
            }
        sa += shfl_xor(sa, 1); sb += shfl_xor(sb, 1); st += shfl_xor(st, 1);
        sa += shfl_xor(sa, 2); sb += shfl_xor(sb, 2); st += shfl_xor(st, 2);
        if (c < n && q == 0) {
            out_a[c] = sa;
            out_b[c] = sb;
            out_t[c] = st;
        }
    }
}

template <typename T, int W, typename Obj>
__device__ __forceinline__ LineSearchResult<T> line_search_cta(Obj& obj, const SolveParams<T>& p, const T* x,
                                                               const T* d, T f0, const T* g, T* xt, T* gt) {
    const int n = Obj::kParams > 0 ? Obj::kParams : p.n, tid = threadIdx.x, lane = tid & 31;
    const T g0 = wide_dot(d, g, n, lane);       // :77 (each warp, same order: identical in every thread)
    bool widening = true, zooming = false;      // :80-82
    T lo = T(0), hi = T(0), cand = T(1);        // :97-108
    T lo_f = f0, hi_f = f0, cand_f = f0;        // :109-111
    int probes = 0;
    const T neg_c2_g0 = mul_rn(T(-1) * p.c2, g0);
    for (int i = 0; i < p.max_ls; ++i) {        // :116
        if (!(widening || zooming)) break;      // :119-121
        if (i > 0) {
            if (widening) { hi = cand; hi_f = cand_f; cand = mul_rn(T(2), cand); }  // :125-127
            if (zooming) cand = mul_rn(T(0.5), add_rn(lo, hi));                      // :128-131
        }
        if (tid < n) xt[tid] = add_rn(x[tid], mul_rn(cand, d[tid]));                 // :139
        __syncthreads();
        cand_f = obj.eval(xt, gt);
        const T dphi = wide_dot(d, gt, n, lane);                                    // :141
        ++probes;
        bool D = cand_f > add_rn(f0, mul_rn(mul_rn(p.c1, cand), g0));               // :146-150
        if (zooming) D = D || (cand_f >= lo_f);                                     // :151-153
        if (widening && i > 0) D = D || (cand_f >= hi_f);                           // :154-157
        const bool C = p.strong ? (fabs(dphi) <= neg_c2_g0) : (mul_rn(T(-1), dphi) <= neg_c2_g0);  // :160-169
        const bool G = widening ? (dphi >= T(0)) : (mul_rn(dphi, sub_rn(hi, lo)) >= T(0));         // :174-180
        if (zooming) {                                                              // :187-207
            if (D) { hi = cand; hi_f = cand_f; }
            else if (C) { hi = lo = cand; hi_f = lo_f = cand_f; zooming = false; }
            else { if (G) { hi = lo; hi_f = lo_f; } lo = cand; lo_f = cand_f; }
        } else {                                                                    // :216-237
            if (D) { lo = hi; lo_f = hi_f; hi = cand; hi_f = cand_f; widening = false; zooming = true; }
            else if (C) { hi = lo = cand; hi_f = lo_f = cand_f; widening = false; }
            else if (G) { lo = cand; lo_f = cand_f; widening = false; zooming = true; }
        }
        if (zooming && !(lo != hi)) zooming = false;                                // :236
    }
    LineSearchResult<T> r;
    r.alpha = hi; r.last_cand = cand; r.last_f = cand_f; r.last_g = T(0); r.probes = probes;
    return r;
}

// One probe of wolfe_conditions.py:23-239 consumed by the state machine (the body of line_search_cta's loop), with the
// top-of-loop update of the NEXT trial step folded in.  Returns false when the search has ended.
template <typename T>
struct WolfeState {
    bool widening, zooming;
    T lo, hi, cand, lo_f, hi_f, cand_f, f0, g0, neg_c2_g0;
    int i, probes;
    __device__ __forceinline__ void start(T f0_, T g0_, T c2) {                      // :77-114
        widening = true; zooming = false;
        lo = T(0); hi = T(0); cand = T(1);
        f0 = lo_f = hi_f = cand_f = f0_;
        g0 = g0_;
        neg_c2_g0 = mul_rn(T(-1) * c2, g0_);
        i = 0; probes = 0;
    }
    __device__ __forceinline__ bool consume(const SolveParams<T>& p, T f, T dphi) {  // :143-237, then :116-131
        cand_f = f;
        ++probes;
        bool D = cand_f > add_rn(f0, mul_rn(mul_rn(p.c1, cand), g0));               // :146-150
        if (zooming) D = D || (cand_f >= lo_f);                                     // :151-153
        if (widening && i > 0) D = D || (cand_f >= hi_f);                           // :154-157
        const bool C = p.strong ? (fabs(dphi) <= neg_c2_g0) : (mul_rn(T(-1), dphi) <= neg_c2_g0);  // :160-169
        const bool G = widening ? (dphi >= T(0)) : (mul_rn(dphi, sub_rn(hi, lo)) >= T(0));         // :174-180
        if (zooming) {                                                              // :187-207
            if (D) { hi = cand; hi_f = cand_f; }
            else if (C) { hi = lo = cand; hi_f = lo_f = cand_f; zooming = false; }
            else { if (G) { hi = lo; hi_f = lo_f; } lo = cand; lo_f = cand_f; }
        } else {                                                                    // :216-237
            if (D) { lo = hi; lo_f = hi_f; hi = cand; hi_f = cand_f; widening = false; zooming = true; }
            else if (C) { hi = lo = cand; hi_f = lo_f = cand_f; widening = false; }
            else if (G) { lo = cand; lo_f = cand_f; widening = false; zooming = true; }
        }
        if (zooming && !(lo != hi)) zooming = false;                                // :236
        ++i;
        if (!(widening || zooming) || i >= p.max_ls) return false;                  // :116-121
        if (widening) { hi = cand; hi_f = cand_f; cand = mul_rn(T(2), cand); }      // :125-127
        if (zooming) cand = mul_rn(T(0.5), add_rn(lo, hi));                          // :128-131
        return true;
    }
};

// line_search_cta with K = Obj::kSpecProbes trial points per round, one per group of warps
// (Distort10CtaObjective::probe_partial).  The state machine is wolfe_conditions.py:23-239 unchanged, consumed one
// probe at a time; what is speculative is only WHICH points get evaluated: next to the probe the search asks for, the
// points it will ask for if this probe and its successors fail the sufficient-decrease test (the "D" branch,
// :146-157 -> :187-189 / :216-219: the bracket's upper end moves to the probe and the next one bisects again).  A
// straggler's line searches are such chains, ~45 probes long, repeated for 1000 iterations (a float32 problem that
// has stopped making progress: alpha ~ 4e-6 every time).  A probe is consumed only if its trial step is bitwise the
// one the state machine computed, so the sequence of consumed probes, every decision and the result are those of
// the one-probe-at-a-time search; unconsumed probes are discarded.
//
// Every warp evaluating and then replaying the state machine redundantly made a round cost as many issue slots as
// G sequential probes (ncu: issue slots 60 % busy, no gain).  So the bookkeeping is warp 0's alone, between two
// barriers: it consumes probe 0 with the general code, tests the rest of the chain in PARALLEL (lane j: "is probe j
// another plain D step?" — along a D chain lo, f(lo) and the test's constants do not change), takes the leading run
// in one go, consumes the probe that broke the chain with the general code, and writes the next round's trial points.
// On return gt holds the gradient at the last consumed probe if that is the accepted point.  xt: 16 K words for the
// trial points; ctl: 8 words of scratch.
template <typename T, int W, typename Obj>
__device__ __forceinline__ LineSearchResult<T> line_search_cta_spec(Obj& obj, const SolveParams<T>& p, const T* x,
                                                                    const T* d, T f0, const T* g, T* xt, T* gt,
                                                                    T* ctl) {
    constexpr int K = Obj::kSpecProbes;
    constexpr int n = Obj::kParams;
    static_assert(n == 10 && 16 * K <= 4 * kWideMax && K <= 32, "speculative probes: 10 parameters, K lines of 16 in xt");
    const int tid = threadIdx.x, lane = tid & 31;
    const bool boss = tid < 32;  // warp 0 owns the state machine
    T d2[10];
#pragma unroll
    for (int c = 0; c < 10; ++c) d2[c] = T(2) * d[c];
    T keep[Obj::kPerGroup][Obj::kKeep];
    WolfeState<T> st;
    st.start(f0, wide_dot(d, g, n, lane), p.c2);
    T c_mine = T(0);  // warp 0, lane j < K: trial step of probe j in the current round
    int last = 0;
    // trial points of a round: the step the search asks for, then its successors along the D chain
    auto write_round = [&]() {
        bool z = st.zooming;
        T slo = st.lo, shi = st.hi, sc = st.cand;
        c_mine = sc;
#pragma unroll
        for (int j = 1; j < K; ++j) {
            if (z) { shi = sc; } else { slo = shi; shi = sc; z = true; }   // the D branch
            sc = mul_rn(T(0.5), add_rn(slo, shi));
            if (lane == j) c_mine = sc;
        }
#pragma unroll
        for (int e0 = 0; e0 < 16 * K; e0 += 32) {   // warp-uniform trip count: the shuffle needs every lane
            const int e = e0 + lane, col = e & 15;
            const T cj = shfl_idx(c_mine, (e >> 4) < K ? (e >> 4) : 0);
            if (e < 16 * K) xt[e] = (col < n) ? add_rn(x[col], mul_rn(cj, d[col])) : T(0);   // :139
        }
    };
    if (boss) {
        const bool go = p.max_ls > 0;
        if (go) write_round();
        if (lane == 0) ctl[0] = go ? T(1) : T(0);
    }
    __syncthreads();
    while (ctl[0] != T(0)) {
        obj.probe_partial(xt, d2, keep);
        __syncthreads();
        if (boss) {
            T fl = T(0), dl = T(0);
            if (lane < K) obj.probe_result(lane, fl, dl);
            bool go = st.consume(p, shfl_idx(fl, 0), shfl_idx(dl, 0));
            last = 0;
            if (K > 1 && go && st.zooming && same_bits(st.cand, shfl_idx(c_mine, 1))) {
                // lanes 1 .. K-1: is probe j another plain D step (bracket end moves to it, search goes on)?
                bool D = fl > add_rn(st.f0, mul_rn(mul_rn(p.c1, c_mine), st.g0));   // :146-150
                D = D || (fl >= st.lo_f);                                            // :151-153
                // probe j >= 2 sits on the chain only if it bisects [lo, c_{j-1}] for the lo of NOW (probe 0 may have
                // moved lo and still landed on c_1: widening, no decrease failure, positive slope -> [1, 0] -> 0.5)
                const T c_prev = __shfl_up_sync(kFull, c_mine, 1);
                const bool on_chain = lane == 1 || same_bits(c_mine, mul_rn(T(0.5), add_rn(st.lo, c_prev)));
                const unsigned chain = __ballot_sync(kFull, lane >= 1 && lane < K && on_chain) >> 1;
                const bool plain = lane >= 1 && lane < K && D && (st.lo != c_mine) && (st.i + lane < p.max_ls);
                const unsigned m = (__ballot_sync(kFull, plain) >> 1) & chain;
                int run = __ffs(~m) - 1;        // leading plain-D probes after probe 0 ...
                const int reach = __ffs(~chain) - 1;  // ... among the probes that sit on the chain
                if (run > reach) run = reach;
                if (run > 0) {
                    st.hi = shfl_idx(c_mine, run);
                    st.hi_f = st.cand_f = shfl_idx(fl, run);
                    st.probes += run;
                    st.i += run;
                    st.cand = mul_rn(T(0.5), add_rn(st.lo, st.hi));                  // :128-131
                    last = run;
                }
                if (run + 1 < K && run < reach) {  // the probe that broke the chain sits at the step the search asks for now
                    last = run + 1;
                    go = st.consume(p, shfl_idx(fl, run + 1), shfl_idx(dl, run + 1));
                }
            }
            if (go) {
                write_round();
            } else if (lane == 0) {
                ctl[1] = st.hi; ctl[2] = st.cand; ctl[3] = st.cand_f;
                ctl[4] = T(st.probes); ctl[5] = T(last);
            }
            if (lane == 0) ctl[0] = go ? T(1) : T(0);
        }
        __syncthreads();
    }
    LineSearchResult<T> r;
    r.alpha = ctl[1]; r.last_cand = ctl[2]; r.last_f = ctl[3]; r.last_g = T(0); r.probes = (int)ctl[4];
    last = (int)ctl[5];
    if (p.max_ls <= 0) { r.alpha = T(0); r.last_cand = T(1); r.last_f = f0; r.probes = 0; }
    if (r.probes > 0 && same_bits(r.alpha, r.last_cand)) obj.finish_gradient(keep, last, gt);
    return r;
}

// bfgs_solver.py:80-215 for one problem, executed by the whole CTA.
template <typename T, int W, typename Obj>
__device__ __forceinline__ void solve_one_cta(Obj& obj, const SolveParams<T>& p, int b, CtaWorkspace<T>& ws) {
    const int n = Obj::kParams > 0 ? Obj::kParams : p.n, ld = Obj::kParams > 0 ? Obj::kParams + 1 : ws.ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kThreads = 32 * W;
    T *x = ws.x, *g = ws.g, *gprev = ws.gprev, *gt = ws.gt;
    T *d = ws.d, *s = ws.s, *y = ws.y, *yH = ws.yH, *Hy = ws.Hy, *H = ws.H;
    if (tid < n) {
        x[tid] = p.x0[(size_t)b * n + tid];
        s[tid] = T(0);
        g[tid] = T(0);
    }
    for (int i = warp; i < n; i += W)
        for (int j = lane; j < n; j += 32) H[i * ld + j] = (i == j) ? T(1) : T(0);  // :112-117
    __syncthreads();
    T f = T(0);
    int iters = 0, fevals = 0, reason = DAVO_REASON_CAP;
    bool have_fg = false, have_f = false;

    for (int k = 0; k < p.max_iters; ++k) {  // :118
        if (!have_fg) f = obj.eval(x, g);    // :128-135
        ++fevals;
        have_f = true;
        if (!(f > p.thr)) {                  // :143
            reason = (f <= p.thr) ? DAVO_REASON_THRESHOLD : DAVO_REASON_NAN;
            break;
        }
        if (k == 0) {
            if (tid < n) d[tid] = mul_rn(T(-1), g[tid]);  // :152-155
        } else {
            if (tid < n) y[tid] = sub_rn(g[tid], gprev[tid]);  // :157
            __syncthreads();
            const T sy = wide_dot(s, y, n, lane);
            if (k == 1) {                                      // :159-167, :217-233
                T den = wide_dot(y, y, n, lane);
                den = (den < T(1e-5)) ? T(1e-5) : den;
                T sc = div_rn(sy, den);
                sc = (sc < T(1e-4)) ? T(1e-4) : sc;
                for (int i = warp; i < n; i += W)
                    for (int j = lane; j < n; j += 32) H[i * ld + j] = mul_rn(sc, H[i * ld + j]);
                __syncthreads();
            }
            T rho = div_rn(T(1), sy);                          // func_inverse_curvature.py:8-11
            if (sy <= T(0)) rho = T(0);
            // H' = H + (s rho) s^T (1+q) - (s rho)(y^T H) - (H y)(s rho)^T with the old H on the right
            // (:263-303).  y^T H is formed from the COLUMNS of H — never replaced by (H y)^T: H is symmetric
            // only up to rounding and that shortcut makes the asymmetry grow in float32 (see solver_warp.cuh).
            // One sweep over H forms H y, H g (rows) and y^T H (columns); the new direction -H' g (:173-176)
            // then follows from
            //   H' g = H g + (s rho) [ (1+q) s.g - (y^T H).g ] - (H y) rho s.g
            // without waiting for H' to be written, so the update of H and of d share one barrier.
            cta_matvec3<T, W>(H, ld, y, g, Hy, ws.xt /* scratch: H g */, yH, n);
            __syncthreads();
            const T* Hg = ws.xt;
            T a0 = T(0), a1 = T(0), a2 = T(0);
            for (int c = lane; c < n; c += 32) {
                a0 = fma_t(yH[c], y[c], a0);
                a1 = fma_t(s[c], g[c], a1);
                a2 = fma_t(yH[c], g[c], a2);
            }
            const T q = mul_rn(warp_allreduce(a0), rho);       // y^T H y / (y^T s), :271-274
            const T sg = warp_allreduce(a1), yhg = warp_allreduce(a2);
            const T onepq = add_rn(T(1), q);
            if (tid < n) {
                const T sr = mul_rn(s[tid], rho);
                const T hpg = fma_t(-Hy[tid] * rho, sg, fma_t(sr, fma_t(onepq, sg, -yhg), Hg[tid]));
                d[tid] = mul_rn(T(-1), hpg);
            }
            {   // four threads share a row (as in cta_matvec2): the row's constants are loaded once
                const int qd = tid & 3;
                constexpr int kRowsPerPass = kThreads / 4;
                for (int i0 = 0; i0 < n; i0 += kRowsPerPass) {
                    const int i = i0 + (tid >> 2);
                    if (i < n) {
                        const T sri = mul_rn(s[i], rho), nHyrho = -mul_rn(Hy[i], rho);
                        for (int j = qd; j < n; j += 4) {
                            const T inner = fma_t(s[j], onepq, -yH[j]);   // s_j (1+q) - (y^T H)_j
                            H[i * ld + j] = fma_t(nHyrho, s[j], fma_t(sri, inner, H[i * ld + j]));
                        }
                    }
                }
            }
        }
        __syncthreads();
        LineSearchResult<T> ls;                                                                // :181-190
        if constexpr (Obj::kSpeculative) ls = line_search_cta_spec<T, W>(obj, p, x, d, f, g, ws.s /* s, y, yH, Hy: 4 kWideMax words, free here */, gt, ws.xt);
        else ls = line_search_cta<T, W>(obj, p, x, d, f, g, ws.xt, gt);
        fevals += ls.probes;
        ++iters;
        if (tid < n) {                                         // :191-199
            const T sc = mul_rn(ls.alpha, d[tid]);
            s[tid] = sc;
            x[tid] = add_rn(x[tid], sc);
        }
        __syncthreads();
        const T nrm = sqrt_rn(wide_dot(s, s, n, lane));        // :203-205
        have_fg = same_bits(ls.alpha, ls.last_cand);
        have_f = have_fg;
        {   // rotate gradient buffers: gprev <- g, and g <- gt when the accepted point is the last probe
            T* old = gprev;
            gprev = g;
            if (have_fg) { g = gt; gt = old; f = ls.last_f; }
            else { g = old; }
        }
        if (!(nrm > p.min_step)) {                             // :203-207
            reason = DAVO_REASON_STEP;
            break;
        }
    }
    if (!have_f) f = obj.eval(x, gt);  // cost at the returned parameters (networks/calibration_network.py:71)
    __syncthreads();
    if (tid < n) p.x_out[(size_t)b * n + tid] = x[tid];
    if (tid == 0) {
        if (p.cost_out) p.cost_out[b] = f;
        if (p.converged_out) p.converged_out[b] = (f <= p.thr) ? 1 : 0;
        if (p.iters_out) p.iters_out[b] = iters;
        if (p.fevals_out) p.fevals_out[b] = fevals;
        if (p.reason_out) p.reason_out[b] = reason;
    }
}

}  // namespace davo
