// davo_common.cuh — shared device helpers for the sm_100a calibration-solve kernels.
//
// Vocabulary: a PROBLEM is one independent calibration fit; a MATCH is one 3-D <-> 2-D
// correspondence; a SLOT is one of 16 positions of the warp-distributed parameter vector
// (slot c is held, duplicated, by lanes 2c and 2c+1).
#pragma once
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/davo_b200.h"

#ifndef DAVO_TRACE
#define DAVO_TRACE 0
#endif

namespace davo {

constexpr int kReasonHandoff = 4;  // internal value of reason_out between the two launches of a DISTORT10 solve
constexpr int kReasonClaimed = 6;  // internal: a CTA of the second launch has taken the problem
// Words of the caller's 256-byte workspace used by the two launches of a DISTORT10 solve (all zeroed before the first):
//   [0] problem queue of the first launch      [1] flag-scan position of the second launch
//   [2] hand-offs reserved by the first launch  [3] hand-off tickets taken by the second launch
//   [4] warps of the first launch that have exited
//   [5 ..] the hand-off list: problem index + 1 (0 = not published yet)
#ifndef DAVO_TIMELINE
#define DAVO_TIMELINE 0   // 1: debug builds print hand-off / straggler timestamps (tools/straggler_timeline.py)
#endif
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
constexpr int kWsQueue = 0, kWsScan = 1, kWsReserved = 2, kWsTaken = 3, kWsExited = 4, kWsList = 5;
constexpr int kHandoffList = DAVO_WORKSPACE_BYTES / 4 - kWsList;
constexpr int kSlots = 16;  // distributed-vector width of the warp-per-problem solver (n <= 16)
constexpr unsigned kFull = 0xffffffffu;
// Per-warp shared-memory scratch of the warp-per-problem solver: 32 rows x kRedPitch words.  Cross-lane sums go
// through it as a transpose (every lane stores its partial sums as a row, lane pair c adds up column c) instead
// of a shuffle reduce-scatter: the shuffle form costs two selects per exchanged value, and the kernel is bound
// by instruction issue.  Pitch 20 words: rows stay 16-byte aligned and 8 consecutive rows start in distinct
// 4-bank groups (STS.128 conflict free); a column read hits bank (20 r + c) mod 32.
constexpr int kRedPitch = 20;
constexpr int kScratch = 32 * kRedPitch;

// ---- vector-of-4 type per arithmetic type -----------------------------------------------------
template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };
template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

// ---- non-contracted arithmetic ----------------------------------------------------------------
// The control arithmetic of the solver (trial points, Wolfe tests, the H update) mirrors the
// reference's op-by-op rounding, so these must never be fused into FMAs by nvcc.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float rcp_rn(float a) { return __frcp_rn(a); }
__device__ __forceinline__ double rcp_rn(double a) { return __drcp_rn(a); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }

// interpolate_alpha (utils/func_interpolate_alpha.py:15-33): the zero of the line through (a1, v1), (a2, v2); the
// midpoint when the values are equal or the zero is within 1e-3 of (or outside) the interval.  Operation by operation.
template <typename T>
__device__ __forceinline__ T interpolate_alpha_value(T a1, T a2, T v1, T v2, bool* nonlinear = nullptr) {
    const T lo = a1 < a2 ? a1 : a2, hi = a1 < a2 ? a2 : a1;          // torch.minimum / maximum (:15-16)
    const T diff = sub_rn(v2, v1);                                          // :18
    const T inv_gradient = div_rn(sub_rn(a2, a1), diff);                         // :19
    T cand = sub_rn(a1, mul_rn(v1, inv_gradient));   // two roundings, as torch's mul then sub                                 // :20
    const bool nl = (diff == T(0)) || (cand < add_rn(lo, T(1e-3))) || (cand > sub_rn(hi, T(1e-3)));   // :23-29
    if (nl) cand = div_rn(add_rn(a1, a2), T(2));                                  // :30-32
    if (nonlinear) *nonlinear = nl;
    return cand;
}


// ---- warp shuffles for float / double ---------------------------------------------------------
template <typename T>
__device__ __forceinline__ T shfl_xor(T v, int mask) { return __shfl_xor_sync(kFull, v, mask); }
template <typename T>
__device__ __forceinline__ T shfl_idx(T v, int src) { return __shfl_sync(kFull, v, src); }

// Sum over the 16 slots of a slot-distributed value (lanes 2c and 2c+1 hold the same input):
// butterfly over lane bits 1..4.  Every lane ends with the bitwise-identical total.
template <typename T>
__device__ __forceinline__ T slot_allreduce(T v) {
    v += shfl_xor(v, 2);
    v += shfl_xor(v, 4);
    v += shfl_xor(v, 8);
    v += shfl_xor(v, 16);
    return v;
}

// Reduce-scatter of 16 per-lane values.  kDistinctLanes = true: all 32 lanes hold different partial
// sums (a reduction over matches) -> 16 shuffles; false: lanes 2c/2c+1 hold the same 16 values
// (a reduction over slots, i.e. over rows of H) -> 15 shuffles.  Afterwards lane L holds the total
// of value index L >> 1.
template <bool kDistinctLanes, typename T>
__device__ __forceinline__ T reduce_scatter16(T (&v)[kSlots], int lane) {
    {
        const bool up = lane & 16;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            T send = up ? v[k] : v[k + 8];
            T keep = up ? v[k + 8] : v[k];
            v[k] = keep + shfl_xor(send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            T send = up ? v[k] : v[k + 4];
            T keep = up ? v[k + 4] : v[k];
            v[k] = keep + shfl_xor(send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            T send = up ? v[k] : v[k + 2];
            T keep = up ? v[k + 2] : v[k];
            v[k] = keep + shfl_xor(send, 4);
        }
    }
    {
        const bool up = lane & 2;
        T send = up ? v[0] : v[1];
        T keep = up ? v[1] : v[0];
        v[0] = keep + shfl_xor(send, 2);
    }
    if (kDistinctLanes) v[0] += shfl_xor(v[0], 1);
    return v[0];
}

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk, SASS UBLKCP) --------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; bytes must be a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- device-side copy of the descriptor, typed --------------------------------------------------
template <typename T>
struct SolveParams {
    int B, N, V, n, model, max_iters, max_ls, strong, has_w;
    int zoom;  // 1: secant zoom (davo_problem_desc.zoom_interpolation), generic solver only
    T c1, c2, thr, min_step;
    const T* data0;
    const T* data1;
    const T* w;
    const T* x0;
    T* x_out;
    T* cost_out;
    uint8_t* converged_out;
    int32_t* iters_out;
    int32_t* fevals_out;
    int32_t* reason_out;
    unsigned* queue;  // atomic work-queue counter in the caller's workspace
    // Straggler hand-off (solver_half.cuh): a problem whose reference-equivalent evaluation count passes eval_cap
    // at the top of an outer iteration is abandoned with reason_out = kReasonHandoff, and a second launch
    // (Mode::kResolve, one warp per problem) solves exactly those problems again from x0.  0 = no cap.
    int eval_cap;
    // line-search-only entry point
    const T* dir;
    const T* base_cost;
    const T* base_grad;
    T* alpha_out;
#if DAVO_TRACE
    T* trace;           // debug builds only: 8 values per accepted step of problem `trace_problem`
    int trace_problem;
    int trace_capacity;
#endif
};

}  // namespace davo
