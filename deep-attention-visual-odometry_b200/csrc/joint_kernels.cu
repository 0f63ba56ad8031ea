// joint_kernels.cu — kernels for the JOINT model (intrinsics + per-view pose, n = 10 + 6V <= 64): the
// CTA-per-problem solve (solver_cta.cuh) and, through the generic warp-per-problem kernels of
// wide_kernel.cuh, line search and cost + gradient (and the warp solve kept for A/B runs).
#include "davo_common.cuh"
#include "objectives_joint.cuh"
#include "solver_wide.cuh"
#include "wide_kernel.cuh"
#include "solver_cta.cuh"
#include "launch.h"

namespace davo {

template <typename T, WMode kMode>
static int launch_joint(const SolveParams<T>& p, cudaStream_t stream) {
    if (p.V < 1 || p.V > kMaxViews) return DAVO_ERR_UNSUPPORTED;
    if (p.n > 64) return launch_wide<T, JointObjective<T>, kMode, 4>(p, stream);   // 10 .. 19 views
    return launch_wide<T, JointObjective<T>, kMode>(p, stream);
}

// ---- CTA-per-problem solve (W warps cooperate on one problem) -------------------------------------------
template <typename T, int W>
__host__ __device__ inline size_t joint_cta_smem(int N, int V, int n, bool has_w) {
    return JointCtaObjective<T, W, false>::data_bytes(N, V, has_w) + CtaWorkspace<T>::bytes(n, W) + 32;
}

#ifndef DAVO_JOINT_MIN_BLOCKS
// CTAs of 2 warps per SM.  With the pose gradient accumulated as sum X' x gX' (3 sums instead of the 9 of
// sum gX' (x) X, solver_cta.cuh) the evaluator fits 96 registers without spills: config 3 takes 22.8 ms at 8 CTAs
// (122 registers), 21.6 ms at 10, 23.7 ms at 11-12; the 9-sum form took 25.0 ms at 8 and spilled at 10 (25.7 ms).
#define DAVO_JOINT_MIN_BLOCKS 10
#endif
#ifndef DAVO_JOINT_WARPS
#define DAVO_JOINT_WARPS 2       // warps per problem for V >= 2: 25.1 ms vs 27.3 ms with 4 (fewer barriers; still ~11 problems per CTA)
#endif

template <typename T, int W, bool kWeighted, int kV = 0, int kN = 0>
__global__ void __launch_bounds__(32 * W, (sizeof(T) == 4 && W > 1) ? (DAVO_JOINT_MIN_BLOCKS * 2) / W : 1) joint_cta_kernel(const SolveParams<T> p) {
    extern __shared__ __align__(128) unsigned char smem[];
    using Obj = JointCtaObjective<T, W, kWeighted, kV, kN>;
    const size_t data = Obj::data_bytes(p.N, p.V, p.has_w != 0);
    CtaWorkspace<T> ws;
    ws.carve(smem + data, p.n, W);
    unsigned char* tail = smem + data + CtaWorkspace<T>::bytes(p.n, W);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
    volatile unsigned* slot = reinterpret_cast<volatile unsigned*>(tail + 16);
    Obj obj(p, smem, ws.red, bar);
    obj.init();
    for (;;) {
        if (threadIdx.x == 0) *slot = atomicAdd(p.queue, 1u);
        __syncthreads();
        const unsigned b = *slot;
        if (b >= (unsigned)p.B) break;
        obj.bind((int)b);  // starts with a barrier: every thread has read the slot before it is rewritten
        solve_one_cta<T, W>(obj, p, (int)b, ws);
    }
}

template <typename T, int W>
static int launch_joint_cta(const SolveParams<T>& p, cudaStream_t stream) {
    auto kernel = p.has_w ? joint_cta_kernel<T, W, true> : joint_cta_kernel<T, W, false>;
    if (W == DAVO_JOINT_WARPS && !p.has_w && p.V == 4 && p.N == 256)
        kernel = joint_cta_kernel<T, W, false, 4, 256>;  // BASELINE config 3
    const size_t smem = joint_cta_smem<T, W>(p.N, p.V, p.n, p.has_w != 0);
    int dev = 0, sms = 0, max_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return DAVO_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)max_optin) return DAVO_ERR_UNSUPPORTED;
    if (!ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem))
        return DAVO_ERR_CUDA;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * W, smem) != cudaSuccess || per_sm < 1)
        return DAVO_ERR_CUDA;
    long long grid = (long long)per_sm * sms;
    if (grid > p.B) grid = p.B;
    if (grid < 1) grid = 1;
    if (cudaMemsetAsync(p.queue, 0, sizeof(unsigned), stream) != cudaSuccess) return DAVO_ERR_CUDA;
    kernel<<<(unsigned)grid, 32 * W, smem, stream>>>(p);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}

// Warp w of a problem's CTA evaluates views w, w+W, ...
template <typename T>
static int launch_joint_solve(const SolveParams<T>& p, cudaStream_t s) {
    if (p.V < 1 || p.V > kMaxViews) return DAVO_ERR_UNSUPPORTED;
    if (p.n > 64) return launch_joint<T, WMode::kSolve>(p, s);   // 10 .. 19 views: one warp per problem (wide solver)
    if (p.V >= DAVO_JOINT_WARPS) return launch_joint_cta<T, DAVO_JOINT_WARPS>(p, s);
    if (p.V >= 2) return launch_joint_cta<T, 2>(p, s);
    return launch_joint_cta<T, 1>(p, s);
}

#if DAVO_JOINT_WARP_SOLVE  // A/B: the warp-per-problem wide solver
int launch_solve_joint_f32(const SolveParams<float>& p, cudaStream_t s) { return launch_joint<float, WMode::kSolve>(p, s); }
int launch_solve_joint_f64(const SolveParams<double>& p, cudaStream_t s) { return launch_joint<double, WMode::kSolve>(p, s); }
#else
int launch_solve_joint_f32(const SolveParams<float>& p, cudaStream_t s) { return launch_joint_solve<float>(p, s); }
int launch_solve_joint_f64(const SolveParams<double>& p, cudaStream_t s) { return launch_joint_solve<double>(p, s); }
#endif
int launch_eval_joint_f32(const SolveParams<float>& p, cudaStream_t s) { return launch_joint<float, WMode::kEval>(p, s); }
int launch_eval_joint_f64(const SolveParams<double>& p, cudaStream_t s) { return launch_joint<double, WMode::kEval>(p, s); }
int launch_line_search_joint_f32(const SolveParams<float>& p, cudaStream_t s) { return launch_joint<float, WMode::kLineSearch>(p, s); }
int launch_line_search_joint_f64(const SolveParams<double>& p, cudaStream_t s) { return launch_joint<double, WMode::kLineSearch>(p, s); }

}  // namespace davo
