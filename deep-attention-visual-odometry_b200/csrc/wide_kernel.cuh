// wide_kernel.cuh — persistent warp-per-problem kernels around the wide solver (solver_wide.cuh, n <= 64) for any
// objective with the interface
//   Obj(const SolveParams<T>&, unsigned char* slab, int lane);  init();  bind(problem);
//   T eval(const T* theta_smem, T* grad_smem);                  static size_t slab_bytes(N, V, has_w);
// Modes: solve (atomic work queue: iteration counts vary between problems), line search and cost + gradient
// (static round-robin).  Each warp owns a slab of shared memory: the objective's staged data, the solver's
// n-vectors and the n x n inverse Hessian.
#pragma once
#include "davo_common.cuh"
#include "solver_wide.cuh"
#include "launch.h"

namespace davo {

constexpr int kWideWarpsPerCta = 2;
enum class WMode { kSolve, kLineSearch, kEval };

template <typename T, typename Obj>
__host__ __device__ inline size_t wide_warp_stride(int N, int V, int n, bool has_w) {
    size_t b = Obj::slab_bytes(N, V, has_w) + WideWorkspace<T>::bytes(n);
    return (b + 127) & ~size_t(127);
}

// kCols: components per lane (2 for n <= 64, 4 for n <= 128)
template <typename T, typename Obj, WMode kMode, int kCols = 2>
__global__ void __launch_bounds__(kWideWarpsPerCta * 32) wide_problem_kernel(const SolveParams<T> p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = smem + (size_t)warp * wide_warp_stride<T, Obj>(p.N, p.V, p.n, p.has_w != 0);
    Obj obj(p, mine, lane);
    obj.init();
    WideWorkspace<T> ws;
    ws.carve(mine + Obj::slab_bytes(p.N, p.V, p.has_w != 0), p.n);
    // the n-vectors' tails and the padding columns of H are read by the vectorised sweeps (compile-time n): zero once
    for (size_t c = lane; c < WideWorkspace<T>::bytes(p.n) / sizeof(T); c += 32) ws.x[c] = T(0);
    __syncwarp();
    const int n = p.n;
    const unsigned warps_per_cta = blockDim.x >> 5;  // 2, or 1 when two slabs do not fit in shared memory
    const unsigned total_warps = gridDim.x * warps_per_cta;
    unsigned b_static = blockIdx.x * warps_per_cta + warp;
    for (;;) {
        unsigned b = 0;
        if (kMode == WMode::kSolve) {
            if (lane == 0) b = atomicAdd(p.queue, 1u);
            b = __shfl_sync(kFull, b, 0);
        } else {
            b = b_static;
            b_static += total_warps;
        }
        if (b >= (unsigned)p.B) break;
        obj.bind((int)b);
        if (kMode == WMode::kSolve) {
            solve_one_wide<kCols>(obj, p, (int)b, ws, lane);
        } else {
            for (int c = lane; c < n; c += 32) ws.x[c] = p.x0[(size_t)b * n + c];
            __syncwarp();
            if (kMode == WMode::kLineSearch) {
                for (int c = lane; c < n; c += 32) {
                    ws.d[c] = p.dir[(size_t)b * n + c];
                    ws.g[c] = p.base_grad[(size_t)b * n + c];
                }
                __syncwarp();
                const LineSearchResult<T> r =
                    line_search_wide<kCols>(obj, p, ws.x, ws.d, p.base_cost[b], ws.g, ws.xt, ws.gt, lane);
                if (lane == 0) {
                    p.alpha_out[b] = r.alpha;
                    if (p.fevals_out) p.fevals_out[b] = r.probes;
                }
            } else {
                const T f = obj.eval(ws.x, ws.g);
                if (lane == 0 && p.cost_out) p.cost_out[b] = f;
                if (p.x_out)
                    for (int c = lane; c < n; c += 32) p.x_out[(size_t)b * n + c] = ws.g[c];
            }
            __syncwarp();
        }
    }
}

template <typename T, typename Obj, WMode kMode, int kCols = 2>
static int launch_wide(const SolveParams<T>& p, cudaStream_t stream) {
    if (p.n > 32 * kCols || p.n > kWideMax) return DAVO_ERR_UNSUPPORTED;
    auto kernel = wide_problem_kernel<T, Obj, kMode, kCols>;
    int dev = 0, sms = 0, max_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return DAVO_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const size_t stride = wide_warp_stride<T, Obj>(p.N, p.V, p.n, p.has_w != 0);
    int warps = kWideWarpsPerCta;
    if (stride * warps > (size_t)max_optin) warps = 1;  // e.g. n = 111 in float64: one 120 KB slab per CTA
    const size_t smem = stride * warps;
    if (smem > (size_t)max_optin) return DAVO_ERR_UNSUPPORTED;
    if (!ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem))
        return DAVO_ERR_CUDA;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem) != cudaSuccess || per_sm < 1)
        return DAVO_ERR_CUDA;
    long long grid = (long long)per_sm * sms;
    const long long need = ((long long)p.B + warps - 1) / warps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (kMode == WMode::kSolve && cudaMemsetAsync(p.queue, 0, sizeof(unsigned), stream) != cudaSuccess)
        return DAVO_ERR_CUDA;
    kernel<<<(unsigned)grid, warps * 32, smem, stream>>>(p);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}

}  // namespace davo
