// solve_kernels.cu — persistent warp-per-problem kernels (solve, line search, cost+gradient).
//
// Grid: resident CTAs per SM (occupancy query) x SM count, capped by the work available.  Each warp
// pulls problem indices from an atomic counter in the caller's workspace, so a warp whose problem
// retires early immediately starts the next one; warps never wait for each other (no CTA barrier
// after set-up).
#include "davo_common.cuh"
#include "objectives.cuh"
#include "solver_warp.cuh"
#include "solver_half.cuh"
#include "solver_cta.cuh"
#include "launch.h"

namespace davo {

constexpr int kWarpsPerCta = 4;
#ifndef DAVO_MIN_BLOCKS
#define DAVO_MIN_BLOCKS 4  // 128 registers: no spills (at 5 CTAs = 96 registers the line-search state spills to local memory)
#endif

template <typename T>
__host__ __device__ constexpr size_t warp_lines_bytes() {
    // trial-point line, broadcast line, transpose scratch (kScratch words), mbarrier
    return (2 * kSlots + kScratch) * sizeof(T) + 16;
}

template <typename T, typename Obj>
__host__ __device__ inline size_t warp_smem_stride(int N, bool has_w) {
    size_t b = Obj::slab_bytes(N, has_w) + warp_lines_bytes<T>();
    return (b + 127) & ~size_t(127);
}

enum class Mode { kSolve, kLineSearch, kEval, kResolve };  // kResolve: solve the problems a first launch handed off

template <typename T, int NP, typename Obj, Mode kMode>
__global__ void __launch_bounds__(kWarpsPerCta * 32, DAVO_MIN_BLOCKS) warp_problem_kernel(const SolveParams<T> p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char* mine = smem + (size_t)warp * warp_smem_stride<T, Obj>(p.N, p.has_w != 0);
    const size_t slab = Obj::slab_bytes(p.N, p.has_w != 0);
    T* xt_line = reinterpret_cast<T*>(mine + slab);
    T* bc_line = xt_line + kSlots;
    T* scratch = bc_line + kSlots;
    uint64_t* bar = reinterpret_cast<uint64_t*>(scratch + kScratch);
    if (Obj::kUsesSmemMatches) {
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncwarp();
    }
    Obj obj(p, mine, scratch, bar, lane);
    const int n = p.n;
    const int c = lane >> 1;
    const bool own = c < n;
    const unsigned total_warps = gridDim.x * kWarpsPerCta;
    unsigned b_static = blockIdx.x * kWarpsPerCta + warp;
    if (kMode == Mode::kResolve) {
        // Second launch of a DISTORT10 solve: warps pull 32 problem indices at a time, look at their reason flags
        // and solve the ones the first launch handed off (normally none: one coalesced read per 32 problems).
        for (;;) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(p.queue, 32u);
            base = __shfl_sync(kFull, base, 0);
            if (base >= (unsigned)p.B) break;
            const unsigned idx = base + lane;
            unsigned todo = __ballot_sync(kFull, idx < (unsigned)p.B && p.reason_out[idx] == kReasonHandoff);
            while (todo) {
                const int j = __ffs(todo) - 1;
                todo &= todo - 1;
                obj.bind((int)(base + j));
                solve_one_warp<T, NP, Obj>(obj, p, (int)(base + j), xt_line, bc_line, scratch, lane);
            }
        }
        return;
    }
    for (;;) {
        unsigned b = 0;
        if (kMode == Mode::kSolve) {  // iteration counts vary 10x between problems: dynamic queue
            if (lane == 0) b = atomicAdd(p.queue, 1u);
            b = __shfl_sync(kFull, b, 0);
        } else {                      // uniform work per problem: static round-robin
            b = b_static;
            b_static += total_warps;
        }
        if (b >= (unsigned)p.B) break;
        obj.bind((int)b);
        if (kMode == Mode::kSolve) {
            solve_one_warp<T, NP, Obj>(obj, p, (int)b, xt_line, bc_line, scratch, lane);
        } else if (kMode == Mode::kLineSearch) {
            const size_t o = (size_t)b * n + c;
            const T x = own ? p.x0[o] : T(0);
            const T d = own ? p.dir[o] : T(0);
            const T g = own ? p.base_grad[o] : T(0);
            const LineSearchResult<T> r = line_search_warp(obj, p, x, d, p.base_cost[b], g, xt_line, lane);
            if (lane == 0) {
                p.alpha_out[b] = r.alpha;
                if (p.fevals_out) p.fevals_out[b] = r.probes;
            }
        } else {
            const T x = own ? p.x0[(size_t)b * n + c] : T(0);
            T f, g;
            eval_at(obj, x, xt_line, lane, f, g);
            if (lane == 0 && p.cost_out) p.cost_out[b] = f;
            if (p.x_out && own && !(lane & 1)) p.x_out[(size_t)b * n + c] = g;  // x_out doubles as grad out
        }
    }
}

template <typename T, int NP, typename Obj, Mode kMode>
static int launch_warp_kernel(const SolveParams<T>& p, cudaStream_t stream) {
    auto kernel = warp_problem_kernel<T, NP, Obj, kMode>;
    const size_t smem = warp_smem_stride<T, Obj>(p.N, p.has_w != 0) * kWarpsPerCta;
    int dev = 0, sms = 0, max_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return DAVO_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)max_optin) return DAVO_ERR_UNSUPPORTED;  // N too large for a warp-resident slab
    if (!ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem))
        return DAVO_ERR_CUDA;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarpsPerCta * 32, smem) != cudaSuccess ||
        per_sm < 1)
        return DAVO_ERR_CUDA;
    long long grid = (long long)per_sm * sms;
    const long long need = (kMode == Mode::kResolve) ? ((long long)p.B + 32 * kWarpsPerCta - 1) / (32 * kWarpsPerCta)
                                                      : ((long long)p.B + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (kMode == Mode::kSolve && cudaMemsetAsync(p.queue, 0, sizeof(unsigned), stream) != cudaSuccess)
        return DAVO_ERR_CUDA;  // (kResolve: the first launch has zeroed both counters)
    kernel<<<(unsigned)grid, kWarpsPerCta * 32, smem, stream>>>(p);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}

#ifndef DAVO_HALF_SOLVE
#define DAVO_HALF_SOLVE 1  // 0: A/B builds solve DISTORT10 with one warp per problem
#endif

template <typename T>
static int launch_resolve_cta(const SolveParams<T>& p, cudaStream_t stream);

#ifndef DAVO_RESOLVE_WARP
#define DAVO_RESOLVE_WARP 0  // 1: A/B builds re-solve the stragglers one warp per problem
#endif
#ifndef DAVO_EVAL_CAP_PER_ITER
#define DAVO_EVAL_CAP_PER_ITER 4  // ... or this many evaluations per allowed outer iteration, whichever is larger
#endif
#ifndef DAVO_EVAL_CAP
#define DAVO_EVAL_CAP 4096  // reference-equivalent evaluations after which the two-per-warp launch hands a problem off
#endif

// DISTORT10 solve, unweighted: two problems per warp (solver_half.cuh), then one warp per problem for the
// stragglers the first launch handed off.
template <typename T>
static int launch_half_kernel(const SolveParams<T>& p_in, cudaStream_t stream) {
    SolveParams<T> p = p_in;
    // The hand-off flag lives in reason_out.  The cap grows with the iteration limit (a problem may legitimately use
    // ~2.5 evaluations per outer iteration): only line-search stragglers pass 4 per iteration on average.
#ifdef DAVO_EVAL_CAP_EXACT  // test builds (tools/resolve_check.py): hand nearly every problem to the second launch
    const long long cap = DAVO_EVAL_CAP;
#else
    const long long per_iter = (long long)DAVO_EVAL_CAP_PER_ITER * p.max_iters;
    const long long cap = per_iter > DAVO_EVAL_CAP ? per_iter : DAVO_EVAL_CAP;
#endif
    p.eval_cap = p.reason_out ? (int)(cap > 0x7fffffffLL ? 0x7fffffffLL : cap) : 0;
    auto kernel = (p.N == 256)     ? half_problem_kernel<T, false, 256>   // BASELINE configs 2, 4, 5
                  : (p.N % 32 == 0) ? half_problem_kernel<T, false>
                                    : half_problem_kernel<T, true>;
    const size_t smem = half_stride<T>(p.N) * 2 * kWarpsPerCta;
    int dev = 0, sms = 0, max_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return DAVO_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)max_optin) return DAVO_ERR_UNSUPPORTED;
    if (!ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem))
        return DAVO_ERR_CUDA;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarpsPerCta * 32, smem) != cudaSuccess ||
        per_sm < 1)
        return DAVO_ERR_CUDA;
    long long grid = (long long)per_sm * sms;
    const long long need = ((long long)p.B + 2 * kWarpsPerCta - 1) / (2 * kWarpsPerCta);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    // fewer than ~3 problems per resident half: the batch cannot keep two problems per warp busy and a problem's
    // latency matters more than the instruction count (the streamed host-input path solves 4K-problem chunks)
    if ((long long)p.B < 3 * 2 * kWarpsPerCta * (long long)per_sm * sms) return DAVO_ERR_UNSUPPORTED;
    if (cudaMemsetAsync(p.queue, 0, DAVO_WORKSPACE_BYTES, stream) != cudaSuccess) return DAVO_ERR_CUDA;
    kernel<<<(unsigned)grid, kWarpsPerCta * 32, smem, stream>>>(p);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return DAVO_ERR_CUDA;
    if (p.eval_cap == 0) return DAVO_OK;
    SolveParams<T> p2 = p_in;
    p2.queue = p.queue;
    p2.eval_cap = (int)(grid * kWarpsPerCta);   // warps of the first launch: the second waits for that many exits
#if DAVO_RESOLVE_WARP
    p2.queue = p.queue + kWsScan;
    return launch_warp_kernel<T, 10, Distort10Objective<T, false>, Mode::kResolve>(p2, stream);
#else
    return launch_resolve_cta<T>(p2, stream);
#endif
}

// ---- second launch of a DISTORT10 solve: one CTA (4 warps) per handed-off problem ------------------------------
#ifndef DAVO_RESOLVE_WARPS
// Warps that share one evaluation; the CTA is DAVO_SPEC_PROBES groups of them, each evaluating DAVO_SPEC_PER_GROUP
// trial points per line-search round (solver_cta.cuh).  Measured on config 4 (64K problems; 13.7 ms of it is the first
// launch), whole solve, warps x groups x points per group: 4 x 1 x 1 (no speculation) 43.4 ms, 4 x 4 x 1 32.8,
// 2 x 8 x 1 28.7, 1 x 16 x 1 27.6-28.5, 1 x 16 x 2 27.7, 1 x 8 x 2 28.8, 1 x 8 x 4 27.0 (256 threads), 1 x 24 x 1 30.0,
// 4 x 8 x 1 (1024 threads, spills) 49.4.  The longest straggler alone (47.8 K probes in ~45-probe D chains): 16.0 ms
// of re-solve at 4 x 4 x 1, ~10 ms from 16 trial points per round up (tools/straggler_probe.py).
#define DAVO_RESOLVE_WARPS 1
#endif
#ifndef DAVO_RESOLVE_CTAS
#define DAVO_RESOLVE_CTAS 16   // CTAs of the early (programmatic dependent) straggler launch
#endif
constexpr int kResolveWarps = DAVO_RESOLVE_WARPS;   // warps that share one evaluation
constexpr int kResolveGroups = DAVO_SPEC_PROBES;    // groups of kResolveWarps warps: trial points per line-search round
constexpr int kResolveCtaWarps = kResolveWarps * kResolveGroups;

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }

// The straggler launches of a DISTORT10 solve.
// kEarly = true: a PROGRAMMATIC DEPENDENT launch of a few CTAs (DAVO_RESOLVE_CTAS).  They become resident as the first
// launch's warps run out of problems (its tail); CTA t waits (sleeping) for hand-off t of the workspace's list and
// solves it while the first launch is still finishing its last problems — the longest stragglers are handed off long
// before that — or leaves when every warp of the first launch has exited without publishing a t-th hand-off.
// kEarly = false: the plain stream-ordered launch behind it, a full grid that scans every flag and solves whatever is
// still marked (hand-offs beyond the early CTAs): normally nothing, 17 us.
// Only DAVO_RESOLVE_CTAS early CTAs because each holds three of an SM's four first-launch slots while it waits: with one
// per SM the next batch's first launch (bench.py keeps two batches in flight) could not start in the slots the tail
// frees (config 2: 3.91 ms per batch against 3.67 with 16).  Which CTA solves a problem does not matter: same route,
// same bits (tools/cfg4_digest.py prints the same digest as the single stream-ordered launch).
template <typename T, bool kEarly>
__global__ void __launch_bounds__(32 * kResolveCtaWarps, 1) resolve_cta_kernel(const SolveParams<T> p) {
    extern __shared__ __align__(128) unsigned char smem[];
    using Obj = Distort10CtaObjective<T, kResolveWarps, kResolveGroups>;
#if DAVO_TIMELINE
    if (threadIdx.x == 0 && (kEarly || blockIdx.x == 0)) printf("T %s cta %u resident at %llu\n", kEarly ? "early" : "scan", blockIdx.x, global_ns());
#endif
    const size_t data = Obj::data_bytes(p.N, 1, false);
    CtaWorkspace<T> ws;
    ws.carve(smem + data, p.n, kResolveCtaWarps);
    unsigned char* tail = smem + data + CtaWorkspace<T>::bytes(p.n, kResolveCtaWarps);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
    volatile unsigned* slot = reinterpret_cast<volatile unsigned*>(tail + 16);  // [0] problem / chunk base, [1] flag mask
    Obj obj(p, smem, ws.red, bar);
    obj.init();
    const int lane = threadIdx.x & 31;
    int* flags = p.reason_out;
    if (kEarly) {
        const unsigned first_warps = (unsigned)p.eval_cap;
        constexpr unsigned kNone = 0xffffffffu;
        constexpr unsigned kWatchdog = 1u << 22;   // polls of ~0.3 us: about a second, then leave it to the scan
        // CTA t takes list entries t, t + gridDim.x, ... in turn
        for (unsigned t = blockIdx.x; t < (unsigned)kHandoffList; t += gridDim.x) {
            if (threadIdx.x == 0) {
                unsigned prob = kNone;
                for (unsigned spins = 0; spins < kWatchdog; ++spins) {
                    unsigned v = ld_volatile_u32(p.queue + kWsList + t);
                    if (v) { prob = v - 1u; break; }
                    if (ld_volatile_u32(p.queue + kWsExited) >= first_warps) {   // every hand-off has been published
                        __threadfence();
                        v = ld_volatile_u32(p.queue + kWsList + t);
                        if (v) prob = v - 1u;
                        break;
                    }
                    __nanosleep(256);
                }
                slot[0] = prob;
            }
            __syncthreads();
            const unsigned prob = slot[0];
            __syncthreads();  // everyone has read the slot before thread 0 may rewrite it
            if (prob == kNone) break;   // no t-th hand-off: there is no later one either
#if DAVO_TIMELINE
            if (threadIdx.x == 0) printf("T early cta %u entry %u problem %u starts at %llu\n", blockIdx.x, t, prob, global_ns());
#endif
            obj.bind((int)prob);
            solve_one_cta<T, kResolveCtaWarps>(obj, p, (int)prob, ws);   // rewrites the flag with the final reason
#if DAVO_TIMELINE
            if (threadIdx.x == 0) printf("T early cta %u entry %u problem %u done at %llu\n", blockIdx.x, t, prob, global_ns());
#endif
        }
        // this grid must not complete before the first launch has (the scan launch behind it is ordered after THIS
        // grid only): block until the first launch has finished and flushed its results
        asm volatile("griddepcontrol.wait;" ::: "memory");
        return;
    }
    for (;;) {
        // warp 0 pulls 32 problem indices and looks at their reason flags (normally none is set)
        if (threadIdx.x < 32) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(p.queue + kWsScan, 32u);
            base = __shfl_sync(kFull, base, 0);
            const unsigned idx = base + lane;
            const unsigned todo = __ballot_sync(kFull, idx < (unsigned)p.B && flags[idx] == kReasonHandoff);
            if (lane == 0) {
                slot[0] = base;
                slot[1] = todo;
            }
        }
        __syncthreads();
        const unsigned base = slot[0];
        unsigned todo = slot[1];
        __syncthreads();  // everyone has read the slots before warp 0 may rewrite them
        if (base >= (unsigned)p.B) break;
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            obj.bind((int)(base + j));
            solve_one_cta<T, kResolveCtaWarps>(obj, p, (int)(base + j), ws);
        }
    }
}

template <typename T>
static int launch_resolve_cta(const SolveParams<T>& p, cudaStream_t stream) {
    using Obj = Distort10CtaObjective<T, kResolveWarps, kResolveGroups>;
    auto early = resolve_cta_kernel<T, true>;
    auto scan = resolve_cta_kernel<T, false>;
    const size_t smem = Obj::data_bytes(p.N, 1, false) + CtaWorkspace<T>::bytes(p.n, kResolveCtaWarps) + 32;
    int dev = 0, sms = 0, max_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return DAVO_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)max_optin) return DAVO_ERR_UNSUPPORTED;
    if (!ensure_dynamic_smem(reinterpret_cast<const void*>(early), smem) ||
        !ensure_dynamic_smem(reinterpret_cast<const void*>(scan), smem))
        return DAVO_ERR_CUDA;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan, 32 * kResolveCtaWarps, smem) != cudaSuccess ||
        per_sm < 1)
        return DAVO_ERR_CUDA;
    long long grid = (long long)per_sm * sms;
    const long long need = ((long long)p.B + 31) / 32;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    // the early CTAs: programmatic dependent launch (the first launch calls griddepcontrol.launch_dependents at its top);
    // they synchronise with it through the workspace counters, not through griddepcontrol.wait
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(DAVO_RESOLVE_CTAS < kHandoffList ? DAVO_RESOLVE_CTAS : kHandoffList);
    cfg.blockDim = dim3(32 * kResolveCtaWarps);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, early, p) == cudaSuccess) count_launch();
    else (void)cudaGetLastError();   // no early start: the scan below solves every hand-off
    scan<<<(unsigned)grid, 32 * kResolveCtaWarps, smem, stream>>>(p);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}

template <typename T, Mode kMode>
static int dispatch_model(const SolveParams<T>& p, cudaStream_t stream) {
#if DAVO_HALF_SOLVE
    if (kMode == Mode::kSolve && p.model == DAVO_MODEL_DISTORT10 && !p.has_w && p.n == 10) {
        const int st = launch_half_kernel<T>(p, stream);
        if (st != DAVO_ERR_UNSUPPORTED) return st;  // N too large for two slabs per warp: one warp per problem
    }
#endif
    if (p.model == DAVO_MODEL_DISTORT10)
        return p.has_w ? launch_warp_kernel<T, 10, Distort10Objective<T, true>, kMode>(p, stream)
                       : launch_warp_kernel<T, 10, Distort10Objective<T, false>, kMode>(p, stream);
    if (p.model >= DAVO_MODEL_SPHERE && p.model <= DAVO_MODEL_DISTANCE) {
        if (p.n > kSlots) return DAVO_ERR_UNSUPPORTED;
        return launch_warp_kernel<T, kSlots, AnalyticObjective<T>, kMode>(p, stream);
    }
    return DAVO_ERR_UNSUPPORTED;
}

int launch_solve_warp_f32(const SolveParams<float>& p, cudaStream_t s) { return dispatch_model<float, Mode::kSolve>(p, s); }
int launch_solve_warp_f64(const SolveParams<double>& p, cudaStream_t s) { return dispatch_model<double, Mode::kSolve>(p, s); }
int launch_line_search_warp_f32(const SolveParams<float>& p, cudaStream_t s) { return dispatch_model<float, Mode::kLineSearch>(p, s); }
int launch_line_search_warp_f64(const SolveParams<double>& p, cudaStream_t s) { return dispatch_model<double, Mode::kLineSearch>(p, s); }
int launch_eval_warp_f32(const SolveParams<float>& p, cudaStream_t s) { return dispatch_model<float, Mode::kEval>(p, s); }
int launch_eval_warp_f64(const SolveParams<double>& p, cudaStream_t s) { return dispatch_model<double, Mode::kEval>(p, s); }

}  // namespace davo
