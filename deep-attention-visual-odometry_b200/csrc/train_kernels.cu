// train_kernels.cu — the training-mode / differentiable solve for every objective model:
//   forward:  BFGSSolver.forward with self.training or requires_grad inputs (autograd_solvers/bfgs_solver.py:80-215:
//             training thresholds :88-93, drop-path :122-125, return_second_last :196-212), one warp per problem on
//             the generic solver (solver_wide.cuh) with a TrainRecorder that keeps the accepted iterates;
//   backward: d loss / d x0 from d loss / d x_out through those iterates (solver_train.cuh), float64.
#include "davo_common.cuh"
#include "objectives.cuh"
#include "objectives_wide.cuh"
#include "objectives_joint.cuh"
#include "objectives_ba.cuh"
#include "solver_wide.cuh"
#include "solver_train.cuh"
#include "wide_kernel.cuh"
#include "launch.h"

namespace davo {

constexpr int kTrainWarpsPerCta = 2;

template <typename T, typename Obj>
__host__ __device__ inline size_t train_fwd_stride(int N, int V, int n, bool has_w) {
    size_t b = Obj::slab_bytes(N, V, has_w) + WideWorkspace<T>::bytes(n);
    return (b + 127) & ~size_t(127);
}
template <typename T, typename Obj>
__host__ __device__ inline size_t train_bwd_stride(int N, int V, int n, bool has_w) {
    size_t b = Obj::slab_bytes(N, V, has_w) + BackwardWorkspace<T>::bytes(n);
    return (b + 127) & ~size_t(127);
}

template <typename T, typename Obj, int kCols>
__global__ void __launch_bounds__(kTrainWarpsPerCta * 32) train_forward_kernel(const SolveParams<T> p,
                                                                               const TrainRecorder<T> rec) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = smem + (size_t)warp * train_fwd_stride<T, Obj>(p.N, p.V, p.n, p.has_w != 0);
    Obj obj(p, mine, lane);
    obj.init();
    WideWorkspace<T> ws;
    ws.carve(mine + Obj::slab_bytes(p.N, p.V, p.has_w != 0), p.n);
    for (size_t c = lane; c < WideWorkspace<T>::bytes(p.n) / sizeof(T); c += 32) ws.x[c] = T(0);  // vectors and H are contiguous from ws.x
    __syncwarp();
    for (;;) {
        unsigned b = 0;
        if (lane == 0) b = atomicAdd(p.queue, 1u);
        b = __shfl_sync(kFull, b, 0);
        if (b >= (unsigned)p.B) break;
        obj.bind((int)b);
        solve_one_wide<kCols>(obj, p, (int)b, ws, lane, rec);
    }
}

template <typename T, typename Obj>
__global__ void __launch_bounds__(kTrainWarpsPerCta * 32) train_backward_kernel(const SolveParams<T> p,
                                                                                const BackwardParams<T> bp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = smem + (size_t)warp * train_bwd_stride<T, Obj>(p.N, p.V, p.n, p.has_w != 0);
    Obj obj(p, mine, lane);
    obj.init();
    BackwardWorkspace<T> ws;
    ws.carve(mine + Obj::slab_bytes(p.N, p.V, p.has_w != 0), p.n);
    for (int c = lane; c < BackwardWorkspace<T>::kVecs * wide_vec(p.n); c += 32) ws.vec[0][c] = T(0);
    __syncwarp();
    for (;;) {
        unsigned b = 0;
        if (lane == 0) b = atomicAdd(p.queue, 1u);
        b = __shfl_sync(kFull, b, 0);
        if (b >= (unsigned)p.B) break;
        obj.bind((int)b);
        backward_one(obj, p, bp, (int)b, ws, lane);
    }
}

// grid / shared memory for a persistent warp-per-problem launch with `stride` bytes per warp
template <typename Kernel>
static int plan_launch(Kernel kernel, size_t stride, int B, int& warps, size_t& smem, long long& grid) {
    int dev = 0, sms = 0, max_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return DAVO_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    warps = kTrainWarpsPerCta;
    if (stride * warps > (size_t)max_optin) warps = 1;
    smem = stride * warps;
    if (smem > (size_t)max_optin) return DAVO_ERR_UNSUPPORTED;
    if (!ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem)) return DAVO_ERR_CUDA;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem) != cudaSuccess || per_sm < 1)
        return DAVO_ERR_CUDA;
    grid = (long long)per_sm * sms;
    const long long need = ((long long)B + warps - 1) / warps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    return DAVO_OK;
}

template <typename T, typename Obj, int kCols>
static int launch_train_forward_t(const SolveParams<T>& p, const TrainRecorder<T>& rec, cudaStream_t stream) {
    if (p.n > 32 * kCols || p.n > kWideMax) return DAVO_ERR_UNSUPPORTED;
    auto kernel = train_forward_kernel<T, Obj, kCols>;
    int warps = 0;
    size_t smem = 0;
    long long grid = 0;
    const int st = plan_launch(kernel, train_fwd_stride<T, Obj>(p.N, p.V, p.n, p.has_w != 0), p.B, warps, smem, grid);
    if (st) return st;
    if (cudaMemsetAsync(p.queue, 0, sizeof(unsigned), stream) != cudaSuccess) return DAVO_ERR_CUDA;
    kernel<<<(unsigned)grid, warps * 32, smem, stream>>>(p, rec);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}

template <typename T, typename Obj>
static int launch_train_backward_t(const SolveParams<T>& p, const BackwardParams<T>& bp, cudaStream_t stream) {
    if (p.n > kWideMax) return DAVO_ERR_UNSUPPORTED;
    if (bp.grad_data && !DataGradient<Obj, T>::kSupported) return DAVO_ERR_UNSUPPORTED;
    auto kernel = train_backward_kernel<T, Obj>;
    int warps = 0;
    size_t smem = 0;
    long long grid = 0;
    const int st = plan_launch(kernel, train_bwd_stride<T, Obj>(p.N, p.V, p.n, p.has_w != 0), p.B, warps, smem, grid);
    if (st) return st;
    if (cudaMemsetAsync(p.queue, 0, sizeof(unsigned), stream) != cudaSuccess) return DAVO_ERR_CUDA;
    kernel<<<(unsigned)grid, warps * 32, smem, stream>>>(p, bp);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}

static bool analytic_model(int model) { return model >= DAVO_MODEL_SPHERE && model <= DAVO_MODEL_DISTANCE; }

template <typename T>
int launch_train_forward(const SolveParams<T>& p, const TrainRecorder<T>& rec, cudaStream_t s) {
    if (p.model == DAVO_MODEL_DISTORT10) {
        const int st = launch_train_forward_t<T, Distort10WideObjective<T>, 2>(p, rec, s);
        if (st != DAVO_ERR_UNSUPPORTED) return st;  // N too large for a shared-memory slab: read the matches from global
        return launch_train_forward_t<T, Distort10WideObjective<T, true>, 2>(p, rec, s);
    }
    if (p.model == DAVO_MODEL_JOINT) {
        if (p.V < 1 || p.V > kMaxViews) return DAVO_ERR_UNSUPPORTED;
        if (p.n > 64) return launch_train_forward_t<T, JointObjective<T>, 4>(p, rec, s);
        return launch_train_forward_t<T, JointObjective<T>, 2>(p, rec, s);
    }
    if (p.model == DAVO_MODEL_ANGLE_BA) {
        if (p.n > 64) return launch_train_forward_t<T, AngleBAObjective<T>, 4>(p, rec, s);
        return launch_train_forward_t<T, AngleBAObjective<T>, 2>(p, rec, s);
    }
    if (analytic_model(p.model)) return launch_train_forward_t<T, AnalyticWideObjective<T>, 2>(p, rec, s);
    return DAVO_ERR_UNSUPPORTED;
}

int launch_train_backward(const SolveParams<double>& p, const BackwardParams<double>& bp, cudaStream_t s) {
    if (p.model == DAVO_MODEL_DISTORT10) return launch_train_backward_t<double, Distort10WideObjective<double>>(p, bp, s);
    if (p.model == DAVO_MODEL_JOINT) {
        if (p.V < 1 || p.V > kMaxViews) return DAVO_ERR_UNSUPPORTED;
        return launch_train_backward_t<double, JointObjective<double, true>>(p, bp, s);
    }
    if (p.model == DAVO_MODEL_ANGLE_BA) return launch_train_backward_t<double, AngleBAObjective<double, 0, 0, true>>(p, bp, s);
    if (analytic_model(p.model)) return launch_train_backward_t<double, AnalyticWideObjective<double>>(p, bp, s);
    return DAVO_ERR_UNSUPPORTED;
}

template <typename T>
int launch_eval_generic(const SolveParams<T>& p, cudaStream_t s) {
    if (p.model != DAVO_MODEL_DISTORT10) return DAVO_ERR_UNSUPPORTED;
    const int st = launch_wide<T, Distort10WideObjective<T>, WMode::kEval>(p, s);
    if (st != DAVO_ERR_UNSUPPORTED) return st;
    return launch_wide<T, Distort10WideObjective<T, true>, WMode::kEval>(p, s);
}
template int launch_eval_generic<float>(const SolveParams<float>&, cudaStream_t);
template int launch_eval_generic<double>(const SolveParams<double>&, cudaStream_t);

template <typename T>
int launch_line_search_generic(const SolveParams<T>& p, cudaStream_t s) {
    if (p.model == DAVO_MODEL_DISTORT10) {
        const int st = launch_wide<T, Distort10WideObjective<T>, WMode::kLineSearch>(p, s);
        if (st != DAVO_ERR_UNSUPPORTED) return st;
        return launch_wide<T, Distort10WideObjective<T, true>, WMode::kLineSearch>(p, s);
    }
    if (p.model == DAVO_MODEL_JOINT) {
        if (p.V < 1 || p.V > kMaxViews) return DAVO_ERR_UNSUPPORTED;
        if (p.n > 64) return launch_wide<T, JointObjective<T>, WMode::kLineSearch, 4>(p, s);
        return launch_wide<T, JointObjective<T>, WMode::kLineSearch>(p, s);
    }
    if (p.model == DAVO_MODEL_ANGLE_BA) {
        if (p.n > 64) return launch_wide<T, AngleBAObjective<T>, WMode::kLineSearch, 4>(p, s);
        return launch_wide<T, AngleBAObjective<T>, WMode::kLineSearch>(p, s);
    }
    if (analytic_model(p.model)) return launch_wide<T, AnalyticWideObjective<T>, WMode::kLineSearch>(p, s);
    return DAVO_ERR_UNSUPPORTED;
}
template int launch_line_search_generic<float>(const SolveParams<float>&, cudaStream_t);
template int launch_line_search_generic<double>(const SolveParams<double>&, cudaStream_t);

template int launch_train_forward<float>(const SolveParams<float>&, const TrainRecorder<float>&, cudaStream_t);
template int launch_train_forward<double>(const SolveParams<double>&, const TrainRecorder<double>&, cudaStream_t);

}  // namespace davo
