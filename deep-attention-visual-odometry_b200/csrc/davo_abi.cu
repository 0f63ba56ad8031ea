// davo_abi.cu — the extern "C" boundary declared in include/davo_b200.h: argument validation,
// descriptor -> typed device parameters, dispatch to the kernel launchers.  No torch types, no
// allocation, no host synchronisation.
#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "davo_common.cuh"
#include "launch.h"

namespace davo {
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool ensure_dynamic_smem(const void* kernel, size_t smem) {
    if (smem <= 48 * 1024) return true;
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> granted;  // (device, kernel) -> bytes opted in so far
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = granted[{dev, kernel}];
    if (smem <= have) return true;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return false;
    have = smem;
    return true;
}

#if DAVO_TRACE
static void* g_trace = nullptr;
static int g_trace_problem = -1, g_trace_capacity = 0;
#endif

static bool is_analytic(int model) { return model >= DAVO_MODEL_SPHERE && model <= DAVO_MODEL_DISTANCE; }
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_desc(const davo_problem_desc* d) {
    if (!d) return DAVO_ERR_NULL_POINTER;
    if (d->B < 0 || d->N < 0 || d->n < 1 || d->V < 1) return DAVO_ERR_BAD_SHAPE;
    if (d->dtype != DAVO_F32 && d->dtype != DAVO_F64) return DAVO_ERR_UNSUPPORTED;
    if (d->max_iters < 0 || d->max_ls_iters < 1) return DAVO_ERR_BAD_ARGUMENT;
    if (d->model == DAVO_MODEL_DISTORT10) {
        if (d->n != 10 || d->V != 1) return DAVO_ERR_BAD_SHAPE;
    } else if (d->model == DAVO_MODEL_JOINT) {
        if (d->n != 10 + 6 * d->V) return DAVO_ERR_BAD_SHAPE;
    } else if (d->model == DAVO_MODEL_ANGLE_BA) {
        if (d->V < 2 || d->N < 1 || d->n != 3 + 3 * d->N + 6 * (d->V - 1)) return DAVO_ERR_BAD_SHAPE;
    } else if (is_analytic(d->model)) {
        if (d->model == DAVO_MODEL_ROSENBROCK && d->n != 2) return DAVO_ERR_BAD_SHAPE;
        if (d->n > kSlots) return DAVO_ERR_UNSUPPORTED;
    } else {
        return DAVO_ERR_UNSUPPORTED;
    }
    return DAVO_OK;
}

static int check_data(const davo_problem_desc* d, const void* data0, const void* data1, const void* w) {
    if (d->B == 0) return DAVO_OK;
    if (d->model == DAVO_MODEL_DISTORT10) {
        if (!data0) return DAVO_ERR_NULL_POINTER;
        if (!aligned16(data0)) return DAVO_ERR_MISALIGNED;  // bulk-TMA source rows
    } else if (d->model == DAVO_MODEL_JOINT) {
        if (!data0 || !data1) return DAVO_ERR_NULL_POINTER;
        if (!aligned16(data0) || !aligned16(data1)) return DAVO_ERR_MISALIGNED;
    } else if (d->model == DAVO_MODEL_DISTANCE || d->model == DAVO_MODEL_ANGLE_BA) {
        if (!data0) return DAVO_ERR_NULL_POINTER;
    }
    if (d->has_weights && !w) return DAVO_ERR_NULL_POINTER;
    return DAVO_OK;
}

template <typename T>
static SolveParams<T> make_params(const davo_problem_desc* d, const void* data0, const void* data1, const void* w) {
    SolveParams<T> p{};
    p.B = d->B; p.N = d->N; p.V = d->V; p.n = d->n; p.model = d->model;
    p.max_iters = d->max_iters; p.max_ls = d->max_ls_iters; p.strong = d->strong; p.has_w = d->has_weights;
    p.zoom = d->zoom_interpolation != 0;
    // thresholds are rounded to the arithmetic type exactly like torch rounds a Python float that meets a tensor
    p.c1 = static_cast<T>(d->sufficient_decrease);
    p.c2 = static_cast<T>(d->curvature);
    p.thr = static_cast<T>(d->error_threshold);
    p.min_step = static_cast<T>(d->minimum_step);
    p.data0 = static_cast<const T*>(data0);
    p.data1 = static_cast<const T*>(data1);
    p.w = d->has_weights ? static_cast<const T*>(w) : nullptr;
#if DAVO_TRACE
    p.trace = static_cast<T*>(g_trace); p.trace_problem = g_trace_problem; p.trace_capacity = g_trace_capacity;
#endif
    return p;
}

// model -> launcher (f32 / f64 instantiations live in solve_kernels.cu, joint_kernels.cu, ba_kernels.cu)
#define DAVO_DISPATCH(what, suffix, p, s)                                           \
    ((p).model == DAVO_MODEL_JOINT      ? launch_##what##_joint_##suffix((p), (s))  \
     : (p).model == DAVO_MODEL_ANGLE_BA ? launch_##what##_ba_##suffix((p), (s))     \
                                        : launch_##what##_warp_##suffix((p), (s)))
}  // namespace davo

using namespace davo;

extern "C" {

#if DAVO_TRACE
// debug builds only (not part of the ABI): record 8 values per accepted step of one problem
void davo_debug_trace(void* buffer, int problem, int capacity) {
    g_trace = buffer; g_trace_problem = problem; g_trace_capacity = capacity;
}
#endif

int davo_abi_version(void) { return DAVO_ABI_VERSION; }

const char* davo_strerror(int status) {
    switch (status) {
        case DAVO_OK: return "ok";
        case DAVO_ERR_NULL_POINTER: return "a required pointer is NULL";
        case DAVO_ERR_BAD_SHAPE: return "B, N, V, n are inconsistent with the model";
        case DAVO_ERR_UNSUPPORTED: return "model / n / N / dtype is outside what this build supports";
        case DAVO_ERR_MISALIGNED: return "a vector-loaded buffer is not 16-byte aligned";
        case DAVO_ERR_CUDA: return "CUDA launch or runtime error";
        case DAVO_ERR_BAD_ARGUMENT: return "bad argument";
        default: return "unknown davo status";
    }
}

int64_t davo_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int davo_solve_calibration(const davo_problem_desc* desc, const void* data0, const void* data1,
                           const void* weights, const void* x0, void* x_out, void* cost_out,
                           uint8_t* converged_out, int32_t* iters_out, int32_t* fevals_out,
                           int32_t* reason_out, void* workspace, void* stream) {
    int st = check_desc(desc);
    if (st) return st;
    if (desc->B == 0) return DAVO_OK;
    if (!x0 || !x_out || !workspace) return DAVO_ERR_NULL_POINTER;
    if ((st = check_data(desc, data0, data1, weights))) return st;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (desc->dtype == DAVO_F32) {
        SolveParams<float> p = make_params<float>(desc, data0, data1, weights);
        p.x0 = static_cast<const float*>(x0); p.x_out = static_cast<float*>(x_out);
        p.cost_out = static_cast<float*>(cost_out); p.converged_out = converged_out;
        p.iters_out = iters_out; p.fevals_out = fevals_out; p.reason_out = reason_out;
        p.queue = static_cast<unsigned*>(workspace);
        if (p.zoom) return launch_train_forward<float>(p, TrainRecorder<float>{nullptr, nullptr, nullptr, nullptr, 0, 0.0f, 0, 0, 1}, s);
        st = DAVO_DISPATCH(solve, f32, p, s);
        if (st == DAVO_ERR_UNSUPPORTED && p.model == DAVO_MODEL_DISTORT10)   // N beyond the specialised kernels' slab
            st = launch_train_forward<float>(p, TrainRecorder<float>{nullptr, nullptr, nullptr, nullptr, 0, 0.0f, 0, 0, 0}, s);
        return st;
    }
    SolveParams<double> p = make_params<double>(desc, data0, data1, weights);
    p.x0 = static_cast<const double*>(x0); p.x_out = static_cast<double*>(x_out);
    p.cost_out = static_cast<double*>(cost_out); p.converged_out = converged_out;
    p.iters_out = iters_out; p.fevals_out = fevals_out; p.reason_out = reason_out;
    p.queue = static_cast<unsigned*>(workspace);
    if (p.zoom) return launch_train_forward<double>(p, TrainRecorder<double>{nullptr, nullptr, nullptr, nullptr, 0, 0.0f, 0, 0, 1}, s);
    st = DAVO_DISPATCH(solve, f64, p, s);
    if (st == DAVO_ERR_UNSUPPORTED && p.model == DAVO_MODEL_DISTORT10)
        st = launch_train_forward<double>(p, TrainRecorder<double>{nullptr, nullptr, nullptr, nullptr, 0, 0.0f, 0, 0, 0}, s);
    return st;
}

int davo_solve_training(const davo_problem_desc* desc, const davo_training_desc* train, const void* data0,
                        const void* data1, const void* weights, const void* x0, void* x_out, void* cost_out,
                        uint8_t* converged_out, int32_t* iters_out, int32_t* fevals_out, int32_t* reason_out,
                        void* traj_x, void* traj_g, void* traj_alpha, int32_t* traj_len, void* workspace,
                        void* stream) {
    int st = check_desc(desc);
    if (st) return st;
    if (!train) return DAVO_ERR_NULL_POINTER;
    if (desc->B == 0) return DAVO_OK;
    if (!x0 || !x_out || !workspace) return DAVO_ERR_NULL_POINTER;
    if ((st = check_data(desc, data0, data1, weights))) return st;
    const bool recording = traj_x || traj_g || traj_alpha || traj_len;
    if (recording && !(traj_x && traj_g && traj_alpha && traj_len)) return DAVO_ERR_NULL_POINTER;
    if (recording && train->capacity < desc->max_iters) return DAVO_ERR_BAD_ARGUMENT;
    if (!(train->drop_path_p >= 0.0 && train->drop_path_p <= 1.0)) return DAVO_ERR_BAD_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (desc->dtype == DAVO_F32) {
        SolveParams<float> p = make_params<float>(desc, data0, data1, weights);
        p.x0 = static_cast<const float*>(x0); p.x_out = static_cast<float*>(x_out);
        p.cost_out = static_cast<float*>(cost_out); p.converged_out = converged_out;
        p.iters_out = iters_out; p.fevals_out = fevals_out; p.reason_out = reason_out;
        p.queue = static_cast<unsigned*>(workspace);
        TrainRecorder<float> rec{static_cast<float*>(traj_x), static_cast<float*>(traj_g),
                                 static_cast<float*>(traj_alpha), traj_len, train->capacity,
                                 static_cast<float>(train->drop_path_p), train->seed, train->return_second_last,
                                 desc->zoom_interpolation != 0};
        return launch_train_forward<float>(p, rec, s);
    }
    SolveParams<double> p = make_params<double>(desc, data0, data1, weights);
    p.x0 = static_cast<const double*>(x0); p.x_out = static_cast<double*>(x_out);
    p.cost_out = static_cast<double*>(cost_out); p.converged_out = converged_out;
    p.iters_out = iters_out; p.fevals_out = fevals_out; p.reason_out = reason_out;
    p.queue = static_cast<unsigned*>(workspace);
    TrainRecorder<double> rec{static_cast<double*>(traj_x), static_cast<double*>(traj_g),
                              static_cast<double*>(traj_alpha), traj_len, train->capacity,
                              static_cast<float>(train->drop_path_p), train->seed, train->return_second_last,
                                 desc->zoom_interpolation != 0};
    return launch_train_forward<double>(p, rec, s);
}

int davo_solve_backward(const davo_problem_desc* desc, const davo_training_desc* train, const void* data0,
                        const void* data1, const void* weights, const void* traj_x, const void* traj_g,
                        const void* traj_alpha, const int32_t* traj_len, const int64_t* traj_offset,
                        const int64_t* scratch_offset, void* scratch, const void* grad_out, void* grad_x0,
                        void* grad_data, void* workspace, void* stream) {
    int st = check_desc(desc);
    if (st) return st;
    if (!train) return DAVO_ERR_NULL_POINTER;
    if (desc->dtype != DAVO_F64) return DAVO_ERR_UNSUPPORTED;
    if (desc->B == 0) return DAVO_OK;
    if (!traj_x || !traj_g || !traj_alpha || !traj_len || !traj_offset || !scratch_offset || !grad_out || !grad_x0 || !workspace)
        return DAVO_ERR_NULL_POINTER;  // scratch may be NULL when no problem recorded a step
    if ((st = check_data(desc, data0, data1, weights))) return st;
    SolveParams<double> p = make_params<double>(desc, data0, data1, weights);
    p.queue = static_cast<unsigned*>(workspace);
    BackwardParams<double> bp{static_cast<const double*>(traj_x), static_cast<const double*>(traj_g),
                              static_cast<const double*>(traj_alpha), traj_len, traj_offset, scratch_offset,
                              static_cast<double*>(scratch), static_cast<const double*>(grad_out),
                              static_cast<double*>(grad_x0), static_cast<double*>(grad_data),
                              train->hvp_rel_step > 0.0 ? train->hvp_rel_step : 5e-7};
    return launch_train_backward(p, bp, static_cast<cudaStream_t>(stream));
}

int davo_eval_cost_grad(const davo_problem_desc* desc, const void* data0, const void* data1,
                        const void* weights, const void* x, void* cost, void* grad, void* stream) {
    int st = check_desc(desc);
    if (st) return st;
    if (desc->B == 0) return DAVO_OK;
    if (!x) return DAVO_ERR_NULL_POINTER;
    if ((st = check_data(desc, data0, data1, weights))) return st;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (desc->dtype == DAVO_F32) {
        SolveParams<float> p = make_params<float>(desc, data0, data1, weights);
        p.x0 = static_cast<const float*>(x); p.cost_out = static_cast<float*>(cost);
        p.x_out = static_cast<float*>(grad);
        st = DAVO_DISPATCH(eval, f32, p, s);
        if (st == DAVO_ERR_UNSUPPORTED && p.model == DAVO_MODEL_DISTORT10) st = launch_eval_generic<float>(p, s);
        return st;
    }
    SolveParams<double> p = make_params<double>(desc, data0, data1, weights);
    p.x0 = static_cast<const double*>(x); p.cost_out = static_cast<double*>(cost);
    p.x_out = static_cast<double*>(grad);
    st = DAVO_DISPATCH(eval, f64, p, s);
    if (st == DAVO_ERR_UNSUPPORTED && p.model == DAVO_MODEL_DISTORT10) st = launch_eval_generic<double>(p, s);
    return st;
}

int davo_line_search(const davo_problem_desc* desc, const void* data0, const void* data1,
                     const void* weights, const void* x, const void* direction, const void* base_cost,
                     const void* base_grad, void* alpha_out, int32_t* fevals_out, void* stream) {
    int st = check_desc(desc);
    if (st) return st;
    if (desc->B == 0) return DAVO_OK;
    if (!x || !direction || !base_cost || !base_grad || !alpha_out) return DAVO_ERR_NULL_POINTER;
    if ((st = check_data(desc, data0, data1, weights))) return st;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (desc->dtype == DAVO_F32) {
        SolveParams<float> p = make_params<float>(desc, data0, data1, weights);
        p.x0 = static_cast<const float*>(x); p.dir = static_cast<const float*>(direction);
        p.base_cost = static_cast<const float*>(base_cost); p.base_grad = static_cast<const float*>(base_grad);
        p.alpha_out = static_cast<float*>(alpha_out); p.fevals_out = fevals_out;
        if (p.zoom) return launch_line_search_generic<float>(p, s);
        st = DAVO_DISPATCH(line_search, f32, p, s);
        if (st == DAVO_ERR_UNSUPPORTED && p.model == DAVO_MODEL_DISTORT10) st = launch_line_search_generic<float>(p, s);
        return st;
    }
    SolveParams<double> p = make_params<double>(desc, data0, data1, weights);
    p.x0 = static_cast<const double*>(x); p.dir = static_cast<const double*>(direction);
    p.base_cost = static_cast<const double*>(base_cost); p.base_grad = static_cast<const double*>(base_grad);
    p.alpha_out = static_cast<double*>(alpha_out); p.fevals_out = fevals_out;
    if (p.zoom) return launch_line_search_generic<double>(p, s);
    st = DAVO_DISPATCH(line_search, f64, p, s);
    if (st == DAVO_ERR_UNSUPPORTED && p.model == DAVO_MODEL_DISTORT10) st = launch_line_search_generic<double>(p, s);
    return st;
}

int davo_interpolate_alpha(int32_t dtype, int64_t k, const void* alpha_1, const void* alpha_2, const void* value_1,
                           const void* value_2, void* out, void* stream) {
    if (k < 0) return DAVO_ERR_BAD_SHAPE;
    if (k == 0) return DAVO_OK;
    if (!alpha_1 || !alpha_2 || !value_1 || !value_2 || !out) return DAVO_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == DAVO_F32)
        return launch_interpolate_alpha<float>(k, static_cast<const float*>(alpha_1), static_cast<const float*>(alpha_2),
                                               static_cast<const float*>(value_1), static_cast<const float*>(value_2),
                                               static_cast<float*>(out), nullptr, nullptr, nullptr, nullptr, nullptr, s);
    if (dtype == DAVO_F64)
        return launch_interpolate_alpha<double>(k, static_cast<const double*>(alpha_1), static_cast<const double*>(alpha_2),
                                                static_cast<const double*>(value_1), static_cast<const double*>(value_2),
                                                static_cast<double*>(out), nullptr, nullptr, nullptr, nullptr, nullptr, s);
    return DAVO_ERR_UNSUPPORTED;
}

int davo_interpolate_alpha_backward(int32_t dtype, int64_t k, const void* alpha_1, const void* alpha_2,
                                    const void* value_1, const void* value_2, const void* grad_out,
                                    void* grad_alpha_1, void* grad_alpha_2, void* grad_value_1, void* grad_value_2,
                                    void* stream) {
    if (k < 0) return DAVO_ERR_BAD_SHAPE;
    if (k == 0) return DAVO_OK;
    if (!alpha_1 || !alpha_2 || !value_1 || !value_2 || !grad_out) return DAVO_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == DAVO_F32)
        return launch_interpolate_alpha<float>(k, static_cast<const float*>(alpha_1), static_cast<const float*>(alpha_2),
                                               static_cast<const float*>(value_1), static_cast<const float*>(value_2),
                                               nullptr, static_cast<const float*>(grad_out),
                                               static_cast<float*>(grad_alpha_1), static_cast<float*>(grad_alpha_2),
                                               static_cast<float*>(grad_value_1), static_cast<float*>(grad_value_2), s);
    if (dtype == DAVO_F64)
        return launch_interpolate_alpha<double>(k, static_cast<const double*>(alpha_1), static_cast<const double*>(alpha_2),
                                                static_cast<const double*>(value_1), static_cast<const double*>(value_2),
                                                nullptr, static_cast<const double*>(grad_out),
                                                static_cast<double*>(grad_alpha_1), static_cast<double*>(grad_alpha_2),
                                                static_cast<double*>(grad_value_1), static_cast<double*>(grad_value_2), s);
    return DAVO_ERR_UNSUPPORTED;
}

int64_t davo_mlp_packed_bytes(int32_t N, int32_t K) { return launch_mlp_packed_bytes(N, K); }

int davo_mlp_pack_weights(int32_t N, int32_t K, const void* weight, void* packed, void* stream) {
    if (!weight || !packed) return DAVO_ERR_NULL_POINTER;
    if (!aligned16(weight) || !aligned16(packed)) return DAVO_ERR_MISALIGNED;
    return launch_mlp_pack_weights(N, K, static_cast<const float*>(weight), packed, static_cast<cudaStream_t>(stream));
}

int davo_mlp_forward(const davo_mlp_desc* desc, const void* x, const void* w1_packed, const void* b1,
                     const void* scale1, const void* shift1, const void* w2_packed, const void* b2,
                     const void* scale2, const void* shift2, const void* w3_packed, const void* b3, void* x0_out,
                     void* stream) {
    if (!desc) return DAVO_ERR_NULL_POINTER;
    if (desc->B < 0) return DAVO_ERR_BAD_SHAPE;
    if (desc->B == 0) return DAVO_OK;
    if (!x || !w1_packed || !b1 || !scale1 || !shift1 || !w2_packed || !b2 || !scale2 || !shift2 || !w3_packed || !b3 ||
        !x0_out)
        return DAVO_ERR_NULL_POINTER;
    if (!aligned16(x) || !aligned16(w1_packed) || !aligned16(w2_packed) || !aligned16(w3_packed) || !aligned16(b1) ||
        !aligned16(scale1) || !aligned16(shift1) || !aligned16(b2) || !aligned16(scale2) || !aligned16(shift2))
        return DAVO_ERR_MISALIGNED;
    auto f = [](const void* q) { return static_cast<const float*>(q); };
    return launch_mlp_forward(desc->B, desc->in_features, desc->hidden, desc->out_features, f(x), w1_packed, f(b1),
                              f(scale1), f(shift1), w2_packed, f(b2), f(scale2), f(shift2), w3_packed, f(b3),
                              static_cast<float*>(x0_out), static_cast<cudaStream_t>(stream));
}

int davo_stage_matches(const davo_problem_desc* desc, const void* points_3d, const void* obs,
                       const void* pose, void* staged, void* stream) {
    if (!desc) return DAVO_ERR_NULL_POINTER;
    if (desc->B < 0 || desc->N < 0) return DAVO_ERR_BAD_SHAPE;
    if (desc->B == 0 || desc->N == 0) return DAVO_OK;
    if (!points_3d || !obs || !staged) return DAVO_ERR_NULL_POINTER;
    if (!aligned16(staged)) return DAVO_ERR_MISALIGNED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (desc->dtype == DAVO_F32)
        return launch_stage<float>(desc->B, desc->N, static_cast<const float*>(points_3d),
                                   static_cast<const float*>(obs), static_cast<const float*>(pose),
                                   static_cast<float*>(staged), s);
    if (desc->dtype == DAVO_F64)
        return launch_stage<double>(desc->B, desc->N, static_cast<const double*>(points_3d),
                                    static_cast<const double*>(obs), static_cast<const double*>(pose),
                                    static_cast<double*>(staged), s);
    return DAVO_ERR_UNSUPPORTED;
}

static int project_impl(const davo_problem_desc* desc, const void* pts, const void* th, void* J, void* u, void* v,
                        void* stream, bool want_j) {
    if (!desc) return DAVO_ERR_NULL_POINTER;
    if (desc->B < 0 || desc->N < 0) return DAVO_ERR_BAD_SHAPE;
    if (desc->B == 0 || desc->N == 0) return DAVO_OK;
    if (!pts || !th || !u || !v || (want_j && !J)) return DAVO_ERR_NULL_POINTER;
    if (want_j && !aligned16(J)) return DAVO_ERR_MISALIGNED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (desc->dtype == DAVO_F32)
        return launch_project<float>(desc->B, desc->N, static_cast<const float*>(pts), static_cast<const float*>(th),
                                     static_cast<float*>(u), static_cast<float*>(v),
                                     want_j ? static_cast<float*>(J) : nullptr, s);
    if (desc->dtype == DAVO_F64)
        return launch_project<double>(desc->B, desc->N, static_cast<const double*>(pts),
                                      static_cast<const double*>(th), static_cast<double*>(u),
                                      static_cast<double*>(v), want_j ? static_cast<double*>(J) : nullptr, s);
    return DAVO_ERR_UNSUPPORTED;
}

int davo_project(const davo_problem_desc* desc, const void* points_3d, const void* params16, void* u, void* v,
                 void* stream) {
    return project_impl(desc, points_3d, params16, nullptr, u, v, stream, false);
}

int davo_project_jacobian(const davo_problem_desc* desc, const void* points_3d, const void* params16, void* J,
                          void* u, void* v, void* stream) {
    return project_impl(desc, points_3d, params16, J, u, v, stream, true);
}

int davo_least_squares(int32_t dtype, int32_t B, int32_t R, int32_t P, const void* residuals,
                       const void* jacobian, const void* weights, void* error, void* gradient, void* stream) {
    if (B < 0 || R < 0 || P < 0) return DAVO_ERR_BAD_SHAPE;
    if (B == 0) return DAVO_OK;
    if (!residuals || (!error && !gradient) || (gradient && !jacobian)) return DAVO_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == DAVO_F32)
        return launch_least_squares<float>(B, R, P, static_cast<const float*>(residuals),
                                           static_cast<const float*>(jacobian), static_cast<const float*>(weights),
                                           static_cast<float*>(error), static_cast<float*>(gradient), s);
    if (dtype == DAVO_F64)
        return launch_least_squares<double>(B, R, P, static_cast<const double*>(residuals),
                                            static_cast<const double*>(jacobian),
                                            static_cast<const double*>(weights), static_cast<double*>(error),
                                            static_cast<double*>(gradient), s);
    return DAVO_ERR_UNSUPPORTED;
}

int davo_bfgs_update(int32_t dtype, int32_t k, int32_t n, void* H, const void* s_, const void* y, void* stream) {
    if (k < 0 || n < 1) return DAVO_ERR_BAD_SHAPE;
    if (k == 0) return DAVO_OK;
    if (!H || !s_ || !y) return DAVO_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == DAVO_F32)
        return launch_bfgs_update<float>(k, n, static_cast<float*>(H), static_cast<const float*>(s_),
                                         static_cast<const float*>(y), s);
    if (dtype == DAVO_F64)
        return launch_bfgs_update<double>(k, n, static_cast<double*>(H), static_cast<const double*>(s_),
                                          static_cast<const double*>(y), s);
    return DAVO_ERR_UNSUPPORTED;
}

int davo_bfgs_initial_scale(int32_t dtype, int32_t k, int32_t n, const void* s_, const void* y, void* scale,
                            void* stream) {
    if (k < 0 || n < 1) return DAVO_ERR_BAD_SHAPE;
    if (k == 0) return DAVO_OK;
    if (!s_ || !y || !scale) return DAVO_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == DAVO_F32)
        return launch_bfgs_initial_scale<float>(k, n, static_cast<const float*>(s_), static_cast<const float*>(y),
                                                static_cast<float*>(scale), s);
    if (dtype == DAVO_F64)
        return launch_bfgs_initial_scale<double>(k, n, static_cast<const double*>(s_),
                                                 static_cast<const double*>(y), static_cast<double*>(scale), s);
    return DAVO_ERR_UNSUPPORTED;
}

static int check_gen(const davo_generator_desc* d, int min_views) {
    if (!d) return DAVO_ERR_NULL_POINTER;
    if (d->B < 0 || d->N < 1 || d->V < min_views) return DAVO_ERR_BAD_SHAPE;
    if (d->dtype != DAVO_F32 && d->dtype != DAVO_F64) return DAVO_ERR_UNSUPPORTED;
    if (!(d->fov > 0.0) || d->noise < 0.0 || d->pathological < 0.0 || d->pathological > 1.0) return DAVO_ERR_BAD_ARGUMENT;
    return DAVO_OK;
}

int davo_generate_distort10(const davo_generator_desc* desc, void* points_3d, void* obs, void* pose, void* x0,
                            void* truth, void* stream) {
    int st = check_gen(desc, 1);
    if (st) return st;
    if (desc->B == 0) return DAVO_OK;
    if (!points_3d || !obs || !x0) return DAVO_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (desc->dtype == DAVO_F32)
        return launch_generate_distort10<float>(desc, static_cast<float*>(points_3d), static_cast<float*>(obs),
                                                static_cast<float*>(pose), static_cast<float*>(x0),
                                                static_cast<float*>(truth), s);
    return launch_generate_distort10<double>(desc, static_cast<double*>(points_3d), static_cast<double*>(obs),
                                             static_cast<double*>(pose), static_cast<double*>(x0),
                                             static_cast<double*>(truth), s);
}

int davo_generate_joint(const davo_generator_desc* desc, void* points_3d, void* obs, void* x0, void* truth,
                        void* stream) {
    int st = check_gen(desc, 1);
    if (st) return st;
    if (desc->B == 0) return DAVO_OK;
    if (!points_3d || !obs || !x0) return DAVO_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (desc->dtype == DAVO_F32)
        return launch_generate_joint<float>(desc, static_cast<float*>(points_3d), static_cast<float*>(obs),
                                            static_cast<float*>(x0), static_cast<float*>(truth), s);
    return launch_generate_joint<double>(desc, static_cast<double*>(points_3d), static_cast<double*>(obs),
                                         static_cast<double*>(x0), static_cast<double*>(truth), s);
}

int davo_generate_views_and_points(const davo_generator_desc* desc, void* projected_points, void* visibility_mask,
                                   void* camera_intrinsics, void* camera_orientations, void* camera_translations,
                                   void* world_points, void* x0, void* truth, void* stream) {
    int st = check_gen(desc, 2);
    if (st) return st;
    if (desc->B == 0) return DAVO_OK;
    if (!projected_points || !visibility_mask || !camera_intrinsics || !camera_orientations || !camera_translations ||
        !world_points)
        return DAVO_ERR_NULL_POINTER;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define DAVO_GEN_VIEWS(T)                                                                                         \
    launch_generate_views<T>(desc, static_cast<T*>(projected_points), static_cast<T*>(visibility_mask),           \
                             static_cast<T*>(camera_intrinsics), static_cast<T*>(camera_orientations),            \
                             static_cast<T*>(camera_translations), static_cast<T*>(world_points),                 \
                             static_cast<T*>(x0), static_cast<T*>(truth), s)
    return desc->dtype == DAVO_F32 ? DAVO_GEN_VIEWS(float) : DAVO_GEN_VIEWS(double);
#undef DAVO_GEN_VIEWS
}

}  // extern "C"
