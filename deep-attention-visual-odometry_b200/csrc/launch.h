// launch.h — host-side launcher declarations shared by the .cu translation units and the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include "davo_common.cuh"
#include "train_params.cuh"

namespace davo {

void count_launch();  // increments the process-wide counter behind davo_launch_count()

// Opt a kernel in to `smem` bytes of dynamic shared memory.  The attribute is process-wide per (device, kernel), so
// it is only ever RAISED, under a mutex: two host threads launching the same kernel with different N cannot lower
// each other's limit between the set and the launch.  Returns false on a CUDA error.
bool ensure_dynamic_smem(const void* kernel, size_t smem);

int launch_solve_warp_f32(const SolveParams<float>& p, cudaStream_t s);
int launch_solve_warp_f64(const SolveParams<double>& p, cudaStream_t s);
int launch_line_search_warp_f32(const SolveParams<float>& p, cudaStream_t s);
int launch_line_search_warp_f64(const SolveParams<double>& p, cudaStream_t s);
int launch_eval_warp_f32(const SolveParams<float>& p, cudaStream_t s);
int launch_eval_warp_f64(const SolveParams<double>& p, cudaStream_t s);

int launch_solve_joint_f32(const SolveParams<float>& p, cudaStream_t s);
int launch_solve_joint_f64(const SolveParams<double>& p, cudaStream_t s);
int launch_eval_joint_f32(const SolveParams<float>& p, cudaStream_t s);
int launch_eval_joint_f64(const SolveParams<double>& p, cudaStream_t s);
int launch_line_search_joint_f32(const SolveParams<float>& p, cudaStream_t s);
int launch_line_search_joint_f64(const SolveParams<double>& p, cudaStream_t s);

int launch_solve_ba_f32(const SolveParams<float>& p, cudaStream_t s);
int launch_solve_ba_f64(const SolveParams<double>& p, cudaStream_t s);
int launch_eval_ba_f32(const SolveParams<float>& p, cudaStream_t s);
int launch_eval_ba_f64(const SolveParams<double>& p, cudaStream_t s);
int launch_line_search_ba_f32(const SolveParams<float>& p, cudaStream_t s);
int launch_line_search_ba_f64(const SolveParams<double>& p, cudaStream_t s);

template <typename T>
int launch_train_forward(const SolveParams<T>& p, const TrainRecorder<T>& rec, cudaStream_t s);
int launch_train_backward(const SolveParams<double>& p, const BackwardParams<double>& bp, cudaStream_t s);
// davo_line_search with the secant zoom for the models whose default line search runs on the specialised kernels
template <typename T>
int launch_line_search_generic(const SolveParams<T>& p, cudaStream_t s);
// DISTORT10 cost + gradient on the generic kernels (the route for N beyond the specialised kernels' shared-memory slab)
template <typename T>
int launch_eval_generic(const SolveParams<T>& p, cudaStream_t s);
// interpolate_alpha forward (out != NULL) or backward (grad_out != NULL), elementwise
template <typename T>
int launch_interpolate_alpha(long long k, const T* a1, const T* a2, const T* v1, const T* v2, T* out, const T* grad_out,
                             T* g_a1, T* g_a2, T* g_v1, T* g_v2, cudaStream_t s);

template <typename T>
int launch_stage(int B, int N, const T* pts, const T* obs, const T* pose, T* staged, cudaStream_t s);
template <typename T>
int launch_project(int B, int N, const T* pts, const T* th16, T* u, T* v, T* J, cudaStream_t s);
template <typename T>
int launch_least_squares(int B, int R, int P, const T* res, const T* jac, const T* w, T* err, T* grad,
                         cudaStream_t s);
template <typename T>
int launch_bfgs_update(int k, int n, T* H, const T* s_, const T* y, cudaStream_t s);
template <typename T>
int launch_bfgs_initial_scale(int k, int n, const T* s_, const T* y, T* scale, cudaStream_t s);

// the initial-guess MLP on the tensor cores (mlp_kernels.cu)
long long launch_mlp_packed_bytes(int N, int K);
int launch_mlp_pack_weights(int N, int K, const float* W, void* packed, cudaStream_t s);
int launch_mlp_forward(int B, int K1, int H, int P, const float* x, const void* w1, const float* b1, const float* s1,
                       const float* t1, const void* w2, const float* b2, const float* s2, const float* t2,
                       const void* w3, const float* b3, float* out, cudaStream_t s);

template <typename T>
int launch_generate_distort10(const davo_generator_desc* d, T* pts, T* obs, T* pose, T* x0, T* truth, cudaStream_t s);
template <typename T>
int launch_generate_joint(const davo_generator_desc* d, T* pts, T* obs, T* x0, T* truth, cudaStream_t s);
template <typename T>
int launch_generate_views(const davo_generator_desc* d, T* projected, T* visibility, T* intrinsics, T* orientations,
                          T* translations, T* world, T* x0, T* truth, cudaStream_t s);

}  // namespace davo
