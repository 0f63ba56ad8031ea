// joint_kernels.cu — CTA-per-problem kernels for the JOINT model (intrinsics + per-view pose).
#include "davo_common.cuh"
#include "launch.h"

namespace davo {
// Placeholder until the CTA-per-problem solver lands: the ABI reports the model as unsupported.
int launch_solve_joint_f32(const SolveParams<float>&, cudaStream_t) { return DAVO_ERR_UNSUPPORTED; }
int launch_solve_joint_f64(const SolveParams<double>&, cudaStream_t) { return DAVO_ERR_UNSUPPORTED; }
int launch_eval_joint_f32(const SolveParams<float>&, cudaStream_t) { return DAVO_ERR_UNSUPPORTED; }
int launch_eval_joint_f64(const SolveParams<double>&, cudaStream_t) { return DAVO_ERR_UNSUPPORTED; }
int launch_line_search_joint_f32(const SolveParams<float>&, cudaStream_t) { return DAVO_ERR_UNSUPPORTED; }
int launch_line_search_joint_f64(const SolveParams<double>&, cudaStream_t) { return DAVO_ERR_UNSUPPORTED; }
}  // namespace davo
