// ba_kernels.cu — the ANGLE_BA model (the entry script's bundle-adjustment objective,
// networks/calibration_network.py:58-67; n = 3 + 3N + 6(V-1) <= 128) on the generic warp-per-problem
// kernels of wide_kernel.cuh: solve, line search, cost + gradient.
#include "davo_common.cuh"
#include "objectives_ba.cuh"
#include "wide_kernel.cuh"
#include "launch.h"

namespace davo {

template <typename T, WMode kMode>
static int launch_ba(const SolveParams<T>& p, cudaStream_t s) {
    if (p.V < 2 || p.N < 1 || p.n != 3 + 3 * p.N + 6 * (p.V - 1)) return DAVO_ERR_BAD_SHAPE;
    // the entry script's configuration (4 views x 8 points, n = 45) has its own instantiation
    if (p.V == 4 && p.N == 8) return launch_wide<T, AngleBAObjective<T, 4, 8>, kMode>(p, s);
    if (p.n > 64) return launch_wide<T, AngleBAObjective<T>, kMode, 4>(p, s);  // 65..128 parameters: 4 components per lane
    return launch_wide<T, AngleBAObjective<T>, kMode>(p, s);
}

int launch_solve_ba_f32(const SolveParams<float>& p, cudaStream_t s) { return launch_ba<float, WMode::kSolve>(p, s); }
int launch_solve_ba_f64(const SolveParams<double>& p, cudaStream_t s) { return launch_ba<double, WMode::kSolve>(p, s); }
int launch_eval_ba_f32(const SolveParams<float>& p, cudaStream_t s) { return launch_ba<float, WMode::kEval>(p, s); }
int launch_eval_ba_f64(const SolveParams<double>& p, cudaStream_t s) { return launch_ba<double, WMode::kEval>(p, s); }
int launch_line_search_ba_f32(const SolveParams<float>& p, cudaStream_t s) { return launch_ba<float, WMode::kLineSearch>(p, s); }
int launch_line_search_ba_f64(const SolveParams<double>& p, cudaStream_t s) { return launch_ba<double, WMode::kLineSearch>(p, s); }

}  // namespace davo
