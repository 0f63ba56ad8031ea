/*
 * davo_b200.h — C-ABI of the B200-native batched calibration solve.
 *
 * This is the drop-in boundary for ONE hot path of jskinn/deep-attention-visual-odometry:
 * the batched BFGS + strong-Wolfe fit of pinhole intrinsics and radial/tangential
 * distortion.  The reference has no FFI of its own (it is pure Python/PyTorch), so each
 * entry point below names the reference Python symbol it replaces (path:line relative to
 * /root/reference/deep_attention_visual_odometry/).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every data pointer is DEVICE memory owned by the
 *    caller (PyTorch passes tensor.data_ptr()); the library allocates nothing persistent.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All calls are
 *    asynchronous with respect to the host and stream-ordered.
 *  - `dtype` in the descriptor selects the arithmetic type of every data pointer
 *    (DAVO_F32 / DAVO_F64); integer outputs are int32_t / uint8_t in both cases.
 *  - return value: DAVO_OK (0) or a negative davo_status; never throws, never aborts.
 *  - thread-safe and re-entrant: no state is kept between calls except (a) a mutex-guarded, monotone
 *    record of the dynamic-shared-memory size each kernel has been opted in to (the CUDA attribute is
 *    process-wide per kernel, so it is only ever raised) and (b) the atomic launch counter behind
 *    davo_launch_count(), a statistic that no result depends on.
 */
#ifndef DAVO_B200_H
#define DAVO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAVO_ABI_VERSION 2

/* ---- status codes -------------------------------------------------------------------- */
typedef enum davo_status {
    DAVO_OK = 0,
    DAVO_ERR_NULL_POINTER = -1,
    DAVO_ERR_BAD_SHAPE = -2,      /* B,N,V,n inconsistent with the model              */
    DAVO_ERR_UNSUPPORTED = -3,    /* model / n / N / dtype outside what is compiled in */
    DAVO_ERR_MISALIGNED = -4,     /* a vector-loaded buffer is not 16-byte aligned     */
    DAVO_ERR_CUDA = -5,           /* launch or runtime error (see cudaGetLastError)    */
    DAVO_ERR_BAD_ARGUMENT = -6
} davo_status;

/* ---- arithmetic type ----------------------------------------------------------------- */
#define DAVO_F32 0
#define DAVO_F64 1

/* ---- objective models ----------------------------------------------------------------
 * DISTORT10: fit (cx,cy,k1,k2,k3,p1,p2,fx,s,fy) with the camera pose fixed.
 *            camera_model/distorted_camera_model.py:24-103 with columns 10..15 constant.
 *            data0 = staged matches [B,N,4] = {a=x'/z', b=y'/z', u*, v*}; data1 unused.
 * JOINT:     fit the 10 intrinsics + (rx,ry,rz,tx,ty,tz) for each of V views, n = 10+6V, V <= 19
 *            (up to 9 views: one CTA per problem; 10 .. 19: one warp per problem).
 *            data0 = world points [B,N,3] (shared by the V views);
 *            data1 = observations [B,V,N,2].
 * The analytic models restate tests/autograd_solvers/reference_functions.py:20-62 and
 * tests/autograd_solvers/test_bfgs_solver.py:33-46 so that the reference's solver tests
 * can be mirrored through this ABI (data0/data1 unused).
 */
#define DAVO_MODEL_DISTORT10 0
#define DAVO_MODEL_JOINT 1
/* ANGLE_BA:  the entry script's bundle-adjustment objective (networks/calibration_network.py:58-67):
 *            parameters (f, cx, cy | N world points | V-1 translations | V-1 axis-angle rotations),
 *            n = 3 + 3N + 6(V-1) <= 128, V >= 2; error = sum over views and points of
 *            visibility * angle(pixel ray, camera-relative point), with
 *            unpack_calibration_parameters / get_camera_relative_points
 *            (camera_model/calibration_pinhole_camera_model.py:33-117), rotate_vector_axis_angle
 *            (geometry/axis_angle_rotation.py:25-54), pixel_coordinates_to_homogeneous
 *            (geometry/homogeneous_projection.py:21-44) and projective_plane_angle_distance
 *            (geometry/projective_plane_angle_distance.py:20-64).
 *            data0 = true_projected_points [B,V,N,2]; data1 unused;
 *            weights [B,V,N] = visibility mask (has_weights = 0: every point visible). */
#define DAVO_MODEL_ANGLE_BA 2
#define DAVO_MODEL_SPHERE 16        /* sum x^2                                   */
#define DAVO_MODEL_SPHERE_OFFSET 17 /* sum x^2 + 10                              */
#define DAVO_MODEL_LOG_SPHERE 18    /* log(1 + sum x^2)                          */
#define DAVO_MODEL_ROSENBROCK 19    /* (1-x)^2 + 100 (y-x^2)^2, n = 2            */
#define DAVO_MODEL_COSINE 20        /* (1 - x0/|x|) + (1-|x|)^2                  */
#define DAVO_MODEL_X2_SINE 21       /* |x|^2 (sin|x| + 2)                        */
#define DAVO_MODEL_DISTANCE 22      /* |x - target|, data0 = target [B,n]
                                       (tests/autograd_solvers/line_search/test_wolffe_conditions.py:214-305) */

/* Parameter index layout of the 16-parameter camera model (spatial_maths.camera_model_parameters,
 * pinned by tests/camera_model/test_distorted_camera_model.py:13-30 and the Jacobian column
 * order camera_model/distorted_camera_model.py:364-383). */
#define DAVO_CX 0
#define DAVO_CY 1
#define DAVO_K1 2
#define DAVO_K2 3
#define DAVO_K3 4
#define DAVO_P1 5
#define DAVO_P2 6
#define DAVO_FX 7
#define DAVO_S 8
#define DAVO_FY 9
#define DAVO_RX 10
#define DAVO_RY 11
#define DAVO_RZ 12
#define DAVO_TX 13
#define DAVO_TY 14
#define DAVO_TZ 15

/* ---- termination reason (per problem) ------------------------------------------------- */
#define DAVO_REASON_THRESHOLD 0 /* cost <= error_threshold      (bfgs_solver.py:143)            */
#define DAVO_REASON_STEP 1      /* |step| <= minimum_step       (bfgs_solver.py:203-207)        */
#define DAVO_REASON_CAP 2       /* `iterations` outer iterations (bfgs_solver.py:118)           */
#define DAVO_REASON_NAN 3       /* cost is NaN: `NaN > thr` is false, the reference retires it  */
/* 4 is internal (hand-off between the two launches of a DISTORT10 solve) and never returned */
#define DAVO_REASON_DROPPED 5   /* training mode: retired by drop-path (bfgs_solver.py:122-125) */

/* ---- problem descriptor ---------------------------------------------------------------
 * Mirrors BFGSSolver.__init__ (autograd_solvers/bfgs_solver.py:49-60) and the arguments of
 * line_search_wolfe_conditions (autograd_solvers/line_search/wolfe_conditions.py:23-32). */
typedef struct davo_problem_desc {
    int32_t B;            /* independent problems                                           */
    int32_t N;            /* matches per problem per view                                   */
    int32_t V;            /* views (1 for DISTORT10 and the analytic models)                */
    int32_t n;            /* fitted parameters per problem                                  */
    int32_t model;        /* DAVO_MODEL_*                                                   */
    int32_t dtype;        /* DAVO_F32 | DAVO_F64                                            */
    int32_t max_iters;    /* `iterations`, default 1000                                     */
    int32_t max_ls_iters; /* line-search probe cap, 1000 in the reference (wolfe_conditions.py:116) */
    int32_t strong;       /* 1 = strong Wolfe (what BFGSSolver passes, bfgs_solver.py:189)  */
    int32_t has_weights;  /* 1 = `weights` [B,V,N] multiplies each squared residual pair    */
    int32_t zoom_interpolation; /* 0 = bisection zoom (wolfe_conditions.py:128-131, :242-253); 1 = the older
                                 * solvers/ generation's secant zoom: the zoom step is interpolate_alpha(lo, hi,
                                 * phi'(lo), phi'(hi)) (utils/func_interpolate_alpha.py:5-40, as used by
                                 * solvers/line_search_strong_wolfe_conditions.py:147-155); runs on the generic
                                 * one-warp-per-problem solver for every model */
    int32_t reserved0;    /* 0 */
    double sufficient_decrease; /* c1, default 1e-4                                         */
    double curvature;           /* c2, default 0.9                                          */
    double error_threshold;     /* default 1e-4                                             */
    double minimum_step;        /* default 1e-8                                             */
} davo_problem_desc;

/* Bytes of caller-owned scratch every solve / line-search call needs (work-queue counter). */
#define DAVO_WORKSPACE_BYTES 256

int davo_abi_version(void);
const char* davo_strerror(int status);

/* Replaces BFGSSolver.forward in eval mode (autograd_solvers/bfgs_solver.py:80-215) driving the
 * strong-Wolfe line search (autograd_solvers/line_search/wolfe_conditions.py:23-239) on the
 * objective selected by desc->model.  One persistent launch; problems are pulled from an
 * atomic work queue in `workspace`.
 *   x0        [B,n]   initial parameters
 *   x_out     [B,n]   fitted parameters (last accepted iterate; may alias x0)
 *   cost_out  [B]     objective at x_out
 *   converged_out [B] 1 where cost <= error_threshold
 *   iters_out [B]     accepted steps (= line searches the problem took part in)
 *   fevals_out[B]     objective evaluations the reference would have made
 *   reason_out[B]     DAVO_REASON_* (may be NULL)
 * Output pointers are written for all B rows.
 * DISTORT10 with N beyond the specialised kernels' shared-memory slab (N > ~3 500) runs on the generic one-warp-per-
 * problem solver, and beyond ITS slab (~14 000 matches in float32) with the matches read from global memory at every
 * evaluation: no N is refused.  The same fall-back applies to davo_eval_cost_grad and davo_line_search.
 * DISTORT10 without weights, large batches: two launches on `stream` — two problems per warp, then one CTA per
 * straggler (a problem past max(4096, 4 * max_iters) evaluations is abandoned by the first launch and solved again
 * from x0 by the second).  reason_out carries the hand-off flag between them (an internal value that never
 * survives the call); with reason_out == NULL the first launch solves every problem to the end.  `workspace`
 * holds one work-queue counter per launch. */
int davo_solve_calibration(const davo_problem_desc* desc, const void* data0, const void* data1,
                           const void* weights, const void* x0, void* x_out, void* cost_out,
                           uint8_t* converged_out, int32_t* iters_out, int32_t* fevals_out,
                           int32_t* reason_out, void* workspace, void* stream);

/* ---- training-mode / differentiable solve -------------------------------------------------------------
 * Replaces BFGSSolver.forward when `self.training` is set or the parameters require grad
 * (autograd_solvers/bfgs_solver.py:80-215): the caller passes the training thresholds in `desc`
 * (training_error_threshold / training_iterations, :88-93); drop-path (:122-125) retires each problem at the top
 * of every outer iteration with probability drop_path_p, from a counter-based generator that is a pure function
 * of (seed, problem index, iteration); return_second_last (:196-212) withholds the step that retires a problem on
 * its length.  One warp per problem on the generic solver, every model, n <= 128. */
typedef struct davo_training_desc {
    int32_t capacity;           /* rows per problem of the trajectory buffers (>= max_iters records every step) */
    int32_t return_second_last; /* BFGSSolver(return_second_last=...)                                           */
    double drop_path_p;         /* BFGSSolver(drop_path_p=...); 0 = never                                       */
    uint64_t seed;
    double hvp_rel_step;        /* backward: |h v| = hvp_rel_step (1 + |x|) in the central differences; 0 = 5e-7 */
} davo_training_desc;

/* Forward.  Outputs as davo_solve_calibration (reason_out may also be DAVO_REASON_DROPPED).  The trajectory the
 * backward pass needs (create_graph=True, :85,133-135): traj_x[B,capacity,n] = iterate at which line search k
 * started, traj_g[B,capacity,n] = its gradient, traj_alpha[B,capacity] = step length applied (0 where
 * return_second_last withheld the step), traj_len[B] = recorded steps.  All four NULL: nothing is recorded. */
int davo_solve_training(const davo_problem_desc* desc, const davo_training_desc* train, const void* data0,
                        const void* data1, const void* weights, const void* x0, void* x_out, void* cost_out,
                        uint8_t* converged_out, int32_t* iters_out, int32_t* fevals_out, int32_t* reason_out,
                        void* traj_x, void* traj_g, void* traj_alpha, int32_t* traj_len, void* workspace,
                        void* stream);

/* Backward: grad_x0[B,n] = d loss / d x0 given grad_out[B,n] = d loss / d x_out, i.e. what torch.autograd computes
 * through the reference's unrolled iteration (eq. 6.17 update, eq. 6.20 scale, InverseCurvature.backward of
 * utils/func_inverse_curvature.py:22-37, alpha constant, Hessian-vector products of the objective).
 * desc->dtype must be DAVO_F64 (every buffer float64).  Problem b's recorded steps are rows traj_offset[b] ..
 * traj_offset[b] + traj_len[b] - 1 of traj_x[rows,n], traj_g[rows,n], traj_alpha[rows]: traj_offset[b] = b * capacity
 * reads the forward's buffers as they are, an exclusive prefix sum of traj_len reads a compacted copy.
 * scratch_offset[B] = exclusive prefix sum of traj_len (int64); scratch = sum(traj_len) * (n*n + n) float64 values
 * (the replayed inverse Hessians and search directions).
 * grad_data (may be NULL): d loss / d observations.  DISTORT10: [B,N,2], with respect to the (u*, v*) of the staged
 * matches; JOINT: [B,V,N,2] (data1); ANGLE_BA: [B,V,N,2] (data0); analytic models: DAVO_ERR_UNSUPPORTED when not NULL. */
int davo_solve_backward(const davo_problem_desc* desc, const davo_training_desc* train, const void* data0,
                        const void* data1, const void* weights, const void* traj_x, const void* traj_g,
                        const void* traj_alpha, const int32_t* traj_len, const int64_t* traj_offset,
                        const int64_t* scratch_offset, void* scratch, const void* grad_out, void* grad_x0,
                        void* grad_data, void* workspace, void* stream);

/* One evaluation of the objective and its gradient for every problem: cost[B], grad[B,n]
 * (grad may be NULL).  Replaces `error_function(...)` + torch.autograd.grad
 * (bfgs_solver.py:131-135) == find_error / find_error_gradient
 * (solvers/least_squares_utils.py:16-48) on the camera-model residuals. */
int davo_eval_cost_grad(const davo_problem_desc* desc, const void* data0, const void* data1,
                        const void* weights, const void* x, void* cost, void* grad, void* stream);

/* Replaces line_search_wolfe_conditions (wolfe_conditions.py:23-239) for every problem:
 * alpha_out[B]; fevals_out[B] (probes made, may be NULL). */
int davo_line_search(const davo_problem_desc* desc, const void* data0, const void* data1,
                     const void* weights, const void* x, const void* direction,
                     const void* base_cost, const void* base_grad, void* alpha_out,
                     int32_t* fevals_out, void* stream);

/* Stage matches for DISTORT10: rotate/translate points_3d[B,N,3] by the fixed pose
 * pose[B,6]=(rx,ry,rz,tx,ty,tz) (NULL = identity), divide by z' (z'==0 -> +1e-8,
 * distorted_camera_model.py:57) and interleave with obs[B,N,2] into staged[B,N,4]. */
int davo_stage_matches(const davo_problem_desc* desc, const void* points_3d, const void* obs,
                       const void* pose, void* staged, void* stream);

/* compute_distorted_camera_model (distorted_camera_model.py:106-111):
 * points_3d[B,N,3], params16[B,16] -> u[B,N], v[B,N]. */
int davo_project(const davo_problem_desc* desc, const void* points_3d, const void* params16,
                 void* u, void* v, void* stream);

/* compute_distorted_camera_model_and_jacobian (distorted_camera_model.py:114-385):
 * additionally J[B,2N,16], rows 0..N-1 = du'/dtheta, N..2N-1 = dv'/dtheta, derived from the
 * forward model (the reference's hand-written columns fx,s,fy,tx,ty,rx,ry,rz are wrong; see
 * DESIGN.md). */
int davo_project_jacobian(const davo_problem_desc* desc, const void* points_3d,
                          const void* params16, void* J, void* u, void* v, void* stream);

/* find_error / find_error_gradient (solvers/least_squares_utils.py:16-48) on explicit
 * residuals[B,R] and jacobian[B,R,P] with optional per-residual weights[B,R]:
 * error[B] = sum w r^2, gradient[B,P] = sum 2 w r J.  jacobian/gradient may be NULL. */
int davo_least_squares(int32_t dtype, int32_t B, int32_t R, int32_t P, const void* residuals,
                       const void* jacobian, const void* weights, void* error, void* gradient,
                       void* stream);

/* BFGSSolver.update_inverse_hessian (bfgs_solver.py:235-303) in place on H[k,n,n] with
 * s[k,n], y[k,n]; curvature y.s <= 0 leaves H untouched (utils/func_inverse_curvature.py:8-11). */
int davo_bfgs_update(int32_t dtype, int32_t k, int32_t n, void* H, const void* s, const void* y,
                     void* stream);

/* BFGSSolver.scale_initial_inverse_hessian (bfgs_solver.py:217-233): scale[k]. */
int davo_bfgs_initial_scale(int32_t dtype, int32_t k, int32_t n, const void* s, const void* y,
                            void* scale, void* stream);

/* interpolate_alpha (utils/func_interpolate_alpha.py:5-40), elementwise over k values: the zero of the line through
 * (alpha_1, value_1), (alpha_2, value_2), or the midpoint when the values are equal or the zero falls within 1e-3 of
 * (or outside) the interval.  out[k].  The backward entry point returns the four input gradients of the reference's
 * custom backward (:42-79) given grad_out[k]; any of the four outputs may be NULL. */
int davo_interpolate_alpha(int32_t dtype, int64_t k, const void* alpha_1, const void* alpha_2, const void* value_1,
                           const void* value_2, void* out, void* stream);
int davo_interpolate_alpha_backward(int32_t dtype, int64_t k, const void* alpha_1, const void* alpha_2,
                                    const void* value_1, const void* value_2, const void* grad_out,
                                    void* grad_alpha_1, void* grad_alpha_2, void* grad_value_1, void* grad_value_2,
                                    void* stream);

/* ---- the initial-guess network, fused (networks/calibration_network.py:35-43) ------------------------------------
 * Linear(in, H) - GELU - BatchNorm1d(H) - Linear(H, H) - GELU - BatchNorm1d(H) - Linear(H, P) in inference form, one
 * kernel on the tcgen05 tensor cores (TF32 with operand splitting: float32 accuracy), writing the solver's start
 * parameters x0[B, P].  float32 only; in_features % 8 == 0, hidden % 16 == 0, hidden <= 256, P <= 256.
 * Each Linear's weight [N, K] (row major, as torch stores it) is first re-laid by davo_mlp_pack_weights into a buffer
 * of davo_mlp_packed_bytes(N, K) bytes (split into TF32 high and low parts in the tensor cores' shared-memory
 * layout; repack when the weights change).  scale/shift fold BatchNorm1d's running statistics and affine
 * parameters: scale = weight / sqrt(running_var + eps), shift = bias - running_mean * scale. */
typedef struct davo_mlp_desc {
    int32_t B;             /* rows                                     */
    int32_t in_features;   /* 2 * views * points                       */
    int32_t hidden;        /* H                                        */
    int32_t out_features;  /* P = 3 + 3 N + 6 (M - 1)                  */
} davo_mlp_desc;
int64_t davo_mlp_packed_bytes(int32_t N, int32_t K);
int davo_mlp_pack_weights(int32_t N, int32_t K, const void* weight, void* packed, void* stream);
int davo_mlp_forward(const davo_mlp_desc* desc, const void* x, const void* w1_packed, const void* b1,
                     const void* scale1, const void* shift1, const void* w2_packed, const void* b2,
                     const void* scale2, const void* shift2, const void* w3_packed, const void* b3, void* x0_out,
                     void* stream);

/* ---- synthetic oracle-match generator, on the device -------------------------------------------------
 * Replaces the reference's host-side dataset (data/camera_and_parameters_dataset.py:48-61,85-151, batch layout
 * base_types/camera_views_and_points.py:21-33; that file does not parse at HEAD) and this repo's numpy generators
 * for the BASELINE configurations, so that million-problem batches are produced in HBM and never cross PCIe.
 * Counter-based randomness (Philox4x32-10): every value is a pure function of (seed, first_problem + row,
 * element), so shards generated by different ranks are the rows of one global batch. */
typedef struct davo_generator_desc {
    int32_t B;               /* rows to generate                                                     */
    int32_t N;               /* matches (points) per problem per view                                */
    int32_t V;               /* views (JOINT: poses per problem; views-and-points: M >= 2)           */
    int32_t dtype;           /* DAVO_F32 | DAVO_F64 (arithmetic is float64, rounded at the end)      */
    int32_t ill_conditioned; /* 1 = BASELINE config 4: heavy distortion, start focal x U_log(0.3, 3) */
    int32_t random_pose;     /* DISTORT10: 1 = draw a fixed pose per problem (else identity)         */
    uint64_t seed;
    uint64_t first_problem;  /* global index of row 0                                                */
    double fov;              /* half-width of x,y relative to z (0.5; config 4: 1.0)                 */
    double noise;            /* sigma of Gaussian noise added to the observations (0 = exact)        */
    double pathological;     /* config 4: fraction of rows with points at z -> 0+ or an ascent start */
    double start_noise;      /* views-and-points: scale of the perturbation truth -> x0 (1.0)        */
    double min_camera_distance; /* views-and-points: 0.1 in the reference                            */
} davo_generator_desc;

/* BASELINE configs 2, 4, 5: points_3d[B,N,3], obs[B,N,2], pose[B,6] (may be NULL), x0[B,10], truth[B,10] (may be NULL). */
int davo_generate_distort10(const davo_generator_desc* desc, void* points_3d, void* obs, void* pose, void* x0,
                            void* truth, void* stream);
/* BASELINE config 3: points_3d[B,N,3], obs[B,V,N,2], x0[B,10+6V], truth[B,10+6V] (may be NULL). */
int davo_generate_joint(const davo_generator_desc* desc, void* points_3d, void* obs, void* x0, void* truth,
                        void* stream);
/* The entry script's batches, CameraViewsAndPoints layout: projected_points[B,M,N,2], visibility_mask[B,M,N]
 * (1.0 / 0.0), camera_intrinsics[B,3] = (f', cx, cy), camera_orientations[B,M-1,3] (axis-angle),
 * camera_translations[B,M-1,3], world_points[B,N,3]; optionally the solver's parameter vectors
 * x0[B,n] / truth[B,n], n = 3+3N+6(M-1), laid out as unpack_calibration_parameters expects
 * (camera_model/calibration_pinhole_camera_model.py:33-75).  M = desc->V <= 32. */
int davo_generate_views_and_points(const davo_generator_desc* desc, void* projected_points, void* visibility_mask,
                                   void* camera_intrinsics, void* camera_orientations, void* camera_translations,
                                   void* world_points, void* x0, void* truth, void* stream);

/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t davo_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DAVO_B200_H */
