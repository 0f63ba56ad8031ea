// train_params.cuh — parameter blocks of the training-mode solve and its backward pass (train_kernels.cu), shared by
// the kernels and the C-ABI translation unit.
#pragma once
#include "davo_common.cuh"

namespace davo {

// ---- training-mode hooks of the solve (bfgs_solver.py:88-93 thresholds, :122-125 drop-path, :196-212
// return_second_last, and the iterates create_graph=True keeps alive for the backward pass) -----------------
// NoRecorder compiles to nothing: the eval-mode solve is unchanged.
struct NoRecorder {
    static constexpr bool kActive = false;
    __device__ __forceinline__ bool drop(int, int) const { return false; }
    __device__ __forceinline__ bool second_last() const { return false; }
    __device__ __forceinline__ bool secant() const { return false; }
    template <typename T>
    __device__ __forceinline__ void record(int, int, int, const T*, const T*, T, int) const {}
    __device__ __forceinline__ void finish(int, int, int) const {}
};

// One uniform in [0, 1) on the float32 grid torch.rand uses (multiples of 2^-24), a pure function of
// (seed, problem, iteration): Philox4x32-10 with counter (problem, 0, iteration, 'DROP').
__host__ __device__ inline uint32_t drop_path_bits(uint64_t seed, uint32_t problem, uint32_t iteration) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = problem, c1 = 0u, c2 = iteration, c3 = 0x44524F50u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return c0;
}

template <typename T>
struct TrainRecorder {
    static constexpr bool kActive = true;
    T* traj_x;          // [B, capacity, n]  iterate x_k at which line search k started (NULL: no recording)
    T* traj_g;          // [B, capacity, n]  gradient at x_k
    T* traj_alpha;      // [B, capacity]     step length applied (0 when return_second_last withheld the step)
    int32_t* traj_len;  // [B]               recorded steps
    int capacity;
    float drop_p;       // drop_path_p (0: never)
    uint64_t seed;
    int second;         // return_second_last
    int zoom;           // 1: secant zoom (davo_problem_desc.zoom_interpolation)
    __device__ __forceinline__ bool drop(int b, int k) const {   // bfgs_solver.py:122-125: keep iff rand > p
        if (!(drop_p > 0.0f)) return false;
        const float u = (float)(drop_path_bits(seed, (uint32_t)b, (uint32_t)k) >> 8) * (1.0f / 16777216.0f);
        return !(u > drop_p);
    }
    __device__ __forceinline__ bool second_last() const { return second != 0; }
    __device__ __forceinline__ bool secant() const { return zoom != 0; }
    __device__ __forceinline__ void record(int b, int k, int n, const T* x, const T* g, T alpha, int lane) const {
        if (!traj_x || k >= capacity) return;
        const size_t row = ((size_t)b * capacity + k) * n;
        for (int c = lane; c < n; c += 32) {
            traj_x[row + c] = x[c];
            traj_g[row + c] = g[c];
        }
        if (lane == 0) traj_alpha[(size_t)b * capacity + k] = alpha;
    }
    __device__ __forceinline__ void finish(int b, int steps, int lane) const {
        if (traj_len && lane == 0) traj_len[b] = steps < capacity ? steps : capacity;
    }
};

template <typename T>
struct BackwardParams {
    const T* traj_x;                // [rows, n]  problem b's steps are rows traj_offset[b] .. + traj_len[b]
    const T* traj_g;                // [rows, n]
    const T* traj_alpha;            // [rows]
    const int32_t* traj_len;        // [B]
    const int64_t* traj_offset;     // [B] (b * capacity for the forward's padded buffers, a prefix sum once compacted)
    const int64_t* scratch_offset;  // [B] exclusive prefix sum of traj_len
    T* scratch;                     // sum(traj_len) x (n*n + n): H_k then d_k per recorded step
    const T* grad_out;              // [B, n] d loss / d x_out
    T* grad_x0;                     // [B, n] d loss / d x0
    T* grad_data;                   // DISTORT10: [B, N, 2] d loss / d observations (NULL: not wanted)
    T rel_step;
};

}  // namespace davo
