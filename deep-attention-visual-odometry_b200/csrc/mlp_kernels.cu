// mlp_kernels.cu — the initial-guess network of CalibrationNetwork (networks/calibration_network.py:35-43),
//   Linear(in, H) - GELU - BatchNorm1d(H) - Linear(H, H) - GELU - BatchNorm1d(H) - Linear(H, P)
// in inference form (BatchNorm folded into a per-feature scale and shift), as ONE kernel that writes the solver's
// start parameters x0[B, P].  This is the one dense contraction on the path (64 -> 256 -> 256
// -> 45 at 64K rows = 12 GFLOP), so it runs on the 5th-generation tensor cores:
//
//   * one CTA = one tile of 128 batch rows, 256 threads: threads t and t + 128 own row t (TMEM lane t) in every
//     epilogue and split its columns (a warp may only touch the TMEM lane quarter (warp % 4));
//   * tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = H (or P padded to 16), K = 8 per instruction, issued by one
//     thread, accumulators in TMEM (512 columns: two H-wide accumulators that alternate between the layers);
//   * full float32 accuracy from the TF32 pipe by operand splitting: x = hi + lo with hi = tf32(x) (round to nearest)
//     and lo = x - hi (exact), D += A_hi B_hi + A_hi B_lo + A_lo B_hi; the dropped A_lo B_lo term and the truncation
//     of lo are ~2^-21 relative, so the result agrees with a float32 GEMM to ~1e-6 (tests: <= 1e-5 against torch);
//   * operands in shared memory in the canonical K-major no-swizzle UMMA layout, one 16-byte column of 4 values at a
//     time: element (row r, 16-byte column c) at ((c R + r) 16) bytes, i.e. core matrices of 8 rows x 16 B are
//     contiguous, stride-byte-offset 128 B, leading-byte-offset 16 R;
//   * weights are split and re-laid once (mlp_pack_weights_kernel) into exactly that byte image per 64-wide K chunk,
//     so a chunk is fetched by plain 1-D bulk TMA copies (cp.async.bulk, L2 resident: 0.7 MB in all);
//   * activations never leave the SM: the epilogue of a layer reads its accumulator from TMEM (tcgen05.ld 32x32b),
//     applies bias, exact GELU and the folded BatchNorm, splits and writes the next layer's A operand to shared memory.
//
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor, UMMA::InstrDescriptor) of the
// CUTLASS headers vendored in this image; the PTX forms follow cute/arch/mma_sm100_umma.hpp (SM100_MMA_TF32_SS),
// copy_sm100.hpp (SM100_TMEM_LOAD_32dp32b32x), tmem_allocator_sm100.hpp and cutlass/arch/barrier.h (umma_arrive).
#include "davo_common.cuh"
#include "launch.h"

namespace davo {

constexpr int kMlpTile = 128;      // batch rows per tile (= TMEM lanes)
constexpr int kMlpThreads = 256;   // two threads per row: warps w and w + 4 share TMEM lane quarter w and split the columns
constexpr int kMlpChunk = 64;      // K values per staged chunk (16 columns of 16 bytes)
constexpr int kMlpMaxN = 256;      // widest layer output one MMA covers
constexpr int kMlpTmemCols = 512;

__host__ __device__ inline int mlp_pad_n(int n) { return (n + 15) & ~15; }
// bytes of one packed Linear [N, K]: per K chunk the hi image then the lo image, each cols16 x Np x 16 B
__host__ __device__ inline long long mlp_packed_bytes(int N, int K) {
    return 2LL * (K / 4) * mlp_pad_n(N) * 16;
}

// ---- PTX wrappers --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(to_tf32(x));
    lo = x - hi;
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] B[smem]^T, M = 128, K = 8 tf32
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
// 32 consecutive accumulator columns of this thread's row (TMEM lane = 32 (warp % 4) + lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA::SmemDescriptor, K-major, SWIZZLE_NONE: start address, leading byte offset (between the two 16-byte columns of
// one K = 8 step) and stride byte offset (between 8-row groups), all in 16-byte units; version 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// UMMA::InstrDescriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at bit 17,
// M >> 4 at bit 24.
__host__ __device__ inline uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- weight packing ------------------------------------------------------------------------------------------
// W [N, K] row major (torch.nn.Linear.weight) -> per K chunk of 64 (the last may be shorter): hi image then lo image,
// each [cols16][Np] float4 with rows N .. Np-1 zero.
__global__ void mlp_pack_weights_kernel(int N, int K, const float* __restrict__ W, float4* __restrict__ out) {
    const int Np = mlp_pad_n(N);
    const int cols_total = K / 4;
    const long long total = (long long)cols_total * Np;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i / Np), r = (int)(i % Np);
        const int chunk = col / (kMlpChunk / 4), c_in = col % (kMlpChunk / 4);
        const int cols_here = min(kMlpChunk / 4, cols_total - chunk * (kMlpChunk / 4));
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < N) w = *reinterpret_cast<const float4*>(W + (size_t)r * K + 4 * col);
        float4 hi, lo;
        split_tf32(w.x, hi.x, lo.x); split_tf32(w.y, hi.y, lo.y);
        split_tf32(w.z, hi.z, lo.z); split_tf32(w.w, hi.w, lo.w);
        // chunk base: all earlier chunks are full (16 columns), each 2 images of 16 Np float4
        const size_t base = (size_t)chunk * 2 * (kMlpChunk / 4) * Np;
        out[base + (size_t)c_in * Np + r] = hi;
        out[base + (size_t)cols_here * Np + (size_t)c_in * Np + r] = lo;
    }
}

struct MlpParams {
    int B, K1, H, P;
    const float* x;
    const float4* w1; const float* b1; const float* s1; const float* t1;
    const float4* w2; const float* b2; const float* s2; const float* t2;
    const float4* w3; const float* b3;
    float* out;
};

// One pipeline STEP = 32 K values (8 columns of 16 bytes) of one layer.  Shared memory holds two A buffers (hi and lo
// image, 8 x 128 x 16 B each: 32 KB per buffer) and two W buffers (8 x 256 x 16 B per image: 64 KB per buffer): 192 KB.
constexpr int kMlpStepCols = 8;
constexpr size_t kMlpWBuf = 2ull * kMlpStepCols * kMlpMaxN * 16;
constexpr size_t kMlpABuf = 2ull * kMlpStepCols * kMlpTile * 16;
constexpr size_t kMlpSmem = 2 * kMlpWBuf + 2 * kMlpABuf + 64;

// Software pipeline over the steps of a tile (2 + 8 + 8 at 64-256-256-45): while the tensor core runs step s - 1,
// the threads write step s's A images (the previous layer's epilogue) into the other A buffer and the bulk copy of
// step s's weight images lands in the other W buffer.  A buffer is reused two steps later, after the commit of the
// MMAs that read it has arrived on that buffer's mbarrier; the first step of a layer also waits for the previous
// layer's last MMAs (it reads their accumulator).
__global__ void __launch_bounds__(kMlpThreads, 1) mlp_forward_kernel(const MlpParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* Wbase = smem;
    unsigned char* Abase = smem + 2 * kMlpWBuf;
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + 2 * kMlpWBuf + 2 * kMlpABuf);   // [2]
    uint64_t* bar_mma = bar_w + 2;                                                       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 4);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & (kMlpTile - 1), part = tid / kMlpTile;   // row of the tile, which half of a step's columns
    if (tid == 0) {
        mbar_init(bar_w, 1); mbar_init(bar_w + 1, 1);
        mbar_init(bar_mma, 1); mbar_init(bar_mma + 1, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, kMlpTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    unsigned pw[2] = {0u, 0u}, pm[2] = {0u, 0u};   // barrier parities, tracked identically by every thread
    bool pending[2] = {false, false};              // a commit on bar_mma[buf] has not been waited for yet
    const int Hp = mlp_pad_n(p.H), Pp = mlp_pad_n(p.P);
    const int cols1 = p.K1 / 4, colsH = p.H / 4;
    const int steps1 = (cols1 + kMlpStepCols - 1) / kMlpStepCols, stepsH = (colsH + kMlpStepCols - 1) / kMlpStepCols;
    const int steps = steps1 + 2 * stepsH;

    auto wait_mma = [&](int buf) {
        if (pending[buf]) {
            mbar_wait(bar_mma + buf, pm[buf]);
            pm[buf] ^= 1u;
            pending[buf] = false;
            tc_fence_after();
        }
    };
    // split one float4 of row `r` into the hi / lo images of an A buffer at 16-byte column c
    auto put_a = [&](float4* A, int c, int cols, float4 v) {
        float4 hi, lo;
        split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
        split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
        A[c * kMlpTile + r] = hi;
        A[(cols + c) * kMlpTile + r] = lo;
    };
    // epilogue of a hidden layer for the step's columns: accumulator -> bias, GELU, folded BatchNorm -> A images.
    // The row's two threads take 4 columns of 16 bytes (16 accumulator columns) each.
    auto hidden_to_a = [&](float4* A, uint32_t src_col, int j0, int cols, const float* b, const float* sc, const float* sh) {
        const int c0 = 4 * part;
        if (c0 < cols) {
            // the 16 columns' bias, scale and shift: three runs of 16 consecutive floats, 64-byte aligned
            const float4* b4 = reinterpret_cast<const float4*>(b + j0 + 4 * c0);
            const float4* s4 = reinterpret_cast<const float4*>(sc + j0 + 4 * c0);
            const float4* t4 = reinterpret_cast<const float4*>(sh + j0 + 4 * c0);
            float4 bq[4], sq[4], tq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { bq[q] = __ldg(b4 + q); sq[q] = __ldg(s4 + q); tq[q] = __ldg(t4 + q); }
            float v[16];
            tmem_ld16(tmem + lane_base + src_col + j0 + 4 * c0, v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 o;
                float* op = &o.x;
                const float* bp = &bq[q].x;
                const float* sp = &sq[q].x;
                const float* tp = &tq[q].x;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float z = v[4 * q + e] + bp[e];
                    const float gl = 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f));   // nn.GELU()
                    op[e] = fmaf(gl, sp[e], tp[e]);                                            // BatchNorm1d (eval)
                }
                put_a(A, c0 + q, cols, o);
            }
        }
    };

    const int tiles = (p.B + kMlpTile - 1) / kMlpTile;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row = (long long)tile * kMlpTile + r;
        const bool live = row < p.B;
        for (int s = 0; s < steps; ++s) {
            const int buf = s & 1;
            // ---- which layer, which columns ------------------------------------------------------------------
            int layer, ls;   // layer 0..2, step within the layer
            if (s < steps1) { layer = 0; ls = s; }
            else if (s < steps1 + stepsH) { layer = 1; ls = s - steps1; }
            else { layer = 2; ls = s - steps1 - stepsH; }
            const int cols_total = layer == 0 ? cols1 : colsH;
            const int c_first = ls * kMlpStepCols;
            const int cols = min(kMlpStepCols, cols_total - c_first);
            const int Np = layer == 2 ? Pp : Hp;
            const float4* wl = layer == 0 ? p.w1 : (layer == 1 ? p.w2 : p.w3);
            // packed layout: per 64-wide chunk [hi: cols_here x Np][lo: cols_here x Np] float4 (mlp_pack_weights_kernel)
            const int chunk = c_first / (kMlpChunk / 4), c_in = c_first % (kMlpChunk / 4);
            const int cols_here = min(kMlpChunk / 4, cols_total - chunk * (kMlpChunk / 4));
            const float4* w_hi = wl + (size_t)chunk * 2 * (kMlpChunk / 4) * Np + (size_t)c_in * Np;
            const float4* w_lo = w_hi + (size_t)cols_here * Np;
            const uint32_t d_col = layer == 1 ? (uint32_t)kMlpMaxN : 0u;
            float4* A = reinterpret_cast<float4*>(Abase + buf * kMlpABuf);
            unsigned char* W = Wbase + buf * kMlpWBuf;
            const uint32_t img_bytes = (uint32_t)cols * Np * 16;
            // ---- the buffers of step s - 2 are free once its MMAs have completed -------------------------------
            wait_mma(buf);
            if (ls == 0 && layer > 0) wait_mma(buf ^ 1);   // the previous layer's accumulator is complete
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(bar_w + buf, 2 * img_bytes);
                tma_load_1d(W, w_hi, img_bytes, bar_w + buf);
                tma_load_1d(W + img_bytes, w_lo, img_bytes, bar_w + buf);
            }
            // ---- this step's A images -----------------------------------------------------------------------------
            if (layer == 0) {
                const float4* xr = reinterpret_cast<const float4*>(p.x + (size_t)(live ? row : 0) * p.K1) + c_first;
                for (int c = part; c < cols; c += 2) put_a(A, c, cols, live ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f));
            } else if (layer == 1) {
                hidden_to_a(A, 0u, 4 * c_first, cols, p.b1, p.s1, p.t1);
            } else {
                hidden_to_a(A, (uint32_t)kMlpMaxN, 4 * c_first, cols, p.b2, p.s2, p.t2);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            // ---- one thread issues the step's MMAs (3 per K = 8: lo hi, hi lo, hi hi) and commits ----------------
            if (tid == 0) {
                mbar_wait(bar_w + buf, pw[buf]);
                tc_fence_after();
                const uint32_t idesc = umma_idesc_tf32(kMlpTile, Np);
                const uint32_t a_hi = smem_u32(A), a_lo = a_hi + (uint32_t)cols * kMlpTile * 16;
                const uint32_t b_hi = smem_u32(W), b_lo = b_hi + img_bytes;
                const uint32_t a_lbo = kMlpTile * 16, b_lbo = (uint32_t)Np * 16;
                for (int k = 0; k < cols / 2; ++k) {
                    const uint64_t ah = umma_desc(a_hi + 2 * k * a_lbo, a_lbo, 128), al = umma_desc(a_lo + 2 * k * a_lbo, a_lbo, 128);
                    const uint64_t bh = umma_desc(b_hi + 2 * k * b_lbo, b_lbo, 128), bl = umma_desc(b_lo + 2 * k * b_lbo, b_lbo, 128);
                    mma_tf32(tmem + d_col, al, bh, idesc, (ls == 0 && k == 0) ? 0u : 1u);
                    mma_tf32(tmem + d_col, ah, bl, idesc, 1u);
                    mma_tf32(tmem + d_col, ah, bh, idesc, 1u);
                }
                tc_commit(bar_mma + buf);
            }
            pw[buf] ^= 1u;
            pending[buf] = true;
        }
        wait_mma(0);
        wait_mma(1);
        // ---- output: accumulator 0 + bias -> x0[row, :] -----------------------------------------------------
        for (int j0 = 16 * part; j0 < Pp; j0 += 32) {
            float v[16];
            tmem_ld16(tmem + lane_base + j0, v);
            if (live) {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (j0 + e < p.P) p.out[(size_t)row * p.P + j0 + e] = v[e] + __ldg(p.b3 + j0 + e);
            }
        }
        tc_fence_before();   // the next tile's first MMA overwrites accumulator 0: order our reads before it
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kMlpTmemCols);
}

long long launch_mlp_packed_bytes(int N, int K) { return mlp_packed_bytes(N, K); }

int launch_mlp_pack_weights(int N, int K, const float* W, void* packed, cudaStream_t s) {
    if (N < 1 || K < 8 || K % 8 != 0 || mlp_pad_n(N) > kMlpMaxN) return DAVO_ERR_UNSUPPORTED;
    const long long total = (long long)(K / 4) * mlp_pad_n(N);
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    if (blocks > 148 * 8) blocks = 148 * 8;
    mlp_pack_weights_kernel<<<(unsigned)blocks, threads, 0, s>>>(N, K, W, static_cast<float4*>(packed));
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}

int launch_mlp_forward(int B, int K1, int H, int P, const float* x, const void* w1, const float* b1, const float* s1,
                       const float* t1, const void* w2, const float* b2, const float* s2, const float* t2,
                       const void* w3, const float* b3, float* out, cudaStream_t s) {
    if (K1 < 8 || K1 % 8 != 0 || H < 16 || H % 16 != 0 || H > kMlpMaxN || P < 1 || mlp_pad_n(P) > kMlpMaxN)
        return DAVO_ERR_UNSUPPORTED;
    if (B == 0) return DAVO_OK;
    MlpParams p{B, K1, H, P, x, static_cast<const float4*>(w1), b1, s1, t1, static_cast<const float4*>(w2), b2, s2, t2,
                static_cast<const float4*>(w3), b3, out};
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return DAVO_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (!ensure_dynamic_smem(reinterpret_cast<const void*>(mlp_forward_kernel), kMlpSmem)) return DAVO_ERR_CUDA;
    const int tiles = (B + kMlpTile - 1) / kMlpTile;
    const int grid = tiles < sms ? tiles : sms;   // one CTA per SM (192 KB of shared memory, all 512 TMEM columns)
    mlp_forward_kernel<<<grid, kMlpThreads, kMlpSmem, s>>>(p);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? DAVO_OK : DAVO_ERR_CUDA;
}

}  // namespace davo
