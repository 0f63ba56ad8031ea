// ffma2_probe.cu — does packed fma.rn.f32x2 (SASS FFMA2, sm_100+) free issue slots on B200?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_probe tools/ffma2_probe.cu && ./ffma2_probe
// Four kernels, each thread runs kChains independent dependency chains for kIters rounds:
//   scalar      : kChains FFMA per round
//   packed      : kChains/2 FFMA2 per round (the same number of FMAs)
//   scalar+alu  : kChains FFMA + kAlu integer ops per round
//   packed+alu  : kChains/2 FFMA2 + kAlu integer ops per round
// Reports FMA/clk/SM for each (peak = 128).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kChains = 16;
constexpr int kIters = 4096;

template <bool kPacked, int kAlu>
__global__ void __launch_bounds__(256) probe(float* out, float a, float b, unsigned seed) {
    float acc[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc[i] = threadIdx.x * 0.001f + i;
    unsigned z[4] = {seed + threadIdx.x, seed * 3u, seed * 5u, seed * 7u};
    for (int it = 0; it < kIters; ++it) {
        if (kPacked) {
#pragma unroll
            for (int i = 0; i < kChains; i += 2) {
                float2 v = make_float2(acc[i], acc[i + 1]);
                v = __ffma2_rn(v, make_float2(a, a), make_float2(b, b));
                acc[i] = v.x;
                acc[i + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < kChains; ++i) acc[i] = fmaf(acc[i], a, b);
        }
#pragma unroll
        for (int j = 0; j < kAlu; ++j) z[j & 3] = (z[j & 3] ^ (z[(j + 1) & 3] >> 3)) + 0x9e3779b9u;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + float(z[0] ^ z[1] ^ z[2] ^ z[3]);
}

template <bool kPacked, int kAlu>
void run(const char* name, float* out, int sms, double mhz) {
    const int grid = sms * 4, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<kPacked, kAlu><<<grid, block>>>(out, 1.0001f, 0.5f, 1u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<kPacked, kAlu><<<grid, block>>>(out, 1.0001f, 0.5f, 1u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = double(grid) * block * kChains * kIters;
    const double clk = ms * 1e-3 * mhz * 1e6;
    printf("%-14s %8.3f ms  %6.1f FMA/clk/SM (at %.0f MHz nominal)  alu ops per round %d\n", name, ms, fma / clk / sms, mhz,
           kAlu);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    float* out;
    cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 4 * 256);
    run<false, 0>("scalar", out, p.multiProcessorCount, mhz);
    run<true, 0>("packed", out, p.multiProcessorCount, mhz);
    run<false, 4>("scalar+alu4", out, p.multiProcessorCount, mhz);
    run<true, 4>("packed+alu4", out, p.multiProcessorCount, mhz);
    run<false, 8>("scalar+alu8", out, p.multiProcessorCount, mhz);
    run<true, 8>("packed+alu8", out, p.multiProcessorCount, mhz);
    run<false, 16>("scalar+alu16", out, p.multiProcessorCount, mhz);
    run<true, 16>("packed+alu16", out, p.multiProcessorCount, mhz);
    return 0;
}
