// fp32_peak.cu — measured FP32 FMA peak of this GPU (the denominator of bench.py's roofline).
// Every thread runs 16 independent FFMA chains (register-resident, no memory traffic); 2 flop per FFMA per lane.
// Prints one JSON line.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp32_peak tools/fp32_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <algorithm>

#define CHAINS 16
#define INNER 512

__global__ void __launch_bounds__(256) ffma_kernel(float* out, float a, float b, int outer) {
    float x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int k = 0; k < INNER / CHAINS; ++k) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) x[i] = fmaf(x[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i];
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;     // never true: keeps the chains alive
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clock_khz = 0; cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, 1 << 20);
    const int blocks = p.multiProcessorCount * 8, threads = 256, outer = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) ffma_kernel<<<blocks, threads>>>(out, 0.999f, 1e-4f, outer);
    cudaDeviceSynchronize();
    double best = 0, sum = 0; const int reps = 10;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        ffma_kernel<<<blocks, threads>>>(out, 0.999f, 1e-4f, outer);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * (double)blocks * threads * (double)outer * INNER;
        const double tf = flops / (ms * 1e-3) / 1e12;
        best = std::max(best, tf); sum += tf;
    }
    cudaError_t err = cudaGetLastError();
    const double nominal = (double)p.multiProcessorCount * 128 * 2 * clock_khz * 1e3 / 1e12;
    printf("{\"fp32_tflops_measured\": %.2f, \"fp32_tflops_mean\": %.2f, \"sms\": %d, \"clock_mhz_attr\": %.0f, \"nominal_at_attr_clock\": %.2f, "
           "\"how\": \"16 independent FFMA chains per thread, %d blocks x %d threads, best of %d (CUDA events)\", \"gpu\": \"%s\", \"error\": \"%s\"}\n",
           best, sum / reps, p.multiProcessorCount, clock_khz / 1e3, nominal, blocks, threads, reps, p.name, cudaGetErrorString(err));
    return err == cudaSuccess ? 0 : 1;
}
