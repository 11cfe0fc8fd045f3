// Microbenchmark behind DESIGN.md's index section: sustained FP64 rate of a B200 SM for dependent-chain-free DFMA
// (vector pipe) and mma.sync.m8n8k4.f64 (DMMA).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma_kernel(double *out, int iters, double a, double b) {
    double d0[ILP], d1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) d0[i] = d1[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0[i]), "+d"(d1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d0[i] + d1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = warps * 32;
        float ms = time_ms([&] { dfma_kernel<16><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
        double fma = (double)sms * threads * 16.0 * iters;
        printf("DFMA  %2d warps/SM: %.3f ms  %.2f TFLOPS  (%.1f FMA/clk/SM at %d MHz)\n", warps, ms, 2 * fma / ms / 1e9, fma / sms / (ms * 1e-3) / (p.clockRate * 1e3), p.clockRate / 1000);
        ms = time_ms([&] { dmma_kernel<8><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
        fma = (double)sms * warps * 8.0 * iters * 256.0;
        printf("DMMA  %2d warps/SM: %.3f ms  %.2f TFLOPS  (%.1f FMA/clk/SM)\n", warps, ms, 2 * fma / ms / 1e9, fma / sms / (ms * 1e-3) / (p.clockRate * 1e3));
    }
    return 0;
}
