// tools/micro/dmma.cu -- FP64 tensor-core (DMMA.8x8x4) throughput and latency on sm_100a vs DFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu && ./dmma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void dmma_kernel(double* out, int iters, double a, double b, long long* cyc) {
    double c0[CHAINS], c1[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c0[i] = threadIdx.x + i; c1[i] = 0.5 * i; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma(c0[i], c1[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CHAINS>
__global__ void dfma_kernel(double* out, int iters, double a, double b, long long* cyc) {
    double c[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = threadIdx.x + i;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <class F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 4096;
    long long h;
    // latency: one warp, one chain
    dmma_kernel<1><<<1, 32>>>(out, iters, 1.0000001, 0.5, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA dependent-chain latency: %.1f cycles\n", (double)h / iters);
    dfma_kernel<1><<<1, 32>>>(out, iters, 1.0000001, 0.5, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent-chain latency: %.1f cycles\n", (double)h / iters);
    // single-warp issue rate, 8 independent chains
    dmma_kernel<8><<<1, 32>>>(out, iters, 1.0000001, 0.5, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA one warp, 8 chains: %.1f cycles per DMMA\n", (double)h / iters / 8);
    dmma_kernel<8><<<1, 128>>>(out, iters, 1.0000001, 0.5, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA four warps (one per SMSP), 8 chains: %.1f cycles per DMMA per warp\n", (double)h / iters / 8);
    // full-chip throughput
    for (int warps = 4; warps <= 32; warps *= 2) {
        float ms = timeit([&] { dmma_kernel<8><<<sms, 32 * warps>>>(out, iters, 1.0000001, 0.5, cyc); });
        const double flops = 2.0 * 256 * 8.0 * iters * warps * sms;
        printf("DMMA %2d warps/SM: %.2f TFLOP/s\n", warps, flops / (ms * 1e-3) / 1e12);
    }
    for (int warps = 8; warps <= 32; warps *= 2) {
        float ms = timeit([&] { dfma_kernel<8><<<sms, 32 * warps>>>(out, iters, 1.0000001, 0.5, cyc); });
        const double flops = 2.0 * 32 * 8.0 * iters * warps * sms;
        printf("DFMA %2d warps/SM: %.2f TFLOP/s\n", warps, flops / (ms * 1e-3) / 1e12);
    }
    return 0;
}
