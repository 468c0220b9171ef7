// Latency microbenchmarks behind DESIGN.md's solver notes: dependent DFMA chain, shuffle + DFMA chain and
// the one-warp substitution of LinSys::warp_subst on a 45 x 45 factor.   nvcc -arch=sm_100a -O3 lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>
#define HMPC_NO_API
#include "../../hopper_mpc_inertial_b200/csrc/hmpc_qp.cuh"

__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 1024;
}
__global__ void k_shfl_dfma(double* out, long long* cyc, double a) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { double t = __shfl_sync(0xffffffffu, x, (i + u) & 31); x = fma(t, a, x); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[1] = (t1 - t0) / 1024;
}
__global__ void k_lds_dfma(double* out, long long* cyc) {
    __shared__ double s[64];
    s[threadIdx.x] = out[threadIdx.x]; s[threadIdx.x + 32] = 0.5;
    __syncwarp();
    double x = out[threadIdx.x];
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) { x = fma(s[idx], 0.5, x); idx = ((int)x) & 63; }   // load address depends on the chain
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 1024;
}
// the product's substitution, nk = 45, one warp alone / as one of 4 warps of a CTA with the others at the barrier
__global__ void k_subst(double* out, long long* cyc, int nk, int reps) {
    extern __shared__ double sm[];
    double* Lm = sm; double* dinv = sm + 2700; double* b = dinv + 80; double* o = b + 80; double* sc = o + 80;
    for (int i = threadIdx.x; i < 2700; i += blockDim.x) Lm[i] = 1e-3 * ((i * 7) % 13);
    for (int i = threadIdx.x; i < 80; i += blockDim.x) { dinv[i] = 1.0; b[i] = 1.0 + i; }
    hmpc::LinSys<double> sys;
    sys.n = 60; sys.nF = nk; sys.ng = 0; sys.Lm = Lm; sys.dinv = dinv; sys.H = nullptr; sys.idx = nullptr; sys.grow = nullptr;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) sys.solve(b, o, sc);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[3] = (t1 - t0) / reps;
    out[64 + threadIdx.x] = o[threadIdx.x % nk];
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMemset(out, 0, 4096); cudaMallocManaged(&cyc, 64);
    k_dfma<<<1, 32>>>(out, cyc, 0.999, 1e-3);
    k_shfl_dfma<<<1, 32>>>(out, cyc, 1e-3);
    k_lds_dfma<<<1, 32>>>(out, cyc);
    cudaDeviceSynchronize();
    printf("dependent DFMA: %lld cycles; SHFL+DFMA: %lld cycles; LDS(dep addr)+DFMA+cvt: %lld cycles\n", cyc[0], cyc[1], cyc[2]);
    cudaFuncSetAttribute(k_subst, cudaFuncAttributeMaxDynamicSharedMemorySize, 56160);
    for (int nk : {45, 60}) {
        k_subst<<<1, 128, 56160>>>(out, cyc, nk, 50); cudaDeviceSynchronize();
        printf("solve nk=%d, 1 CTA alone on the GPU: %lld cycles per solve (%lld per column step)\n", nk, cyc[3], cyc[3] / (2 * nk));
        k_subst<<<148 * 4, 128, 56160>>>(out, cyc, nk, 50); cudaDeviceSynchronize();
        printf("solve nk=%d, 4 CTAs per SM all solving: %lld cycles per solve (%lld per column step)\n", nk, cyc[3], cyc[3] / (2 * nk));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
