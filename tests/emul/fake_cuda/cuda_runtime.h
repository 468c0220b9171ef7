// Stand-in for <cuda_runtime.h> used ONLY by tests/emul: lets g++ compile the product's device
// functions (csrc/*.cuh) as ordinary host code executed by one serial "thread".
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#define HMPC_HOST_EMUL 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(x)
struct hmpc_emul_dim { int x; };
static const hmpc_emul_dim threadIdx{0}, blockDim{1}, blockIdx{0}, gridDim{1};
static inline void __syncthreads() {}
static inline int __syncthreads_or(int v) { return v; }
static inline void __syncwarp() {}
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline int min(int a, int b) { return a < b ? a : b; }
