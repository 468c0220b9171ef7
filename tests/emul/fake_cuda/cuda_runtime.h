// Stand-in for <cuda_runtime.h> used ONLY by tests/emul: lets g++ compile the product's device
// functions (csrc/*.cuh) as ordinary host code.
//
// Two execution models:
//   * blockDim.x == 1 (default): one serial "thread" runs the CTA-per-hopper code (round-1 model);
//   * warp mode (hmpc_emul_run_warp): 32 cooperative fibers (ucontext) execute the warp-per-hopper code in
//     lock-step at every warp-synchronous primitive.  __shfl_sync / __shfl_xor_sync / __ballot_sync /
//     __any_sync / __syncwarp are rendezvous points: a lane that does not arrive (divergent control flow
//     around a shuffle) deadlocks the emulation, which the runtime reports and aborts on.
// The emulation sees arithmetic, indexing, lane ownership and shuffle/ballot semantics; it cannot see a
// MISSING __syncwarp() (fibers only switch at the primitives), which the -m gpu tests cover.
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
#define __launch_bounds__(...)
struct hmpc_emul_dim { int x; };
extern int hmpc_emul_tid, hmpc_emul_bdim;          // defined in emul.cpp (set by the fiber scheduler)
#define threadIdx (hmpc_emul_dim{hmpc_emul_tid})
#define blockDim (hmpc_emul_dim{hmpc_emul_bdim})
static const hmpc_emul_dim blockIdx{0}, gridDim{1};
static inline void __syncthreads() {}
static inline int __syncthreads_or(int v) { return v; }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline int min(int a, int b) { return a < b ? a : b; }

// warp-synchronous primitives (emul.cpp)
void hmpc_emul_rendezvous();                       // all live lanes meet
extern double hmpc_emul_xd[32];
extern long long hmpc_emul_xi[32];
static inline void __syncwarp() { if (hmpc_emul_bdim > 1) hmpc_emul_rendezvous(); }
static inline double __shfl_sync(unsigned, double v, int src) {
    if (hmpc_emul_bdim == 1) return v;
    hmpc_emul_xd[hmpc_emul_tid] = v;
    hmpc_emul_rendezvous();
    const double r = hmpc_emul_xd[src & 31];
    hmpc_emul_rendezvous();
    return r;
}
static inline double __shfl_sync(unsigned, double v, int src, int width) {
    if (hmpc_emul_bdim == 1) return v;
    hmpc_emul_xd[hmpc_emul_tid] = v;
    hmpc_emul_rendezvous();
    const double r = hmpc_emul_xd[(hmpc_emul_tid & ~(width - 1)) | (src & (width - 1))];
    hmpc_emul_rendezvous();
    return r;
}
// mma.sync.aligned.m8n8k4.row.col.f64: lane = 4 g + t holds A[g][t], B[t][g], C/D[g][2t], [g][2t+1]
extern double hmpc_emul_ma[32], hmpc_emul_mb[32];
extern long long hmpc_emul_count[8];               // [0] DMMA lane-calls, [1..] user counters (HMPC_EMUL_COUNT)
#define HMPC_EMUL_COUNT(k) (++hmpc_emul_count[k])
static inline void hmpc_emul_dmma(double* c0, double* c1, double a, double b) {
    ++hmpc_emul_count[0];
    hmpc_emul_ma[hmpc_emul_tid] = a;
    hmpc_emul_mb[hmpc_emul_tid] = b;
    hmpc_emul_rendezvous();
    const int g = hmpc_emul_tid >> 2, t = hmpc_emul_tid & 3;
    double d0 = *c0, d1 = *c1;
    for (int k = 0; k < 4; ++k) {
        d0 = fma(hmpc_emul_ma[4 * g + k], hmpc_emul_mb[4 * (2 * t) + k], d0);
        d1 = fma(hmpc_emul_ma[4 * g + k], hmpc_emul_mb[4 * (2 * t + 1) + k], d1);
    }
    hmpc_emul_rendezvous();
    *c0 = d0; *c1 = d1;
}
static inline int __shfl_sync(unsigned, int v, int src) {
    if (hmpc_emul_bdim == 1) return v;
    hmpc_emul_xi[hmpc_emul_tid] = v;
    hmpc_emul_rendezvous();
    const int r = (int)hmpc_emul_xi[src & 31];
    hmpc_emul_rendezvous();
    return r;
}
static inline double __shfl_xor_sync(unsigned, double v, int o) {
    if (hmpc_emul_bdim == 1) return v;
    hmpc_emul_xd[hmpc_emul_tid] = v;
    hmpc_emul_rendezvous();
    const double r = hmpc_emul_xd[(hmpc_emul_tid ^ o) & 31];
    hmpc_emul_rendezvous();
    return r;
}
static inline unsigned __ballot_sync(unsigned, int p) {
    if (hmpc_emul_bdim == 1) return p ? 1u : 0u;
    hmpc_emul_xi[hmpc_emul_tid] = p ? 1 : 0;
    hmpc_emul_rendezvous();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) if (hmpc_emul_xi[l]) m |= (1u << l);
    hmpc_emul_rendezvous();
    return m;
}
static inline int __any_sync(unsigned mask, int p) { return __ballot_sync(mask, p) != 0u; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
