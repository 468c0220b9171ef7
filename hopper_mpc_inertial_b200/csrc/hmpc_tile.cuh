// hmpc_tile.cuh -- FP64 tensor-core tile primitives shared by the warp-per-hopper kernel (hmpc_warp.cuh) and the
// tiled multi-warp linear algebra of the wide CTA kernels (hmpc_qp.cuh: LinSys::factor_tiled / solve_tiled).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace hmpc {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ double fast_rsqrt(double a) {
#ifdef HMPC_HOST_EMUL
    return 1.0 / sqrt(a);
#else
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    // two Newton steps: r <- r (1.5 - 0.5 a r^2)
    const double h = 0.5 * a;
    r = fma(r, fma(-h * r, r, 0.5), r);
    r = fma(r, fma(-h * r, r, 0.5), r);
    return r;
#endif
}

// ------------------------------------------------------------------------------------------------
// FP64 tensor-core building blocks.  The kernel is bound by instruction issue, not by FP64 throughput
// (profiles/README.md): one DMMA.8x8x4 does the work of eight warp-wide DFMAs at the full FP64 rate
// (tools/micro/dmma.cu: 17.5 cycles per DMMA per SM sub-partition = 14.6 FMA/clk, 26 cycles latency), so the
// dense linear algebra -- factorisation, substitutions, Hessian products -- is organised in 8x8 tiles.
//
// mma.sync.m8n8k4.f64 fragments, lane = 4 g + t:  A[g][t],  B[t][g],  C/D[g][2t], [g][2t+1].
// tile_mac(acc, a, b) adds X Y^T for two row-major 8x8 tiles X, Y when lane (g, t) passes a = X[g][2t..2t+1] and
// b = Y[g][2t..2t+1]: the two DMMAs sum over the even and the odd columns (the summation index may be permuted
// freely).  With this order an accumulator fragment (c0, c1) IS the A fragment of the next product, so chained
// products need no data movement, and every operand is one conflict-free 16-byte load per lane.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxTiles = 7;                       // compact KKT systems of (padded) order <= 56
struct d2 { double x, y; };
__device__ __forceinline__ d2 ld2(const double* p) {
#ifdef HMPC_HOST_EMUL
    return d2{p[0], p[1]};
#else
    const double2 v = *reinterpret_cast<const double2*>(p);
    return d2{v.x, v.y};
#endif
}
__device__ __forceinline__ void st2(double* p, double x, double y) {
#ifdef HMPC_HOST_EMUL
    p[0] = x; p[1] = y;
#else
    *reinterpret_cast<double2*>(p) = make_double2(x, y);
#endif
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
#ifdef HMPC_HOST_EMUL
    hmpc_emul_dmma(&c0, &c1, a, b);
#else
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
#endif
}
__device__ __forceinline__ void tile_mac(double& c0, double& c1, const d2& a, const d2& b) {
    dmma(c0, c1, a.x, b.x);
    dmma(c0, c1, a.y, b.y);
}
// lower block triangle of 8x8 row-major tiles: tile (I, J), J <= I
__host__ __device__ __forceinline__ int tile_off(int I, int J) { return ((I * (I + 1)) / 2 + J) * 64; }
// a vector piece v[0..8) held as "lane 4 g has v[g]" (column 0 of an accumulator) -> B fragment of the 8x1 operand:
// lanes 0..3 get (v[2t], v[2t+1]), every other lane zero
__device__ __forceinline__ d2 vec_to_b(double v, int lane) {
    const int t = lane & 3;
    const double x = __shfl_sync(kFullMask, v, 8 * t), y = __shfl_sync(kFullMask, v, 8 * t + 4);
    return (lane < 4) ? d2{x, y} : d2{0.0, 0.0};
}


// B fragment of an 8x1 operand v[0..8) read from shared memory: lanes 0..3 get (v[2t], v[2t+1]), all others zero
__device__ __forceinline__ d2 vec_b(const double* v, int lane) {
    return (lane < 4) ? ld2(v + 2 * lane) : d2{0.0, 0.0};
}

// ------------------------------------------------------------------------------------------------
// 8x8 diagonal block of the system, in place in registers: signed Cholesky  C = V S V'  of the tile D (row-major,
// lower triangle; the strict upper triangle is never read as data), then the inverse  W = V^-1  (lower triangular,
// zeros above the diagonal) into Wt.  Lane l holds entries (l >> 3, l & 7) and (4 + (l >> 3), l & 7).
//   MIXED = false: all eight pivots have the same sign sgn (+1 variables, -1 active rows; a block of rows is the
//                  plain Cholesky factorisation of -C);
//   MIXED = true : pivots 0 .. jrel-1 are variables (+), jrel .. 7 active rows (-), 0 < jrel < 8, and the diagonal of
//                  the remaining (Schur complement) rows is scaled by 1 + eps once the variables are eliminated
//                  (hmpc_qp.cuh: LinSys::factor).
// The eight elimination steps are unrolled: which register holds column j / row j is then known at compile time, and
// one step costs five 64-bit shuffles, one reciprocal square root and a handful of FMAs (profiles/README.md: the
// rolled loop spent 1300 instructions per block, a third of the kernel's instruction stream).
// Returns nonzero when a pivot has the wrong sign or is not finite.  Ends with a __syncwarp().
// ------------------------------------------------------------------------------------------------
template <bool MIXED>
__device__ __forceinline__ int wdiag8(const double* D, double* Wt, int jrel, double sgn, double eps, int lane) {
    const int kk = lane & 7, i0 = lane >> 3, i1 = i0 + 4;
    const int rowsrc = lane & 24, colsrc = 8 * (kk & 3);   // lanes holding (i0, 0) / (kk mod 4, 0): + j gives column j
    const bool lowk = kk < 4;
    int bad = 0;
    // W starts as the identity and receives the row operations of the elimination (forward substitution on I):
    // after pivot j, row j of W is final and rows i > j have  W(i, 0..j) -= V(i, j) W(j, 0..j)
    double w0 = (i0 == kk) ? 1.0 : 0.0, w1 = (i1 == kk) ? 1.0 : 0.0;     // W(i0, kk), W(i1, kk)
    double c0 = D[8 * i0 + kk], c1 = D[8 * i1 + kk];                      // C(i0, kk), C(i1, kk)
    if (!MIXED) { c0 *= sgn; c1 *= sgn; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        double s = 1.0;
        if (MIXED) {
            if (j == jrel) {            // all variables eliminated: regularise the rows' Schur complement
                if (i0 == kk && kk >= j) c0 *= 1.0 + eps;
                if (i1 == kk && kk >= j) c1 *= 1.0 + eps;
            }
            s = (j < jrel) ? 1.0 : -1.0;
        }
        // pivot C(j, j): lane 8 (j mod 4) + j, register c0 for j < 4 and c1 otherwise
        const double piv = __shfl_sync(kFullMask, j < 4 ? c0 : c1, 8 * (j & 3) + j);
        const double ap = MIXED ? s * piv : piv;
        const bool ok = (ap > 0.0) && (ap < 1e30);
        if (!ok) bad = 1;
        const double rsq = fast_rsqrt(ok ? ap : 1.0);
        const double srs = MIXED ? s * rsq : rsq;
        // V(i, j) = S_j C(i, j) / sqrt(|C(j, j)|) for this lane's two rows and for its column
        const double vi0 = __shfl_sync(kFullMask, c0, rowsrc + j) * srs, vi1 = __shfl_sync(kFullMask, c1, rowsrc + j) * srs;
        const double ck0 = __shfl_sync(kFullMask, c0, colsrc + j), ck1 = __shfl_sync(kFullMask, c1, colsrc + j);
        const double vk = (lowk ? ck0 : ck1) * (MIXED ? rsq : srs);     // S_j V(kk, j)
        // row j of W, scaled: W(j, kk) / V(j, j)
        const double wj = __shfl_sync(kFullMask, j < 4 ? w0 : w1, 8 * (j & 3) + kk) * rsq;
        if (kk > j) {                               // rank-1 update of the trailing part (upper entries: don't care)
            c0 = fma(-vi0, vk, c0);
            c1 = fma(-vi1, vk, c1);
        } else {
            w0 = (i0 == j) ? wj : (i0 > j ? fma(-vi0, wj, w0) : w0);
            w1 = (i1 == j) ? wj : (i1 > j ? fma(-vi1, wj, w1) : w1);
        }
    }
    Wt[8 * i0 + kk] = w0;
    Wt[8 * i1 + kk] = w1;
    __syncwarp();
    return bad;
}


}  // namespace hmpc
