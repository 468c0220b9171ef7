// Prototype + micro-benchmark of a register-tiled right-looking LDL' for the <= 64-unknown systems of
// mpc_kernel (see experiments/README.md): the matrix lives in REGISTERS, distributed block-cyclically over a
// 16 x 8 thread grid (thread (tr, tc) owns rows tr + 16 a, columns tc + 8 b); per pivot the owners publish
// the pivot column through a double-buffered shared-memory vector, ONE barrier, and every thread updates its
// tile.  Compared against LinSys<double>::factor (the shipped blocked smem algorithm) on the same matrix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include tiled.cu -o tiled
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../hopper_mpc_inertial_b200/csrc/hmpc_qp.cuh"

using hmpc::tri_off;

// reciprocal without the library's slow path: MUFU seed + two Newton steps (relative error ~1e-16, not
// correctly rounded -- the factor only feeds an iteratively refined, KKT-verified solve)
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(fma(-d, r, 1.0), r, r);
    r = fma(fma(-d, r, 1.0), r, r);
    return r;
}

template <int RT, int CT, int JB>
__device__ __forceinline__ void publish(const double (&a)[RT][CT], double* buf, int tr, bool own, int k, int nk) {
    if constexpr (JB < CT) {
#pragma unroll
        for (int ia = 0; ia < RT; ++ia) {
            const int i = tr + 16 * ia;
            if (own && i >= k && i < nk) buf[i] = a[ia][JB];
        }
    }
}

template <int RT, int CT, int MODE>
__device__ __forceinline__ int factor_tiled(double* Lm, double* dinv, int nk, int nF, int ng, double eps, double* scratch) {
    const int tid = threadIdx.x, tr = tid & 15, tc = tid >> 4;
    double* buf0 = scratch;            // [64] pivot column, double-buffered
    double* buf1 = scratch + 64;
    double a[RT][CT];
#pragma unroll
    for (int jb = 0; jb < CT; ++jb) {
        const int j = tc + 8 * jb;
        const double* colp = Lm + tri_off(j < nk ? j : 0, nk) - (j < nk ? j : 0);
#pragma unroll
        for (int ia = 0; ia < RT; ++ia) {
            const int i = tr + 16 * ia;
            a[ia][jb] = (i < nk && j <= i) ? colp[i] : 0.0;
        }
    }
    int bad = 0;
    for (int k = 0; k < nk; ++k) {
        double* buf = (k & 1) ? buf1 : buf0;
        if (k == nF && ng > 0) {          // relative regularisation of the row block (see LinSys::factor)
#pragma unroll
            for (int ia = 0; ia < RT; ++ia)
#pragma unroll
                for (int jb = 0; jb < CT; ++jb) {
                    const int i = tr + 16 * ia, j = tc + 8 * jb;
                    if (i == j && i >= nF) a[ia][jb] *= 1.0 + eps;
                }
        }
        // owners publish column k (rows >= k).  The column block k >> 3 is CTA-uniform; an explicit switch keeps
        // the tile statically indexed (a compare inside an unrolled loop is turned into a dynamic index by the
        // compiler, which sends the whole tile to local memory).
        const bool own = tc == (k & 7);
        switch (k >> 3) {
            case 0: publish<RT, CT, 0>(a, buf, tr, own, k, nk); break;
            case 1: publish<RT, CT, 1>(a, buf, tr, own, k, nk); break;
            case 2: publish<RT, CT, 2>(a, buf, tr, own, k, nk); break;
            case 3: publish<RT, CT, 3>(a, buf, tr, own, k, nk); break;
            case 4: publish<RT, CT, 4>(a, buf, tr, own, k, nk); break;
            case 5: publish<RT, CT, 5>(a, buf, tr, own, k, nk); break;
            case 6: publish<RT, CT, 6>(a, buf, tr, own, k, nk); break;
            default: publish<RT, CT, 7>(a, buf, tr, own, k, nk); break;
        }
        __syncthreads();
        const double piv = buf[k];
        const double ap = (k < nF) ? piv : -piv;
        const bool ok = (ap > 0.0) && (ap < 1e30);
        if (!ok) bad = 1;
        const double rinv = (MODE & 1) ? 1.0 : fast_rcp(ok ? piv : 1.0);
        if (tid == 0) dinv[k] = rinv;
        double li[RT], uj[CT];
        double* colp = Lm + tri_off(k, nk) - k;       // the unit-lower column goes to the packed factor
#pragma unroll
        for (int ia = 0; ia < RT; ++ia) {
            const int i = tr + 16 * ia;
            const bool live = i > k && i < nk;
            li[ia] = live ? buf[live ? i : 0] * rinv : 0.0;
            if (own && live) colp[i] = li[ia];
        }
#pragma unroll
        for (int jb = 0; jb < CT; ++jb) {
            const int j = tc + 8 * jb;
            const bool live = j > k && j < nk;
            uj[jb] = live ? buf[live ? j : 0] : 0.0;
        }
        if (!(MODE & 2)) {
#pragma unroll
            for (int ia = 0; ia < RT; ++ia)
#pragma unroll
                for (int jb = 0; jb < CT; ++jb) a[ia][jb] -= li[ia] * uj[jb];
        }
    }
    return __syncthreads_or(bad);
}

__global__ void __launch_bounds__(128, 4) k_factor(const double* Hin, double* out, long long* cyc, int nk, int reps, int which) {
    extern __shared__ double sm[];
    double* H = sm;                    // packed SPD input, order nk
    double* Lm = H + 1900;             // packed factor
    double* dinv = Lm + 2700; double* scr = dinv + 80; int* idx = reinterpret_cast<int*>(scr + 320);
    const int tot = nk * (nk + 1) / 2;
    for (int i = threadIdx.x; i < tot; i += blockDim.x) H[i] = Hin[i];
    for (int i = threadIdx.x; i < nk; i += blockDim.x) idx[i] = i;
    hmpc::LinSys<double> sys;
    sys.n = nk; sys.nF = nk; sys.ng = 0; sys.Lm = Lm; sys.dinv = dinv; sys.H = H; sys.idx = idx; sys.grow = nullptr;
    hmpc::AOp A{};                      // unused: no weights, no rows
    __syncthreads();
    int bad = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (which == 0) bad |= sys.factor(A, nullptr, 0.0, 0.0, scr);
        else {
            for (int i = threadIdx.x; i < tot; i += blockDim.x) Lm[i] = H[i];     // "assembly"
            __syncthreads();
            if (which == 1) bad |= (nk <= 48) ? factor_tiled<3, 6, 0>(Lm, dinv, nk, nk, 0, 0.0, scr) : factor_tiled<4, 8, 0>(Lm, dinv, nk, nk, 0, 0.0, scr);
            else if (which == 2) bad |= factor_tiled<3, 6, 1>(Lm, dinv, nk, nk, 0, 0.0, scr);
            else if (which == 3) bad |= factor_tiled<3, 6, 3>(Lm, dinv, nk, nk, 0, 0.0, scr);
            if (which == 1) {          // diagonal entries are not stored by the tiled version; not read by the solves
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0) / reps; cyc[1] = bad; }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < tot; i += blockDim.x) out[i] = Lm[i];
        for (int i = threadIdx.x; i < nk; i += blockDim.x) out[3000 + i] = dinv[i];
    }
}

int main() {
    for (int nk : {45, 60}) {
        const int tot = nk * (nk + 1) / 2;
        std::vector<double> M(nk * nk), Hp(tot);
        // SPD test matrix: B B' + nk I
        std::vector<double> Bm(nk * nk);
        unsigned s = 12345;
        for (auto& v : Bm) { s = s * 1664525u + 1013904223u; v = ((s >> 8) & 0xffff) / 65536.0 - 0.5; }
        for (int i = 0; i < nk; ++i) for (int j = 0; j < nk; ++j) { double acc = (i == j) ? 1.0 : 0.0; for (int k = 0; k < nk; ++k) acc += Bm[i * nk + k] * Bm[j * nk + k]; M[i * nk + j] = acc; }
        for (int j = 0; j < nk; ++j) for (int i = j; i < nk; ++i) Hp[tri_off(j, nk) + i - j] = M[i * nk + j];
        double *dH, *dout; long long* cyc;
        cudaMalloc(&dH, tot * 8); cudaMemcpy(dH, Hp.data(), tot * 8, cudaMemcpyHostToDevice);
        cudaMalloc(&dout, 4096 * 8); cudaMallocManaged(&cyc, 64);
        cudaFuncSetAttribute(k_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, 56160);
        std::vector<double> ref(4096), got(4096);
        for (int which = 0; which < (nk <= 48 ? 4 : 2); ++which) {
            for (int grid : {1, 148 * 4}) {
                k_factor<<<grid, 128, 56160>>>(dH, dout, cyc, nk, 20, which); cudaDeviceSynchronize();
                printf("nk=%d %s, %s: %lld cycles per factorisation (bad=%lld) %s\n", nk, which == 0 ? "shipped blocked smem" : which == 1 ? "register-tiled" : which == 2 ? "tiled, no division (timing only)" : "tiled, no division, no update (timing only)",
                       grid == 1 ? "1 CTA alone" : "4 CTAs per SM", cyc[0], cyc[1], cudaGetErrorString(cudaGetLastError()));
            }
            if (which < 2) cudaMemcpy(which ? got.data() : ref.data(), dout, 4096 * 8, cudaMemcpyDeviceToHost);
        }
        double e = 0, ed = 0;
        for (int j = 0; j < nk; ++j) for (int i = j + 1; i < nk; ++i) e = fmax(e, fabs(ref[tri_off(j, nk) + i - j] - got[tri_off(j, nk) + i - j]));
        for (int i = 0; i < nk; ++i) ed = fmax(ed, fabs(ref[3000 + i] - got[3000 + i]) / fabs(ref[3000 + i]));
        printf("nk=%d max |L' diff| = %.3e, max rel dinv diff = %.3e\n", nk, e, ed);
    }
    return 0;
}
