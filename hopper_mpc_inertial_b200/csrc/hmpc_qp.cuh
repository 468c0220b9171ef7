// hmpc_qp.cuh -- per-hopper MPC QP on the device: linearise, condense, ADMM + verified polish.
//
// One CTA owns one hopper's QP.  All vectors live in shared memory; the condensed Hessian H and the
// factor of the KKT operator live in shared memory when they fit, else in a per-CTA slice of a global
// workspace that stays L2-resident (persistent grid).
//
// What is computed (reference file:line into the reference's src/):
//   linearise   gen_dt_dynamics          mpc_cvx_euler_3f.py:71-94 / mpc_cvx_euler_2f.py:70-94
//   condense    build_qp                 mpc_cvx_euler_3f.py:96-153 / 2f:96-151  
//   solve       cp.Problem(...).solve(solver=cp.OSQP)   mpc_cvx_euler_3f.py:155-160
//
// Condensed form (inputs only, n = 6N):  X = c + S U,  H = 2(S'QS + R),  g = 2(S'Q(c - xref) - R ubar).
// Because A_k A_j = 0 S has closed-form blocks
//     p-rows  dt (i-a-1) Bv_a     theta-rows  dt W_{a+1,i} Bw_a     v-rows  Bv_a     w-rows  Bw_a
// with W_{a+1,i} = sum_{l=a+1}^{i-1} Rz_l, so H is assembled block-by-block without forming S.
// Constraint rows use a fixed slot layout (m = 11N):  [6N identity box | 4N friction | N height].
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace hmpc {

constexpr double kInf = 1e30;
constexpr double kInfThresh = 1e26;   // OSQP: OSQP_INFTY * MIN_SCALING
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6;

struct QpConst {
    int N, dyn, uref_mode, mode, max_iter, check, polish, adaptive_rho;
    double dt, m, g, mu;
    double Jinv[9], rh[3], tau_max[3];
    double fz_max, z_min, kf;
    double eps_abs, eps_rel, rho0, sigma, alpha, delta, polish_tol;
};

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up (all doubles unless noted)
// ------------------------------------------------------------------------------------------------
struct Work {
    // linearisation
    double *gp;      // [N][4]  guess position + yaw per stage
    double *cz, *sz; // [N]
    double *PC, *PS; // [N+1]   prefix sums of cos/sin
    double *Bv;      // [N][9]  dt * B[6:9, 0:3]
    double *Bw;      // [N][18] dt * B[9:12, 0:6]
    double *pfw;     // [N][3]
    double *xin;     // [12]
    double *cfree;   // [N+1][12] free response
    double *err;     // [N+1][12] cfree[i] - xref[i-1]  (row 0 unused)
    double *Qd, *Rd; // [12], [6]
    // QP vectors
    double *g, *lo, *hi, *rv;          // [n] [m] [m] [m]
    double *x, *xt, *rhs, *z, *y, *wv; // n n n m m m
    double *xp, *mul, *wp, *bnd, *tmp; // n m m m n   (polish: point, multipliers, penalties, bounds)
    double *ts, *sc;                   // [m] scratch rows, [n] solve scratch
    double *dinv;                      // [n]
    double *red;                       // [64] reduction scratch
    int *fixed;                        // [n]  1: variable eliminated a priori (lo == hi)
    int *pin;                          // [n]  polish: variable pinned at a bound
    int *code;                         // [m]  polish: +1 upper active, -1 lower active, 0 inactive
    int *stance;                       // [N]
    // matrices (shared or global)
    double *H, *LC, *LR;
};

__host__ __device__ inline size_t work_vec_doubles(int N) {
    const int n = 6 * N, m = 11 * N;
    size_t d = 0;
    d += 4 * N + 2 * N + 2 * (N + 1) + 9 * N + 18 * N + 3 * N + 12 + 2 * 12 * (N + 1) + 12 + 6;
    d += n + 3 * m;           // g lo hi rv
    d += 3 * n + 3 * m;       // x xt rhs z y wv
    d += n + 3 * m + n;       // xp mul wp bnd tmp
    d += m + n;               // ts sc
    d += n;                   // dinv
    d += 64;                  // red
    d += (2 * n + m + N + 1) / 2 + 1; // fixed, pin, code, stance (ints)
    return d;
}

__device__ inline void carve(Work& w, double* base, int N) {
    const int n = 6 * N, m = 11 * N;
    double* p = base;
    auto take = [&](size_t k) { double* r = p; p += k; return r; };
    w.gp = take(4 * N); w.cz = take(N); w.sz = take(N); w.PC = take(N + 1); w.PS = take(N + 1);
    w.Bv = take(9 * N); w.Bw = take(18 * N); w.pfw = take(3 * N); w.xin = take(12);
    w.cfree = take(12 * (N + 1)); w.err = take(12 * (N + 1)); w.Qd = take(12); w.Rd = take(6);
    w.g = take(n); w.lo = take(m); w.hi = take(m); w.rv = take(m);
    w.x = take(n); w.xt = take(n); w.rhs = take(n); w.z = take(m); w.y = take(m); w.wv = take(m);
    w.xp = take(n); w.mul = take(m); w.wp = take(m); w.bnd = take(m); w.tmp = take(n);
    w.ts = take(m); w.sc = take(n);
    w.dinv = take(n); w.red = take(64);
    w.fixed = reinterpret_cast<int*>(p);
    w.pin = w.fixed + n;
    w.code = w.pin + n;
    w.stance = w.code + m;
}

// ------------------------------------------------------------------------------------------------
// block reductions (max) -- warp shuffles + one shared round
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// reduces K values at once; result valid in all threads. NaN-propagating via flag in slot K.
template <int K>
__device__ inline void block_max(double (&v)[K], double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_max(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[wid * K + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double r = red[k];
        for (int q = 1; q < nw; ++q) r = fmax(r, red[q * K + k]);
        v[k] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// constraint operator A (implicit)
//   rows 0..n-1        identity
//   rows n+4k+s        friction, stance stages only:  s=0: fx-mu fz, 1: -fx-mu fz, 2: fy-mu fz, 3: -fy-mu fz
//                      (s>=2 disabled for 2f)
//   rows n+4N+k        height z_k (k>=2):  sum_{j<=k-2} dt^2 (k-j-1)/m * fz_j
// ------------------------------------------------------------------------------------------------
struct AOp {
    int N, n, fy_rows;   // fy_rows: 1 for 3f
    double mu, zc;       // zc = dt^2/m
    const int* stance;
    __device__ __forceinline__ bool fr_on(int k, int s) const { return stance[k] && (s < 2 || fy_rows); }
    __device__ inline double row(int r, const double* x) const {
        if (r < n) return x[r];
        r -= n;
        if (r < 4 * N) {
            const int k = r >> 2, s = r & 3;
            if (!fr_on(k, s)) return 0.0;
            const double sg = (s & 1) ? -1.0 : 1.0;
            return sg * x[6 * k + (s >> 1)] - mu * x[6 * k + 2];
        }
        const int k = r - 4 * N;
        double acc = 0.0;
        for (int j = 0; j + 2 <= k; ++j) acc += zc * (double)(k - j - 1) * x[6 * j + 2];
        return acc;
    }
    // (A^T v)_i
    __device__ inline double colT(int i, const double* v) const {
        double acc = v[i];
        const int k = i / 6, c = i - 6 * k;
        if (c > 2) return acc;
        const double* f = v + n + 4 * k;
        if (stance[k]) {
            if (c == 0) acc += f[0] - f[1];
            else if (c == 1) { if (fy_rows) acc += f[2] - f[3]; }
            else acc -= mu * (f[0] + f[1] + (fy_rows ? f[2] + f[3] : 0.0));
        }
        if (c == 2) {
            const double* zr = v + n + 4 * N;
            for (int kk = k + 2; kk < N; ++kk) acc += zc * (double)(kk - k - 1) * zr[kk];
        }
        return acc;
    }
    // (A^T diag(w) A)_{ij} without the identity rows
    __device__ inline double gram(int i, int j, const double* w) const {
        const int ki = i / 6, ci = i - 6 * ki, kj = j / 6, cj = j - 6 * kj;
        if (ci > 2 || cj > 2) return 0.0;
        double acc = 0.0;
        if (ki == kj && stance[ki]) {
            const double* f = w + n + 4 * ki;
            const double w0 = f[0], w1 = f[1], w2 = fy_rows ? f[2] : 0.0, w3 = fy_rows ? f[3] : 0.0;
            const int lo_ = ci < cj ? ci : cj, hi_ = ci < cj ? cj : ci;
            if (lo_ == 0 && hi_ == 0) acc += w0 + w1;
            else if (lo_ == 1 && hi_ == 1) acc += w2 + w3;
            else if (lo_ == 2 && hi_ == 2) acc += mu * mu * (w0 + w1 + w2 + w3);
            else if (lo_ == 0 && hi_ == 2) acc += -mu * (w0 - w1);
            else if (lo_ == 1 && hi_ == 2) acc += -mu * (w2 - w3);
        }
        if (ci == 2 && cj == 2) {
            const double* zr = w + n + 4 * N;
            const int k0 = (ki > kj ? ki : kj) + 2;
            for (int kk = k0; kk < N; ++kk)
                acc += zr[kk] * zc * zc * (double)(kk - ki - 1) * (double)(kk - kj - 1);
        }
        return acc;
    }
};

// ------------------------------------------------------------------------------------------------
// linearisation of one stage (gen_dt_dynamics); thread k handles stage k
// ------------------------------------------------------------------------------------------------
__device__ inline void linearize_stage(const QpConst& c, int k, Work& w) {
    const double psi = w.gp[4 * k + 3];
    double sn, cs;
    sincos(psi, &sn, &cs);
    w.cz[k] = cs; w.sz[k] = sn;
    const double Rz[9] = {cs, sn, 0, -sn, cs, 0, 0, 0, 1};
    const double d[3] = {w.pfw[3 * k] - w.gp[4 * k], w.pfw[3 * k + 1] - w.gp[4 * k + 1],
                         w.pfw[3 * k + 2] - w.gp[4 * k + 2]};
    double rf[3];
    mat3_vec(Rz, d, rf);
    rf[0] += c.rh[0]; rf[1] += c.rh[1]; rf[2] += c.rh[2];
    // Jw = Rz Jinv Rz^T
    double T1[9], Jw[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0;
            for (int l = 0; l < 3; ++l) a += Rz[3 * i + l] * c.Jinv[3 * l + j];
            T1[3 * i + j] = a;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0;
            for (int l = 0; l < 3; ++l) a += T1[3 * i + l] * Rz[3 * j + l];
            Jw[3 * i + j] = a;
        }
    // JwRzT = Jw Rz^T
    double JwRzT[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0;
            for (int l = 0; l < 3; ++l) a += Jw[3 * i + l] * Rz[3 * j + l];
            JwRzT[3 * i + j] = a;
        }
    double Bf[9];   // B[9:12, 0:3]
    double* Bv = w.Bv + 9 * k;
    if (c.dyn == 3) {
        double rw[3];
        mat3T_vec(Rz, rf, rw);   // Rz^T rf
        const double hatm[9] = {0, -rw[2], rw[1], rw[2], 0, -rw[0], -rw[1], rw[0], 0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double a = 0;
                for (int l = 0; l < 3; ++l) a += Jw[3 * i + l] * hatm[3 * l + j];
                Bf[3 * i + j] = a;
            }
        for (int i = 0; i < 9; ++i) Bv[i] = 0.0;
        Bv[0] = Bv[4] = Bv[8] = (1.0 / c.m) * c.dt;
    } else {
        const double hatm[9] = {0, -rf[2], rf[1], rf[2], 0, -rf[0], -rf[1], rf[0], 0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double a = 0;
                for (int l = 0; l < 3; ++l) a += JwRzT[3 * i + l] * hatm[3 * l + j];
                Bf[3 * i + j] = a;
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Bv[3 * i + j] = (Rz[3 * j + i] / c.m) * c.dt;   // Rz^T / m
    }
    double* Bw = w.Bw + 18 * k;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            Bw[6 * i + j] = Bf[3 * i + j] * c.dt;
            Bw[6 * i + 3 + j] = JwRzT[3 * i + j] * c.dt;
        }
}

// ------------------------------------------------------------------------------------------------
// Condense.  Requires w.gp, w.pfw, w.xin, w.Qd, w.Rd, w.stance filled and a __syncthreads() before.
// xref points at this hopper's window, element (k, c) at xref[(k*12 + c) * xs].
// Leaves H (row-major n x n, symmetric, full), g, lo, hi, fixed, cfree in place.  Returns (all threads)
// 1 when the u-independent height rows k=0,1 are violated.
// ------------------------------------------------------------------------------------------------
__device__ inline int condense(const QpConst& c, Work& w, const double* xref, size_t xs) {
    const int N = c.N, n = 6 * N, m = 11 * N, tid = threadIdx.x, T = blockDim.x;
    for (int k = tid; k < N; k += T) linearize_stage(c, k, w);
    __syncthreads();
    if (tid == 0) {
        double pc = 0, ps = 0;
        w.PC[0] = 0; w.PS[0] = 0;
        for (int k = 0; k < N; ++k) { pc += w.cz[k]; ps += w.sz[k]; w.PC[k + 1] = pc; w.PS[k + 1] = ps; }
    }
    __syncthreads();
    const double dt = c.dt, gdt = -c.g * dt;
    // free response + tracking error
    for (int i = tid; i <= N; i += T) {
        double* cf = w.cfree + 12 * i;
        const double* x0 = w.xin;
        const double di = (double)i;
        cf[6] = x0[6]; cf[7] = x0[7]; cf[8] = x0[8] + di * gdt;
        cf[9] = x0[9]; cf[10] = x0[10]; cf[11] = x0[11];
        cf[0] = x0[0] + dt * (di * x0[6]);
        cf[1] = x0[1] + dt * (di * x0[7]);
        cf[2] = x0[2] + dt * (di * x0[8] + gdt * (0.5 * di * (di - 1.0)));
        const double pc = w.PC[i], ps = w.PS[i];
        cf[3] = x0[3] + dt * (pc * x0[9] + ps * x0[10]);
        cf[4] = x0[4] + dt * (-ps * x0[9] + pc * x0[10]);
        cf[5] = x0[5] + dt * (di * x0[11]);
        if (i >= 1)
            for (int q = 0; q < 12; ++q) w.err[12 * i + q] = cf[q] - xref[((size_t)(i - 1) * 12 + q) * xs];
    }
    // bounds
    for (int r = tid; r < m; r += T) {
        double lo = -kInf, hi = kInf;
        if (r < n) {
            const int k = r / 6, cc = r - 6 * k;
            if (cc >= 3) { lo = -c.tau_max[cc - 3]; hi = c.tau_max[cc - 3]; }
            else if (!w.stance[k]) { lo = 0.0; hi = 0.0; }
            else if (cc == 2) { lo = 0.0; hi = c.fz_max; }
            if (cc == 1 && c.dyn == 2) { lo = 0.0; hi = 0.0; }
            w.fixed[r] = (hi - lo) < 1e-12;
        } else if (r < n + 4 * N) {
            const int k = (r - n) >> 2, s = (r - n) & 3;
            if (w.stance[k] && (s < 2 || c.dyn == 3)) hi = 0.0;
        }
        w.lo[r] = lo; w.hi[r] = hi;
    }
    __syncthreads();
    // height rows need cfree
    int infeasible = (w.cfree[2] < c.z_min) || (N >= 1 && w.cfree[12 + 2] < c.z_min);
    for (int k = 2 + tid; k < N; k += T) w.lo[n + 4 * N + k] = c.z_min - w.cfree[12 * k + 2];
    // Hessian blocks, lower block-triangle a >= b
    const int nblk = N * (N + 1) / 2;
    const double q3 = w.Qd[3], q4 = w.Qd[4], q5 = w.Qd[5];
    for (int p = tid; p < nblk; p += T) {
        // decode p -> (a, b), a >= b
        int a = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
        while ((a + 1) * (a + 2) / 2 <= p) ++a;
        while (a * (a + 1) / 2 > p) --a;
        const int b = p - a * (a + 1) / 2;
        const double pca = w.PC[a + 1], psa = w.PS[a + 1], pcb = w.PC[b + 1], psb = w.PS[b + 1];
        double s0 = 0, NN = 0, CC = 0, SS = 0, CS = 0, SC = 0;
        for (int i = a + 1; i <= N; ++i) {
            const double kap = (i == N) ? c.kf : 1.0;
            const double wca = w.PC[i] - pca, wsa = w.PS[i] - psa, wcb = w.PC[i] - pcb, wsb = w.PS[i] - psb;
            const double na = (double)(i - a - 1), nb = (double)(i - b - 1);
            s0 += kap; NN += kap * na * nb;
            CC += kap * wca * wcb; SS += kap * wsa * wsb; CS += kap * wca * wsb; SC += kap * wsa * wcb;
        }
        const double dt2 = dt * dt;
        double M3[9] = {s0 * w.Qd[9] + dt2 * (q3 * CC + q4 * SS), dt2 * (q3 * CS - q4 * SC), 0,
                        dt2 * (q3 * SC - q4 * CS), s0 * w.Qd[10] + dt2 * (q3 * SS + q4 * CC), 0,
                        0, 0, s0 * w.Qd[11] + dt2 * q5 * NN};
        const double dv[3] = {dt2 * NN * w.Qd[0] + s0 * w.Qd[6], dt2 * NN * w.Qd[1] + s0 * w.Qd[7],
                              dt2 * NN * w.Qd[2] + s0 * w.Qd[8]};
        const double* Bwa = w.Bw + 18 * a; const double* Bwb = w.Bw + 18 * b;
        const double* Bva = w.Bv + 9 * a;  const double* Bvb = w.Bv + 9 * b;
        double MB[18];   // M3 * Bw_b  (3x6)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 6; ++j)
                MB[6 * i + j] = M3[3 * i] * Bwb[j] + M3[3 * i + 1] * Bwb[6 + j] + M3[3 * i + 2] * Bwb[12 + j];
        for (int r = 0; r < 6; ++r)
            for (int cc = 0; cc < 6; ++cc) {
                double v = Bwa[r] * MB[cc] + Bwa[6 + r] * MB[6 + cc] + Bwa[12 + r] * MB[12 + cc];
                if (r < 3 && cc < 3)
                    v += Bva[r] * dv[0] * Bvb[cc] + Bva[3 + r] * dv[1] * Bvb[3 + cc] + Bva[6 + r] * dv[2] * Bvb[6 + cc];
                v *= 2.0;
                if (a == b && r == cc && a != N - 1) v += 2.0 * w.Rd[r];
                w.H[(size_t)(6 * a + r) * n + 6 * b + cc] = v;
                w.H[(size_t)(6 * b + cc) * n + 6 * a + r] = v;
            }
    }
    // gradient
    const double ubar_alias = (c.uref_mode == 0) ? (w.stance[N - 1] ? 2.0 * c.m * c.g : 0.0) : 0.0;
    for (int a = tid; a < N; a += T) {
        double ap[3] = {0, 0, 0}, av[3] = {0, 0, 0}, aw[3] = {0, 0, 0}, at[3] = {0, 0, 0};
        const double pca = w.PC[a + 1], psa = w.PS[a + 1];
        for (int i = a + 1; i <= N; ++i) {
            const double kap = (i == N) ? c.kf : 1.0;
            const double* e = w.err + 12 * i;
            const double na = (double)(i - a - 1);
            const double wca = w.PC[i] - pca, wsa = w.PS[i] - psa;
            for (int q = 0; q < 3; ++q) { ap[q] += kap * na * e[q]; av[q] += kap * e[6 + q]; aw[q] += kap * e[9 + q]; }
            at[0] += kap * (wca * q3 * e[3] - wsa * q4 * e[4]);
            at[1] += kap * (wsa * q3 * e[3] + wca * q4 * e[4]);
            at[2] += kap * (na * q5 * e[5]);
        }
        double tv[3], tw[3];
        for (int q = 0; q < 3; ++q) {
            tv[q] = dt * w.Qd[q] * ap[q] + w.Qd[6 + q] * av[q];
            tw[q] = dt * at[q] + w.Qd[9 + q] * aw[q];
        }
        const double* Bwa = w.Bw + 18 * a; const double* Bva = w.Bv + 9 * a;
        for (int r = 0; r < 6; ++r) {
            double v = Bwa[r] * tw[0] + Bwa[6 + r] * tw[1] + Bwa[12 + r] * tw[2];
            if (r < 3) v += Bva[r] * tv[0] + Bva[3 + r] * tv[1] + Bva[6 + r] * tv[2];
            v *= 2.0;
            if (r == 2 && a != N - 1) {
                const double ub = (c.uref_mode == 0) ? ubar_alias : (w.stance[a] ? 2.0 * c.m * c.g : 0.0);
                v -= 2.0 * w.Rd[2] * ub;
            }
            w.g[6 * a + r] = v;
        }
    }
    __syncthreads();
    return infeasible;
}

// ------------------------------------------------------------------------------------------------
// Generic linear-system policy: dense Cholesky of
//     K = H + dadd I + diag(wbox) + A_g' diag(w_g) A_g       with eliminated variables as identity rows
// stored column-major in LC and row-major in LR (so both substitutions stream contiguous memory).
// ------------------------------------------------------------------------------------------------
struct CholSys {
    int n;
    double *LC, *LR, *dinv;
    const double* H;
    // wts: [m] row weights (identity rows first).  fixed: variables pinned -> identity row/col.
    __device__ inline int factor(const AOp& A, const double* wts, double dadd, const int* fixed, double* red) {
        const int tid = threadIdx.x, T = blockDim.x;
        int bad = 0;
        for (int j = 0; j < n; ++j) {
            const bool fj = fixed[j] != 0;
            for (int i = j + tid; i < n; i += T) {
                double s;
                if (fj || fixed[i]) s = (i == j) ? 1.0 : 0.0;
                else {
                    s = H[(size_t)j * n + i] + A.gram(i, j, wts);
                    if (i == j) s += dadd + wts[i];
                    const double* li = LC + i; const double* lj = LC + j;
                    for (int k = 0; k < j; ++k) s -= li[(size_t)k * n] * lj[(size_t)k * n];
                }
                LC[(size_t)j * n + i] = s;
                if (i == j) red[0] = s;
            }
            __syncthreads();
            const double piv = red[0];
            if (!(piv > 0.0)) bad = 1;
            const double inv = rsqrt(piv > 0.0 ? piv : 1.0);
            for (int i = j + tid; i < n; i += T) {
                const double v = LC[(size_t)j * n + i] * inv;
                LC[(size_t)j * n + i] = v;
                LR[(size_t)i * n + j] = v;
                if (i == j) dinv[j] = 1.0 / v;
            }
            __syncthreads();
        }
        return bad;
    }
    // solves K out = b; b is destroyed; out may not alias b.  Ends with a __syncthreads().
    __device__ inline void solve(double* b, double* out, double* scratch) {
        const int tid = threadIdx.x, T = blockDim.x;
        __syncthreads();
        for (int j = 0; j < n; ++j) {
            const double wj = b[j] * dinv[j];
            const double* col = LC + (size_t)j * n;
            for (int i = j + 1 + tid; i < n; i += T) b[i] -= col[i] * wj;
            if (tid == 0) scratch[j] = wj;
            __syncthreads();
        }
        for (int j = n - 1; j >= 0; --j) {
            const double xj = scratch[j] * dinv[j];
            const double* row = LR + (size_t)j * n;
            for (int i = tid; i < j; i += T) scratch[i] -= row[i] * xj;
            if (tid == 0) out[j] = xj;
            __syncthreads();
        }
    }
};

__device__ inline void sym_matvec(const double* H, int n, const double* x, double* out) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc += H[(size_t)j * n + i] * x[j];
        out[i] = acc;
    }
}

struct SolveInfo { int status, iters, nfac, npolish; double rho; };

// ------------------------------------------------------------------------------------------------
// Verified active-set polish.  Guess the active set from the ADMM iterate (z, y) with OSQP's rule,
// pin box-active variables exactly (identity rows), put a 1/delta penalty on active friction/height
// rows and run a few method-of-multipliers steps in correction form (each step is also a step of
// iterative refinement).  The result is ACCEPTED only if it passes the KKT conditions of the original
// QP: stationarity on free variables, feasibility of every row, equality on active rows and the sign
// of every multiplier.  On success x <- solution, y <- multipliers and 1 is returned (all threads).
// The system factor is overwritten either way.
// ------------------------------------------------------------------------------------------------
template <class Sys>
__device__ inline int polish_verified(const QpConst& c, Work& w, Sys& sys, const AOp& A, int& nfac) {
    const int N = c.N, n = 6 * N, m = 11 * N, tid = threadIdx.x, T = blockDim.x;
    int ngen_loc = 0;
    for (int r = tid; r < m; r += T) {
        const double z = w.z[r], y = w.y[r], lo = w.lo[r], hi = w.hi[r];
        const bool low = (z - lo) < -y, upp = (hi - z) < y;
        const int code = low ? -1 : (upp ? 1 : 0);
        w.code[r] = code;
        w.bnd[r] = low ? lo : hi;
        w.mul[r] = 0.0;
        if (r < n) {
            const int pin = (w.fixed[r] || code != 0) ? 1 : 0;
            w.pin[r] = pin;
            w.wp[r] = 0.0;
            w.xp[r] = pin ? (w.fixed[r] ? lo : (low ? lo : hi)) : w.x[r];
        } else {
            w.wp[r] = code ? 1.0 / c.delta : 0.0;
            ngen_loc += (code != 0);
        }
    }
    const int ngen = __syncthreads_or(ngen_loc);
    const double dprox = ngen ? c.delta : 0.0;
    ++nfac;
    if (sys.factor(A, w.wp, dprox, w.pin, w.red)) return 0;
    const int kmom = ngen ? 8 : 2;
    for (int k = 0; k < kmom; ++k) {
        // Newton step on the augmented Lagrangian at xp (exact for a quadratic; step 2+ also refines)
        sym_matvec(w.H, n, w.xp, w.tmp);
        for (int r = tid; r < m; r += T)
            w.ts[r] = (r >= n && w.code[r]) ? w.mul[r] + w.wp[r] * (A.row(r, w.xp) - w.bnd[r]) : 0.0;
        __syncthreads();
        for (int i = tid; i < n; i += T)
            w.rhs[i] = w.pin[i] ? 0.0 : -(w.tmp[i] + w.g[i] + A.colT(i, w.ts));
        sys.solve(w.rhs, w.xt, w.sc);
        for (int i = tid; i < n; i += T) w.xp[i] += w.xt[i];
        __syncthreads();
        for (int r = n + tid; r < m; r += T)
            if (w.code[r]) w.mul[r] += w.wp[r] * (A.row(r, w.xp) - w.bnd[r]);
        __syncthreads();
    }
    sym_matvec(w.H, n, w.xp, w.tmp);
    for (int r = tid; r < m; r += T) w.ts[r] = (r >= n && w.code[r]) ? w.mul[r] : 0.0;
    __syncthreads();
    // ---- KKT verification on the ORIGINAL problem (ts = multipliers of general rows, tmp = H xp) ----
    double v[5] = {0, 0, 0, 0, 0};   // stat, scale, feas, sign, |mult|
    for (int i = tid; i < n; i += T) {
        const double G = w.tmp[i] + w.g[i] + A.colT(i, w.ts);   // ts[i] == 0 on box rows
        v[1] = fmax(v[1], fmax(fabs(w.tmp[i]), fabs(w.g[i])));
        if (!w.pin[i]) v[0] = fmax(v[0], fabs(G));
        else {
            const double lam = -G;   // box multiplier from stationarity
            w.mul[i] = lam;
            v[4] = fmax(v[4], fabs(lam));
            if (!w.fixed[i]) v[3] = fmax(v[3], w.code[i] > 0 ? -lam : lam);
        }
    }
    for (int r = tid; r < m; r += T) {
        const double ax = A.row(r, w.xp), lo = w.lo[r], hi = w.hi[r];
        double f = fmax(lo - ax, ax - hi) / (1.0 + fmin(fabs(lo), fabs(hi)));
        if (r >= n && w.code[r]) {
            f = fmax(f, fabs(ax - w.bnd[r]) / (1.0 + fabs(w.bnd[r])));
            const double lam = w.mul[r];
            v[4] = fmax(v[4], fabs(lam));
            v[3] = fmax(v[3], w.code[r] > 0 ? -lam : lam);
        }
        v[2] = fmax(v[2], f);
    }
    block_max<5>(v, w.red);
    const double scale = fmax(1.0, v[1]);
    const double tol = c.polish_tol;
    const bool ok = (v[0] <= tol * scale) && (v[2] <= tol) && (v[3] <= tol * fmax(scale, v[4])) &&
                    (v[0] == v[0]) && (v[2] == v[2]) && (v[3] == v[3]);
    if (ok) {
        for (int i = tid; i < n; i += T) { w.x[i] = w.xp[i]; w.y[i] = w.pin[i] ? w.mul[i] : 0.0; }
        for (int r = n + tid; r < m; r += T) w.y[r] = w.code[r] ? w.mul[r] : 0.0;
        __syncthreads();
        for (int r = tid; r < m; r += T) w.z[r] = A.row(r, w.x);
        __syncthreads();
    }
    return ok ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// ADMM (OSQP iteration, SURVEY App. C2, dense-friendly form) + verified polish.
// On entry: H, g, lo, hi, fixed set; x, y hold the warm start (or zeros).  On exit x = solution,
// y = multipliers.  Variables with lo == hi (swing forces, 2f's fy) are eliminated exactly.
// ------------------------------------------------------------------------------------------------
template <class Sys>
__device__ inline SolveInfo admm_solve(const QpConst& c, Work& w, Sys& sys, const AOp& A) {
    const int N = c.N, n = 6 * N, m = 11 * N, tid = threadIdx.x, T = blockDim.x;
    SolveInfo info{1 /*HMPC_MAX_ITER*/, 0, 0, 0, c.rho0};
    double rho = c.rho0;
    const double sigma = c.sigma, alpha = c.alpha;

    auto set_rho = [&](double r) {
        for (int i = tid; i < m; i += T) {
            const double lo = w.lo[i], hi = w.hi[i];
            double v = r;
            if (lo < -kInfThresh && hi > kInfThresh) v = kRhoMin;
            else if (hi - lo < 1e-4) v = fmin(1e3 * r, kRhoMax);
            w.rv[i] = v;
        }
        __syncthreads();
    };
    set_rho(rho);
    for (int i = tid; i < n; i += T) if (w.fixed[i]) w.x[i] = w.lo[i];
    __syncthreads();
    for (int r = tid; r < m; r += T) w.z[r] = fmin(fmax(A.row(r, w.x), w.lo[r]), w.hi[r]);
    __syncthreads();
    info.nfac = 1;
    if (sys.factor(A, w.rv, sigma, w.fixed, w.red)) { info.status = 3; return info; }

    const int last_it = c.max_iter;
    bool conv = false;
    for (int it = 1; it <= last_it; ++it) {
        for (int r = tid; r < m; r += T) w.wv[r] = w.rv[r] * w.z[r] - w.y[r];
        __syncthreads();
        for (int i = tid; i < n; i += T)
            w.rhs[i] = w.fixed[i] ? w.lo[i] : (sigma * w.x[i] - w.g[i] + A.colT(i, w.wv));
        sys.solve(w.rhs, w.xt, w.sc);
        for (int r = tid; r < m; r += T) {
            const double zt = A.row(r, w.xt);
            const double zr = alpha * zt + (1.0 - alpha) * w.z[r];
            const double rv = w.rv[r];
            const double zn = fmin(fmax(zr + w.y[r] / rv, w.lo[r]), w.hi[r]);
            w.y[r] += rv * (zr - zn);
            w.z[r] = zn;
        }
        for (int i = tid; i < n; i += T) w.x[i] = alpha * w.xt[i] + (1.0 - alpha) * w.x[i];
        __syncthreads();
        info.iters = it;
        const bool do_check = (c.mode == 1) ? (it == last_it) : (it % c.check == 0 || it == last_it);
        if (!do_check) continue;

        // ---- residuals of the unscaled problem (OSQP termination test) ----
        sym_matvec(w.H, n, w.x, w.tmp);
        __syncthreads();
        double v[6] = {0, 0, 0, 0, 0, 0};   // pri, npri, dua, |Hx|, |A'y|, |g|
        for (int r = tid; r < m; r += T) {
            const double ax = A.row(r, w.x);
            v[0] = fmax(v[0], fabs(ax - w.z[r]));
            v[1] = fmax(v[1], fmax(fabs(ax), fabs(w.z[r])));
        }
        for (int i = tid; i < n; i += T) {
            if (w.fixed[i]) continue;   // eliminated variables carry an implicit multiplier
            const double aty = A.colT(i, w.y);
            v[2] = fmax(v[2], fabs(w.tmp[i] + w.g[i] + aty));
            v[3] = fmax(v[3], fabs(w.tmp[i]));
            v[4] = fmax(v[4], fabs(aty));
            v[5] = fmax(v[5], fabs(w.g[i]));
        }
        block_max<6>(v, w.red);
        const double pri = v[0], npri = v[1], dua = v[2], ndua = fmax(v[3], fmax(v[4], v[5]));
        if (!(pri == pri) || !(dua == dua)) { info.status = 3; break; }
        conv = pri <= c.eps_abs + c.eps_rel * npri && dua <= c.eps_abs + c.eps_rel * ndua;
        if (!c.polish) {
            if (conv) { info.status = 4; break; }
        } else {
            ++info.npolish;
            if (polish_verified(c, w, sys, A, info.nfac)) { info.status = 0; break; }
        }
        if (it == last_it) break;
        // ---- rho adaptation (OSQP residual balancing), and restore of the ADMM factor ----
        bool refactor = c.polish != 0;
        if (c.adaptive_rho) {
            double rn = rho * sqrt((pri / fmax(npri, 1e-10)) / fmax(dua / fmax(ndua, 1e-10), 1e-10));
            rn = fmin(fmax(rn, kRhoMin), kRhoMax);
            if (rn > 5.0 * rho || rn < 0.2 * rho) { rho = rn; set_rho(rho); refactor = true; }
        }
        if (refactor) {
            ++info.nfac;
            if (sys.factor(A, w.rv, sigma, w.fixed, w.red)) { info.status = 3; break; }
        }
    }
    if (info.status == 1 && conv) info.status = 4;
    info.rho = rho;
    return info;
}

}  // namespace hmpc
