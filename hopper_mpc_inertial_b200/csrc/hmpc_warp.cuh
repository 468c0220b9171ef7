// hmpc_warp.cuh -- the hot path of mpcontrol: ONE WARP per hopper, no CTA barrier anywhere.
//
// Round 1's kernel gave every hopper a 128-thread CTA; its profile (profiles/r1j_full_batch_mpc_kernel_ncu.txt)
// showed 97 k warp-instructions per hopper-tick for ~2.2 k warp-DFMAs of arithmetic, 35 % of the stall samples on
// CTA barriers and four barrier-coupled hoppers per SM.  This file re-states the warm path -- time shift,
// linearise, condense, verified primal-dual active-set refinement (hmpc_qp.cuh: polish_verified), solution
// roll-out -- for a single warp:
//   * every synchronisation is a __syncwarp(); reductions, broadcasts and compactions are shuffles / ballots;
//   * the factor of the compact KKT system (order <= kcap) is the only matrix in shared memory; the condensed
//     Hessian lives in a per-warp slice of a global workspace that stays L2-resident (compact over the non-fixed
//     variables, full square so that every access is one contiguous row segment);
//   * the factorisation is a left-looking, 4-column-blocked SIGNED CHOLESKY  K = V S V'  (S = +1 on the variable
//     columns, -1 on the active-row columns of the quasi-definite polish system): accumulators in registers, one
//     own-row load + four broadcast loads per four FMAs, no store inside the update loop, the summation index split
//     over idle lanes once fewer than 17 rows remain; K's entries are gathered on the fly (no assembly pass);
//   * substitutions keep the right-hand side in registers (rows lane, lane+32, ...) and broadcast the pivot entry
//     with one shuffle per column.
// A hopper the warm path cannot finish (no valid previous tick, infeasible height row, active set not verified
// within the retry budget, system larger than kcap) is appended to a deferral list and handled by the CTA kernel
// (hmpc_kernel.cuh: interior point + polish) in a second launch; results do not depend on which kernel ran first
// because every accepted point passes the same KKT test of the original QP.
//
// Reference behaviour implemented (file:line into the reference's src/): mpc_cvx_euler_3f.py:59-68 (time shift,
// linearise, build, solve), :71-94 (gen_dt_dynamics), :96-153 (build_qp), 2f: the same lines of mpc_cvx_euler_2f.py.
#pragma once
#include <stdlib.h>

#include "hmpc_qp.cuh"
#include "hmpc_mpc.cuh"

namespace hmpc {

constexpr unsigned kFullMask = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// per-warp shared memory
// ------------------------------------------------------------------------------------------------
struct WWork {
    // linearisation / condensing
    double *cz, *sz, *PC, *PS;   // [N] [N] [N+1] [N+1]
    double *Bw;                  // [N][18]  dt * B[9:12, 0:6]
    double *xin, *Qd, *Rd;       // [12] [12] [6]
    double *hinv, *hlo;          // [N] height-row normalisation 1/(k-1), height-row lower bounds
    double *blo, *bhi;           // [2][6] box bounds by (stance, component)
    double *err;                 // [N+1][12] cfree[i] - xref[i-1]; later the solution trajectory
    // QP + solver vectors
    double *g, *xp, *hx;         // [n]
    double *mul;                 // [m]
    double *rhs;                 // [kcap]  right-hand side / solution of the compact system
    double *rs;                  // [kcap]  1 / |v_jj| of the factor
    double *xc;                  // [n]     compact copy of xp for the Hessian product
    double *L;                   // packed lower triangle of V, order <= kcap (+4 doubles of slack)
    double *Hc;                  // GLOBAL: Hessian over the non-fixed variables, [nf][nf]
    uint8_t *fr;                 // [n]  compact -> variable, the non-fixed variables in order
    uint8_t *cpos;               // [n]  variable -> compact (undefined for fixed variables)
    uint8_t *idx;                // [n]  variables of the current KKT system (compact Hessian index)
    uint8_t *grow;               // [kcap] active general rows of the current system (row index - n)
    int8_t *code, *pin, *fixed, *side, *stance;   // [m] [n] [n] [m] [N]
    int nf;
};

__host__ __device__ inline size_t warp_tri(int k) { return (size_t)k * (k + 1) / 2; }
// doubles of one warp's shared-memory slice
__host__ __device__ inline size_t warp_work_doubles(int N, int kcap) {
    const size_t n = 6 * (size_t)N, m = 11 * (size_t)N;
    size_t d = 0;
    d += 4 * N + 2;                 // cz sz PC PS
    d += 18 * N;                    // Bw
    d += 12 + 12 + 6;               // xin Qd Rd
    d += 2 * N + 24;                // hinv hlo blo bhi
    d += 12 * (N + 1);              // err
    d += 4 * n;                     // g xp hx xc
    d += m;                         // mul
    d += 2 * (size_t)kcap;          // rhs rs
    d += warp_tri(kcap) + 4;        // L
    d += (4 * n + (size_t)kcap + 2 * m + N + 7) / 8;   // bytes: fr cpos idx fixed(+pin shares below) ...
    d += (n + 7) / 8;               // pin
    return (d + 1) & ~(size_t)1;
}
__device__ inline void wcarve(WWork& w, double* base, int N, int kcap) {
    const int n = 6 * N, m = 11 * N;
    double* p = base;
    auto take = [&](size_t k) { double* r = p; p += k; return r; };
    w.cz = take(N); w.sz = take(N); w.PC = take(N + 1); w.PS = take(N + 1);
    w.Bw = take(18 * N);
    w.xin = take(12); w.Qd = take(12); w.Rd = take(6);
    w.hinv = take(N); w.hlo = take(N); w.blo = take(12); w.bhi = take(12);
    w.err = take(12 * (N + 1));
    w.g = take(n); w.xp = take(n); w.hx = take(n); w.xc = take(n);
    w.mul = take(m);
    w.rhs = take(kcap); w.rs = take(kcap);
    w.L = take(warp_tri(kcap) + 4);
    uint8_t* q = reinterpret_cast<uint8_t*>(p);
    w.fr = q; q += n; w.cpos = q; q += n; w.idx = q; q += n; w.grow = q; q += kcap;
    w.code = reinterpret_cast<int8_t*>(q); q += m;
    w.side = reinterpret_cast<int8_t*>(q); q += m;
    w.fixed = reinterpret_cast<int8_t*>(q); q += n;
    w.pin = reinterpret_cast<int8_t*>(q); q += n;
    w.stance = reinterpret_cast<int8_t*>(q); q += N;
    w.nf = 0;
}

// ------------------------------------------------------------------------------------------------
// warp primitives
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}
// ordered compaction by one warp: list[0..count) = { first + i : pred(first + i) }, at most cap stored, all counted
template <class Pred>
__device__ inline int wcompact(int first, int len, int cap, uint8_t* list, int sub, int lane, Pred pred) {
    int c = 0;
    for (int base = 0; base < len; base += 32) {
        const int i = base + lane;
        const bool p = (i < len) && pred(first + i);
        const unsigned mask = __ballot_sync(kFullMask, p);
        const int pos = c + __popc(mask & ((1u << lane) - 1u));
        if (p && pos < cap) list[pos] = (uint8_t)(first + i - sub);
        c += __popc(mask);
    }
    return c;
}

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

// box bounds of variable v (mpc_cvx_euler_3f.py:123-128,134-136,146; 2f: fy == 0)
__device__ __forceinline__ double wbox_lo(const WWork& w, int v) { const int k = v / 6; return w.blo[6 * w.stance[k] + (v - 6 * k)]; }
__device__ __forceinline__ double wbox_hi(const WWork& w, int v) { const int k = v / 6; return w.bhi[6 * w.stance[k] + (v - 6 * k)]; }
// bounds of row r in the slot layout [6N box | 4N friction | N height]
__device__ __forceinline__ double wrow_lo(const QpConst& c, const WWork& w, int r) {
    const int N = c.N, n = 6 * N;
    if (r < n) return wbox_lo(w, r);
    if (r < n + 4 * N) return -kInf;
    const int k = r - n - 4 * N;
    return k >= 2 ? w.hlo[k] : -kInf;
}
__device__ __forceinline__ double wrow_hi(const QpConst& c, const WWork& w, int r) {
    const int N = c.N, n = 6 * N;
    if (r < n) return wbox_hi(w, r);
    if (r < n + 4 * N) {
        const int k = (r - n) >> 2, s = (r - n) & 3;
        return (w.stance[k] && (s < 2 || c.dyn == 3)) ? 0.0 : kInf;
    }
    return kInf;
}

// Does general row g (row index - n) touch any unpinned variable?  (A decoupled row gets a unit pivot of its own.)
__device__ __forceinline__ bool wrow_coupled(const QpConst& c, const WWork& w, int g) {
    const int N = c.N;
    if (g < 4 * N) {
        const int k = g >> 2, s = g & 3;
        if (!(w.stance[k] && (s < 2 || c.dyn == 3))) return false;
        return !w.pin[6 * k + (s >> 1)] || !w.pin[6 * k + 2];
    }
    const int k = g - 4 * N;
    for (int j = 0; j + 2 <= k; ++j) if (!w.pin[6 * j + 2]) return true;
    return false;
}

// ------------------------------------------------------------------------------------------------
// linearisation of one stage (gen_dt_dynamics, mpc_cvx_euler_3f.py:82-92 / 2f:82-92); one lane per stage.
// Same arithmetic as linearize_stage (hmpc_qp.cuh); Bv is not stored: 3f  dt/m I,  2f  dt/m Rz^T (wbv).
// ------------------------------------------------------------------------------------------------
__device__ inline void wlinearize_stage(const QpConst& c, const double gp[4], const double pf[3], double* czk,
                                        double* szk, double* Bw) {
    double sn, cs;
    sincos(gp[3], &sn, &cs);
    *czk = cs; *szk = sn;
    const double Rz[9] = {cs, sn, 0, -sn, cs, 0, 0, 0, 1};
    const double d[3] = {pf[0] - gp[0], pf[1] - gp[1], pf[2] - gp[2]};
    double rf[3];
    mat3_vec(Rz, d, rf);
    rf[0] += c.rh[0]; rf[1] += c.rh[1]; rf[2] += c.rh[2];
    double T1[9], Jw[9], JwRzT[9], Bf[9];
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
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0;
            for (int l = 0; l < 3; ++l) a += Jw[3 * i + l] * Rz[3 * j + l];
            JwRzT[3 * i + j] = a;
        }
    if (c.dyn == 3) {
        double rw[3];
        mat3T_vec(Rz, rf, rw);
        const double hatm[9] = {0, -rw[2], rw[1], rw[2], 0, -rw[0], -rw[1], rw[0], 0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double a = 0;
                for (int l = 0; l < 3; ++l) a += Jw[3 * i + l] * hatm[3 * l + j];
                Bf[3 * i + j] = a;
            }
    } else {
        const double hatm[9] = {0, -rf[2], rf[1], rf[2], 0, -rf[0], -rf[1], rf[0], 0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double a = 0;
                for (int l = 0; l < 3; ++l) a += JwRzT[3 * i + l] * hatm[3 * l + j];
                Bf[3 * i + j] = a;
            }
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            Bw[6 * i + j] = Bf[3 * i + j] * c.dt;
            Bw[6 * i + 3 + j] = JwRzT[3 * i + j] * c.dt;
        }
}
// dt * B[6:9, 0:3] of stage k, row-major (linearize_stage's Bv)
__device__ __forceinline__ void wbv(const QpConst& c, const WWork& w, int k, double Bv[9]) {
    const double s = (1.0 / c.m) * c.dt;
    if (c.dyn == 3) {
        Bv[0] = s; Bv[1] = 0; Bv[2] = 0; Bv[3] = 0; Bv[4] = s; Bv[5] = 0; Bv[6] = 0; Bv[7] = 0; Bv[8] = s;
    } else {
        const double cs = w.cz[k], sn = w.sz[k];
        // (Rz^T / m) dt with Rz = [[c, s, 0], [-s, c, 0], [0, 0, 1]]
        Bv[0] = (cs / c.m) * c.dt; Bv[1] = (-sn / c.m) * c.dt; Bv[2] = 0;
        Bv[3] = (sn / c.m) * c.dt; Bv[4] = (cs / c.m) * c.dt; Bv[5] = 0;
        Bv[6] = 0; Bv[7] = 0; Bv[8] = (1.0 / c.m) * c.dt;
    }
}

// ------------------------------------------------------------------------------------------------
// load + time shift + linearise + condense for hopper b (warm tick).  Returns 1 (all lanes) when a height row
// cannot be met.  Leaves Hc (global, compact), g, bounds, fixed / fr / cpos in place.
// ------------------------------------------------------------------------------------------------
__device__ inline int wcondense(const QpConst& c, WWork& w, int b, int B, const MpcIo& io, int lane) {
    const int N = c.N, n = 6 * N;
    const size_t Bs = (size_t)B;
    for (int i = lane; i < 12; i += 32) { w.xin[i] = io.x_in[i * Bs + b]; w.Qd[i] = io.Qd[i * Bs + b]; }
    if (lane < 6) w.Rd[lane] = io.Rd[lane * Bs + b];
    {
        const uint64_t bits = io.Cbits[b];
        for (int k = lane; k < N; k += 32) w.stance[k] = (int8_t)((bits >> k) & 1ull);
    }
    if (lane < 12) {   // box bounds by (stance, component)
        const int st = lane / 6, cc = lane - 6 * st;
        double lo = -kInf, hi = kInf;
        if (cc >= 3) { lo = -c.tau_max[cc - 3]; hi = c.tau_max[cc - 3]; }
        else if (!st) { lo = 0.0; hi = 0.0; }
        else if (cc == 2) { lo = 0.0; hi = c.fz_max; }
        if (cc == 1 && c.dyn == 2) { lo = 0.0; hi = 0.0; }
        w.blo[lane] = lo; w.bhi[lane] = hi;
    }
    __syncwarp();
    // linearisation point: x_guess[0] = x_in, x_guess[k] = x.value[k+1] (mpc_cvx_euler_3f.py:59-62); p and yaw only
    for (int k = lane; k < N; k += 32) {
        double gp[4], pf[3];
        if (k == 0) { gp[0] = w.xin[0]; gp[1] = w.xin[1]; gp[2] = w.xin[2]; gp[3] = w.xin[5]; }
        else {
            const size_t o = (size_t)(k + 1) * 12;
            gp[0] = io.Xsol[(o + 0) * Bs + b]; gp[1] = io.Xsol[(o + 1) * Bs + b];
            gp[2] = io.Xsol[(o + 2) * Bs + b]; gp[3] = io.Xsol[(o + 5) * Bs + b];
        }
        for (int i = 0; i < 3; ++i) pf[i] = io.pf[(size_t)(3 * k + i) * Bs + b];
        wlinearize_stage(c, gp, pf, w.cz + k, w.sz + k, w.Bw + 18 * k);
        w.hinv[k] = (k >= 2) ? 1.0 / (double)(k - 1) : 0.0;
    }
    for (int v = lane; v < n; v += 32) w.fixed[v] = (wbox_hi(w, v) - wbox_lo(w, v)) < 1e-12 ? 1 : 0;
    __syncwarp();
    // prefix sums of cos / sin (same summation order as condense())
    for (int i = lane; i <= N; i += 32) {
        double pc = 0, ps = 0;
        for (int k = 0; k < i; ++k) { pc += w.cz[k]; ps += w.sz[k]; }
        w.PC[i] = pc; w.PS[i] = ps;
    }
    // non-fixed variables, in order
    w.nf = wcompact(0, n, n, w.fr, 0, lane, [&](int v) { return w.fixed[v] == 0; });
    __syncwarp();
    for (int i = lane; i < w.nf; i += 32) w.cpos[w.fr[i]] = (uint8_t)i;
    const double dt = c.dt, gdt = -c.g * dt;
    const double zc = dt * dt / c.m;
    int infeasible = 0;
    // free response + tracking error; height rows
    for (int i = lane; i <= N; i += 32) {
        double cf[12];
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
            for (int q = 0; q < 12; ++q) w.err[12 * i + q] = cf[q] - io.x_ref[((size_t)(i - 1) * 12 + q) * Bs + b];
        if (i < N) {
            double up = 0.0;
            for (int j = 0; j + 2 <= i; ++j) if (w.stance[j]) up += (double)(i - j - 1);
            if (cf[2] + zc * c.fz_max * up < c.z_min) infeasible = 1;
            w.hlo[i] = (i >= 2) ? (c.z_min - cf[2]) / (zc * (double)(i - 1)) : -kInf;
        }
    }
    infeasible = __any_sync(kFullMask, infeasible);
    __syncwarp();
    // ---- Hessian blocks (a >= b), non-fixed entries only, written to both triangles of the compact square ----
    const int nf = w.nf;
    double* Hc = w.Hc;
    const int nblk = N * (N + 1) / 2;
    const double q3 = w.Qd[3], q4 = w.Qd[4], q5 = w.Qd[5];
    const double dt2 = dt * dt;
    for (int p = lane; p < nblk; p += 32) {
        int a = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
        while ((a + 1) * (a + 2) / 2 <= p) ++a;
        while (a * (a + 1) / 2 > p) --a;
        const int bb = p - a * (a + 1) / 2;
        const double pca = w.PC[a + 1], psa = w.PS[a + 1], pcb = w.PC[bb + 1], psb = w.PS[bb + 1];
        double s0 = 0, NN = 0, CC = 0, SS = 0, CS = 0, SC = 0;
        for (int i = a + 1; i <= N; ++i) {
            const double kap = (i == N) ? c.kf : 1.0;
            const double wca = w.PC[i] - pca, wsa = w.PS[i] - psa, wcb = w.PC[i] - pcb, wsb = w.PS[i] - psb;
            const double na = (double)(i - a - 1), nb = (double)(i - bb - 1);
            s0 += kap; NN += kap * na * nb;
            CC += kap * wca * wcb; SS += kap * wsa * wsb; CS += kap * wca * wsb; SC += kap * wsa * wcb;
        }
        const double M3[9] = {s0 * w.Qd[9] + dt2 * (q3 * CC + q4 * SS), dt2 * (q3 * CS - q4 * SC), 0,
                              dt2 * (q3 * SC - q4 * CS), s0 * w.Qd[10] + dt2 * (q3 * SS + q4 * CC), 0,
                              0, 0, s0 * w.Qd[11] + dt2 * q5 * NN};
        const double dv[3] = {dt2 * NN * w.Qd[0] + s0 * w.Qd[6], dt2 * NN * w.Qd[1] + s0 * w.Qd[7],
                              dt2 * NN * w.Qd[2] + s0 * w.Qd[8]};
        const double* Bwa = w.Bw + 18 * a; const double* Bwb = w.Bw + 18 * bb;
        double Bva[9], Bvb[9];
        wbv(c, w, a, Bva); wbv(c, w, bb, Bvb);
        double MB[18];   // M3 * Bw_b (3x6)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j)
                MB[6 * i + j] = M3[3 * i] * Bwb[j] + M3[3 * i + 1] * Bwb[6 + j] + M3[3 * i + 2] * Bwb[12 + j];
        // fully unrolled so that MB / Bva / Bvb stay in registers; fixed rows and columns are skipped by predicate
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const int vi = 6 * a + r;
            if (w.fixed[vi]) continue;
            const int ci = w.cpos[vi];
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) {
                const int vj = 6 * bb + cc;
                if (vi < vj || w.fixed[vj]) continue;        // a == b: the lower half only, mirrored below
                double v = Bwa[r] * MB[cc] + Bwa[6 + r] * MB[6 + cc] + Bwa[12 + r] * MB[12 + cc];
                if (r < 3 && cc < 3)
                    v += Bva[r] * dv[0] * Bvb[cc] + Bva[3 + r] * dv[1] * Bvb[3 + cc] + Bva[6 + r] * dv[2] * Bvb[6 + cc];
                v *= 2.0;
                if (a == bb && r == cc && a != N - 1) v += 2.0 * w.Rd[r];
                const int cj = w.cpos[vj];
                Hc[ci * nf + cj] = v;
                Hc[cj * nf + ci] = v;
            }
        }
    }
    // ---- gradient ----
    const double ubar_alias = (c.uref_mode == 0) ? (w.stance[N - 1] ? 2.0 * c.m * c.g : 0.0) : 0.0;
    for (int a = lane; a < N; a += 32) {
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
        const double* Bwa = w.Bw + 18 * a;
        double Bva[9];
        wbv(c, w, a, Bva);
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
    __syncwarp();
#ifndef HMPC_HOST_EMUL
    __threadfence_block();   // Hc is written and read by lanes of this warp only
#endif
    return infeasible;
}

// ------------------------------------------------------------------------------------------------
// hx = H xp over the non-fixed variables (H compact square in global memory): lanes over the compact row,
// one coalesced row segment per column.  Ends with a __syncwarp().
// ------------------------------------------------------------------------------------------------
template <int SLOTS>
__device__ inline void wmatvec(WWork& w, int lane) {
    const int nf = w.nf;
    for (int j = lane; j < nf; j += 32) w.xc[j] = w.xp[w.fr[j]];
    __syncwarp();
    double acc[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) acc[s] = 0.0;
    const double* Hc = w.Hc;
    int off = lane;
#pragma unroll 4
    for (int j = 0; j < nf; ++j) {
        const double xj = w.xc[j];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int i = lane + 32 * s;
            if (i < nf) acc[s] = fma(Hc[off + 32 * s], xj, acc[s]);
        }
        off += nf;
    }
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int i = lane + 32 * s;
        if (i < nf) w.hx[w.fr[i]] = acc[s];
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Signed Cholesky  K = V S V'  of the compact KKT system of order nk = nF + ng (see LinSys in hmpc_qp.cuh for
// the system: variables idx[0..nF), active general rows grow[0..ng); K_vv = H, K_rv = A.coef, K_rr = 0 with the
// Schur complement's diagonal scaled by (1 + eps) once the variables are eliminated, -1 for a decoupled row).
// V is stored packed, column by column: V(i, j), i >= j, at tri_off(j, nk) + i - j;  rs[j] = 1 / V(j, j).
// Returns nonzero (all lanes) when a pivot has the wrong sign or is not finite.
// ------------------------------------------------------------------------------------------------
template <int SLOTS>
__device__ inline int wfactor(const QpConst& c, WWork& w, const AOp& A, int nF, int ng, double eps, int lane) {
    const int nk = nF + ng, n = 6 * c.N, nf = w.nf;
    double* L = w.L;
    const double* Hc = w.Hc;
    int bad = 0;
    for (int J0 = 0; J0 < nk;) {
        const int lim = (J0 < nF ? nF : nk) - J0;      // blocks never straddle the variable / row boundary
        const int bs = lim < 4 ? lim : 4;
        const int R = nk - J0;
        // lanes -> (row in block, split of the summation index)
        const int S = R > 16 ? 1 : (R > 8 ? 2 : (R > 4 ? 4 : 8));
        const int RP = 32 / S, rl = lane & (RP - 1), q = lane / RP;
        int irow[SLOTS];
        bool on[SLOTS];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int i = J0 + rl + 32 * s;
            on[s] = (i < nk) && (s == 0 || S == 1);
            irow[s] = on[s] ? i : nk - 1;
        }
        // K's entries of the block (independent of the sums below: the loads overlap the update loop)
        double kv[SLOTS][4];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int i = irow[s];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int j = J0 + cc;
                double v = 0.0;
                if (on[s] && cc < bs && i >= j) {
                    if (i < nF) v = Hc[(int)w.idx[j] * nf + (int)w.idx[i]];
                    else if (j < nF) v = A.coef(n + (int)w.grow[i - nF], (int)w.fr[w.idx[j]]);
                    else if (i == j) v = wrow_coupled(c, w, (int)w.grow[i - nF]) ? 0.0 : -1.0;
                }
                kv[s][cc] = v;
            }
        }
        double acc[SLOTS][4];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) acc[s][cc] = 0.0;
        // variable columns: + V(i,k) V(j,k)
        const int k1 = J0 < nF ? J0 : nF;
#pragma unroll 2
        for (int k = q; k < k1; k += S) {
            const double* col = L + (k * nk - ((k * (k + 1)) >> 1));   // col[i] = V(i, k)
            const double b0 = col[J0], b1 = col[J0 + 1], b2 = col[J0 + 2], b3 = col[J0 + 3];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const double a = col[irow[s]];
                acc[s][0] = fma(a, b0, acc[s][0]); acc[s][1] = fma(a, b1, acc[s][1]);
                acc[s][2] = fma(a, b2, acc[s][2]); acc[s][3] = fma(a, b3, acc[s][3]);
            }
        }
        if (J0 >= nF) {
            // the Schur complement of the variables: relative regularisation of its diagonal
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    if (irow[s] == J0 + cc) { acc[s][cc] *= (1.0 + eps); kv[s][cc] *= (1.0 + eps); }
            // active-row columns: - V(i,k) V(j,k)
            for (int k = nF + q; k < J0; k += S) {
                const double* col = L + (k * nk - ((k * (k + 1)) >> 1));
                const double b0 = col[J0], b1 = col[J0 + 1], b2 = col[J0 + 2], b3 = col[J0 + 3];
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const double a = col[irow[s]];
                    acc[s][0] = fma(-a, b0, acc[s][0]); acc[s][1] = fma(-a, b1, acc[s][1]);
                    acc[s][2] = fma(-a, b2, acc[s][2]); acc[s][3] = fma(-a, b3, acc[s][3]);
                }
            }
        }
        // combine the splits (every lane ends up with the total of its row)
        for (int o = RP; o < 32; o <<= 1) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) acc[0][cc] += __shfl_xor_sync(kFullMask, acc[0][cc], o);
        }
        double cv[SLOTS][4];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) cv[s][cc] = kv[s][cc] - acc[s][cc];
        // ---- the block's own columns: pivot by pivot, rows J0 + p live in lane p (slot 0) ----
        const double sgn = (J0 < nF) ? 1.0 : -1.0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (p < bs) {
                const int j = J0 + p;
                const double piv = __shfl_sync(kFullMask, cv[0][p], p);
                const double ap = sgn * piv;
                const bool ok = (ap > 0.0) && (ap < 1e30);
                if (!ok) bad = 1;
                const double rsq = fast_rsqrt(ok ? ap : 1.0);
                double* colj = L + (j * nk - ((j * (j + 1)) >> 1));
                double vn[SLOTS];
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    vn[s] = (irow[s] == j) ? ap * rsq : sgn * cv[s][p] * rsq;
                    if (on[s] && q == 0 && irow[s] >= j) colj[irow[s]] = vn[s];
                }
                if (lane == 0) w.rs[j] = rsq;
#pragma unroll
                for (int p2 = p + 1; p2 < 4; ++p2) {
                    if (p2 < bs) {
                        const double u = sgn * __shfl_sync(kFullMask, vn[0], p2);   // S_j V(J0 + p2, j)
#pragma unroll
                        for (int s = 0; s < SLOTS; ++s) cv[s][p2] = fma(-vn[s], u, cv[s][p2]);
                    }
                }
            }
        }
        __syncwarp();
        J0 += bs;
    }
    return bad;
}

// ------------------------------------------------------------------------------------------------
// Solves K x = b in place on w.rhs (length nk) with K = V S V': forward V z = b, w = S z, backward V' x = w.
// Each lane keeps the entries of its rows (lane, lane + 32, ...) in registers; the pivot entry travels by shuffle.
// Ends with a __syncwarp().
// ------------------------------------------------------------------------------------------------
template <int SLOTS>
__device__ inline void wsolve(WWork& w, int nF, int ng, int lane) {
    const int nk = nF + ng;
    const double* L = w.L;
    double bv[SLOTS], rsv[SLOTS];
    int rowoff[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int i = lane + 32 * s;
        bv[s] = i < nk ? w.rhs[i] : 0.0;
        rsv[s] = i < nk ? w.rs[i] : 0.0;
        rowoff[s] = i < nk ? (i * nk - ((i * (i + 1)) >> 1)) : 0;     // V(j, i) = L[rowoff + j], j >= i
    }
    // forward
#pragma unroll
    for (int sj = 0; sj < SLOTS; ++sj) {
        const int jend = nk - 32 * sj < 32 ? nk - 32 * sj : 32;
#pragma unroll 2
        for (int jl = 0; jl < jend; ++jl) {
            const int j = 32 * sj + jl;
            const double t = __shfl_sync(kFullMask, bv[sj], jl) * w.rs[j];
            const double* col = L + (j * nk - ((j * (j + 1)) >> 1));
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int i = lane + 32 * s;
                if (s >= sj && i > j && i < nk) bv[s] = fma(-col[i], t, bv[s]);
            }
        }
    }
    // z = b / v_jj, w = S z
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int i = lane + 32 * s;
        bv[s] *= (i < nF) ? rsv[s] : -rsv[s];
    }
    // backward
#pragma unroll
    for (int sj = SLOTS - 1; sj >= 0; --sj) {
        const int jend = nk - 32 * sj < 32 ? nk - 32 * sj : 32;
#pragma unroll 2
        for (int jl = jend - 1; jl >= 0; --jl) {
            const int j = 32 * sj + jl;
            const double xj = __shfl_sync(kFullMask, bv[sj], jl) * w.rs[j];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int i = lane + 32 * s;
                if (s <= sj && i < j) bv[s] = fma(-L[rowoff[s] + j], xj, bv[s]);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int i = lane + 32 * s;
        if (i < nk) w.rhs[i] = bv[s] * rsv[s];
    }
    __syncwarp();
}

struct WInfo { int nfac; double flops; };

// ------------------------------------------------------------------------------------------------
// Verified primal-dual active-set refinement for one warp: the same algorithm, tolerances and update rules
// as polish_verified (hmpc_qp.cuh; numpy statement oracle/device_port.py polish_verified).
// On entry w.xp = starting point, w.code = active-set guess.  Returns 1 (all lanes) with w.xp = solution,
// w.mul = multipliers, w.code = active set; 0 when the warm path gives up.
// ------------------------------------------------------------------------------------------------
__device__ inline void wpolish_init(const QpConst& c, WWork& w, int lane) {
    const int N = c.N, n = 6 * N, m = 11 * N;
    for (int r = lane; r < m; r += 32) {
        int cd = w.code[r];
        if (cd > 0 && wrow_hi(c, w, r) > kInfThresh) cd = 0;
        if (cd < 0 && wrow_lo(c, w, r) < -kInfThresh) cd = 0;
        if (r < n && w.fixed[r]) cd = 0;
        w.code[r] = (int8_t)cd;
    }
    __syncwarp();
}
// One trial: 1 = verified optimum (w.xp, w.mul, w.code), 0 = active set updated, try again, -1 = give up.
template <int SLOTS>
__device__ inline int wtrial(const QpConst& c, WWork& w, const AOp& A, int kcap, WInfo& info, int lane) {
    const int N = c.N, n = 6 * N, m = 11 * N;
    const double tol = c.polish_tol;
    double* mul = w.mul;
    {
        for (int r = lane; r < m; r += 32) {
            const int cd = w.code[r];
            mul[r] = 0.0;
            if (r < n) {
                const int pin = (w.fixed[r] || cd != 0) ? 1 : 0;
                w.pin[r] = (int8_t)pin;
                if (pin) w.xp[r] = w.fixed[r] ? 0.0 : (cd < 0 ? wbox_lo(w, r) : wbox_hi(w, r));
            }
        }
        __syncwarp();
        // system variables as compact Hessian indices (the non-fixed, unpinned variables in order)
        const int nF = wcompact(0, w.nf, n, w.idx, 0, lane, [&](int i) { return w.pin[w.fr[i]] == 0; });
        const int ng = wcompact(n, m - n, kcap, w.grow, n, lane, [&](int r) { return w.code[r] != 0; });
        __syncwarp();
        const int nk = nF + ng;
        if (nk > kcap || nk > 32 * SLOTS) return -1;
        ++info.nfac;
        info.flops += flops_factor(nk);
        if (wfactor<SLOTS>(c, w, A, nF, ng, c.kkt_eps, lane)) return -1;
        double prev = 1e300;
        bool hx_current = false;
        for (int k = 0; k < c.max_refine; ++k) {
            wmatvec<SLOTS>(w, lane);
            info.flops += flops_matvec(n);
            double v0 = 0.0, v1 = 0.0;   // residual; largest residual relative to the terms it is the difference of
            for (int i = lane; i < nk; i += 32) {
                double r_, mag;
                if (i < nF) {
                    const int vi = w.fr[w.idx[i]];
                    const double aty = A.colT(vi, mul);
                    r_ = -(w.hx[vi] + w.g[vi] + aty);
                    mag = fabs(w.hx[vi]) + fabs(w.g[vi]) + fabs(aty);
                } else {
                    const int rr = n + (int)w.grow[i - nF];
                    const double ax = A.row(rr, w.xp);
                    const double bnd = w.code[rr] < 0 ? wrow_lo(c, w, rr) : wrow_hi(c, w, rr);
                    r_ = bnd - ax;
                    mag = fabs(bnd) + fabs(ax);
                }
                w.rhs[i] = r_;
                v0 = fmax(v0, fabs(r_));
                v1 = fmax(v1, fabs(r_) / (mag + 1e-300));
            }
            v0 = wmax(v0); v1 = wmax(v1);
            __syncwarp();
            if (!(v0 == v0)) return -1;
            if (k == 1 && ng == 0 && v0 <= 1e-7 * prev) { hx_current = true; break; }
            if (k >= 1 && (v1 <= 1e-12 || (k >= 2 && v0 > c.stagnation * prev))) { hx_current = true; break; }
            prev = v0;
            wsolve<SLOTS>(w, nF, ng, lane);
            info.flops += flops_solve(nk);
            for (int i = lane; i < nk; i += 32) {
                if (i < nF) w.xp[w.fr[w.idx[i]]] += w.rhs[i];
                else mul[n + (int)w.grow[i - nF]] += w.rhs[i];
            }
            __syncwarp();
        }
        // ---- pass 1: multipliers of pinned variables, scales ----
        if (!hx_current) { wmatvec<SLOTS>(w, lane); info.flops += flops_matvec(n); }
        double s_stat = 0.0, s_scale = 0.0, s_mult = 0.0;
        for (int i = lane; i < n; i += 32) {
            if (w.fixed[i]) continue;                     // eliminated a priori: no Hessian row, multiplier unused
            const double aty = A.colT(i, mul);            // mul[i] == 0 on box rows at this point
            const double G = w.hx[i] + w.g[i] + aty;
            s_scale = fmax(s_scale, fmax(fabs(w.hx[i]), fmax(fabs(w.g[i]), fabs(aty))));
            if (!w.pin[i]) s_stat = fmax(s_stat, fabs(G));
            else { mul[i] = -G; s_mult = fmax(s_mult, fabs(G)); }
        }
        for (int r = n + lane; r < m; r += 32) if (w.code[r]) s_mult = fmax(s_mult, fabs(mul[r]));
        s_stat = wmax(s_stat); s_scale = wmax(s_scale); s_mult = wmax(s_mult);
        __syncwarp();
        const double scale = fmax(1.0, s_scale);
        const double stol = tol * fmax(scale, s_mult);
        // ---- pass 2: per-row verdicts and the refined active set ----
        int bad = (s_stat <= 1e-10 * scale) ? 0 : 1, anywrong = 0;
        const int nonfinite = (!(s_stat == s_stat) || !(s_mult == s_mult)) ? 1 : 0;
        for (int r = lane; r < m; r += 32) {
            const bool apriori = (r < n) && w.fixed[r];
            const int cd = w.code[r];
            int change = 0;
            if (!apriori) {
                const double ax = A.row(r, w.xp), lo = wrow_lo(c, w, r), hi = wrow_hi(c, w, r);
                if (cd != 0) {
                    const double lam = mul[r];
                    const double bnd = cd < 0 ? lo : hi;
                    if ((cd > 0 && lam < -stol) || (cd < 0 && lam > stol)) { change = 1; bad |= 1; anywrong = 1; }
                    if (r >= n && fabs(ax - bnd) > tol * (1.0 + fabs(bnd))) bad |= 1;   // singular / inconsistent set
                } else {
                    if (lo - ax > tol * (1.0 + fabs(lo))) { change = 2; bad |= 1; }
                    else if (ax - hi > tol * (1.0 + fabs(hi))) { change = 3; bad |= 1; }
                }
            }
            w.side[r] = (int8_t)change;
        }
        bad = __any_sync(kFullMask, bad);
        anywrong = __any_sync(kFullMask, anywrong);
        __syncwarp();
        if (nonfinite) return -1;
        if (!bad) {
            for (int r = n + lane; r < m; r += 32) if (!w.code[r]) mul[r] = 0.0;
            __syncwarp();
            return 1;
        }
        int changed = 0;
        for (int r = lane; r < m; r += 32) {
            const int change = w.side[r];
            if (change == 1) { w.code[r] = 0; changed = 1; }
            else if (change >= 2 && !anywrong) { w.code[r] = (int8_t)(change == 2 ? -1 : 1); changed = 1; }
        }
        changed = __any_sync(kFullMask, changed);
        __syncwarp();
        return changed ? 0 : -1;
    }
}
// Verified primal-dual active-set refinement, all trials (free-running form).
template <int SLOTS>
__device__ inline int wpolish(const QpConst& c, WWork& w, const AOp& A, int kcap, WInfo& info, int lane) {
    wpolish_init(c, w, lane);
    for (int trial = 0; trial <= c.retries; ++trial) {
        const int r = wtrial<SLOTS>(c, w, A, kcap, info, lane);
        if (r != 0) return r > 0;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// One warm tick of hopper b by one warp.  Returns 1 when the hopper is done (outputs and handle state written),
// 0 when it has to take the CTA kernel (nothing written).
// ------------------------------------------------------------------------------------------------
// wbegin: load, time shift, linearise, condense, warm start.  Returns 0 when the hopper must take the CTA kernel.
__device__ inline int wbegin(const QpConst& c, WWork& w, int b, int B, const MpcIo& io, int lane) {
    const int N = c.N, n = 6 * N, m = 11 * N;
    const size_t Bs = (size_t)B;
    if (!io.valid[b]) return 0;
    if (wcondense(c, w, b, B, io, lane)) return 0;
    // warm start: stage k starts from the previous tick's stage k+1, the last two stages keep their own previous
    // pattern (mpc_hopper in hmpc_mpc.cuh)
    for (int i = lane; i < n; i += 32) {
        const int src = (i / 6 < N - 2) ? i + 6 : i;
        w.xp[i] = io.Usol[(size_t)src * Bs + b];
    }
    for (int r = lane; r < m; r += 32) {
        int src;
        if (r < n) src = (r / 6 < N - 2) ? r + 6 : r;
        else if (r < n + 4 * N) src = ((r - n) / 4 < N - 2) ? r + 4 : r;
        else src = (r - n - 4 * N < N - 2) ? r + 1 : r;
        w.code[r] = io.code[(size_t)src * Bs + b];
    }
    __syncwarp();
    wpolish_init(c, w, lane);
    return 1;
}
// wfinish: roll the verified solution out, store outputs and the handle state of hopper b.
__device__ inline void wfinish(const QpConst& c, WWork& w, int b, int B, const MpcIo& io, const WInfo& info, int lane) {
    const int N = c.N, n = 6 * N, m = 11 * N;
    const size_t Bs = (size_t)B;
    // ---- roll the solution out (mpc_cvx_euler_3f.py:133,140 dynamics rows) and store ----
    double* xs = w.err;
    const double* u = w.xp;
    const double dt = c.dt, gdt = -c.g * dt;
    if (lane < 12) xs[lane] = w.xin[lane];
    if (lane < 6) {          // velocities
        const int q = lane;
        double acc = w.xin[6 + q];
        for (int k = 0; k < N; ++k) {
            const double* uk = u + 6 * k;
            if (q < 3) {
                // row q of dt B[6:9, 0:3]: 3f  dt/m e_q,  2f  dt/m Rz^T (same products as rollout_solution)
                const double sm = (1.0 / c.m) * c.dt;
                if (c.dyn == 3) acc += sm * uk[q];
                else {
                    const double cs = w.cz[k], sn = w.sz[k];
                    if (q == 0) acc += ((cs / c.m) * c.dt) * uk[0] + ((-sn / c.m) * c.dt) * uk[1];
                    else if (q == 1) acc += ((sn / c.m) * c.dt) * uk[0] + ((cs / c.m) * c.dt) * uk[1];
                    else acc += sm * uk[2];
                }
                if (q == 2) acc += gdt;
            } else {
                const double* Bw = w.Bw + 18 * k + 6 * (q - 3);
                acc += Bw[0] * uk[0] + Bw[1] * uk[1] + Bw[2] * uk[2] + Bw[3] * uk[3] + Bw[4] * uk[4] + Bw[5] * uk[5];
            }
            xs[12 * (k + 1) + 6 + q] = acc;
        }
    }
    __syncwarp();
    if (lane < 6) {          // positions / Euler angles integrate the stage-k velocities
        const int q = lane;
        double acc = w.xin[q];
        for (int k = 0; k < N; ++k) {
            const double* xk = xs + 12 * k;
            if (q < 3) acc += dt * xk[6 + q];
            else {
                const double cs = w.cz[k], sn = w.sz[k];
                const double wx = xk[9], wy = xk[10], wz = xk[11];
                const double r = (q == 3) ? (cs * wx + sn * wy) : (q == 4) ? (-sn * wx + cs * wy) : wz;
                acc += dt * r;
            }
            xs[12 * (k + 1) + q] = acc;
        }
    }
    __syncwarp();
    for (int i = lane; i < (N + 1) * 12; i += 32) {
        io.Xsol[(size_t)i * Bs + b] = xs[i];
        if (io.X_out) io.X_out[(size_t)i * Bs + b] = xs[i];
    }
    for (int i = lane; i < n; i += 32) {
        io.Usol[(size_t)i * Bs + b] = u[i];
        if (io.U_out) io.U_out[(size_t)i * Bs + b] = u[i];
    }
    for (int r = lane; r < m; r += 32) io.code[(size_t)r * Bs + b] = w.code[r];
    if (io.U0_out && lane < 6) io.U0_out[(size_t)lane * Bs + b] = u[lane];
    if (lane == 0) {
        io.valid[b] = 1;
        if (io.flops) io.flops[b] = (io.accumulate ? io.flops[b] : 0.0) + info.flops;
        io.st_tick[b] = ST_SOLVED;
        io.path[b] = PATH_WARM;
        if (io.accumulate) {
            io.nfac[b] += info.nfac;
        } else {
            io.status[b] = ST_SOLVED;
            io.iters[b] = 0;
            io.nfac[b] = info.nfac;
            io.ninf[b] = 0;
        }
    }
}
template <int SLOTS>
__device__ inline int mpc_hopper_warp(const QpConst& c, WWork& w, int kcap, int b, int B, const MpcIo& io, int lane) {
    if (!wbegin(c, w, b, B, io, lane)) return 0;
    AOp A{c.N, 6 * c.N, c.dyn == 3 ? 1 : 0, c.mu, w.stance, w.hinv};
    WInfo info{0, c.condense_flops};
    for (int trial = 0; trial <= c.retries; ++trial) {
        const int r = wtrial<SLOTS>(c, w, A, kcap, info, lane);
        if (r < 0) return 0;
        if (r > 0) { wfinish(c, w, b, B, io, info, lane); return 1; }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Host-side dispatch rules, shared by the library (hmpc_api.cu) and the test emulation (tests/emul).
// ------------------------------------------------------------------------------------------------
constexpr int kWarpMaxN = 10;     // horizons the warp kernel is instantiated for (n = 6N <= 64: two rows per lane)
// cap on the order of the compact KKT system (unpinned variables + active friction / height rows): it sizes the
// factor in shared memory; a trial that needs more goes to the CTA kernel
inline int warp_kcap(const hmpc_config& cfg) {
    int k = 7 * cfg.N + 2;
    const int lim = cfg.N <= 10 ? 56 : 104;
    if (k > lim) k = lim;
    if (const char* e = getenv("HMPC_WARP_KCAP")) { const int v = atoi(e); if (v >= 8 && v <= 128) k = v; }
    return k;
}
inline bool warp_path_applies(const hmpc_config& cfg, int init) {
    return cfg.hot_path == HMPC_HOT_AUTO && !init && cfg.warm_start && cfg.solver == HMPC_SOLVER_EXACT &&
           cfg.precision == HMPC_FP64 && cfg.sqp_sweeps <= 1 && cfg.N >= 3 && cfg.N <= kWarpMaxN;
}

#ifndef HMPC_HOST_EMUL
// ------------------------------------------------------------------------------------------------
// Persistent kernel: WPC independent warps per CTA, hoppers handed out one at a time (the solve times differ).
// hws: per-warp Hessian workspace [grid * WPC][hstride] doubles.  Deferred hoppers are appended to defer_list.
// ------------------------------------------------------------------------------------------------
template <int SLOTS, int WPC, int MIN_CTAS>
__global__ void __launch_bounds__(32 * WPC, MIN_CTAS)
mpc_warp_kernel(QpConst c, int B, int kcap, int wdoubles, double* __restrict__ hws, size_t hstride,
                int* __restrict__ work_ctr, int* __restrict__ defer_list, int* __restrict__ defer_cnt, MpcIo io) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WWork w;
    wcarve(w, smem + (size_t)wid * wdoubles, c.N, kcap);
    w.Hc = hws + ((size_t)blockIdx.x * WPC + wid) * hstride;
    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(work_ctr, 1);
        b = __shfl_sync(kFullMask, b, 0);
        if (b >= B) break;
        const int done = mpc_hopper_warp<SLOTS>(c, w, kcap, b, B, io, lane);
        __syncwarp();
        if (!done && lane == 0) defer_list[atomicAdd(defer_cnt, 1)] = b;
    }
}

// Lock-step variant: the WPC warps of a CTA still own one hopper each, but start every active-set trial together
// (one CTA barrier per round), so that the warps sharing an SM execute the same code at the same time and share its
// instruction fetches.  A warp whose hopper is finished fetches and condenses the next one while the others wait.
template <int SLOTS, int WPC, int MIN_CTAS>
__global__ void __launch_bounds__(32 * WPC, MIN_CTAS)
mpc_warp_rounds_kernel(QpConst c, int B, int kcap, int wdoubles, double* __restrict__ hws, size_t hstride,
                       int* __restrict__ work_ctr, int* __restrict__ defer_list, int* __restrict__ defer_cnt, MpcIo io) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WWork w;
    wcarve(w, smem + (size_t)wid * wdoubles, c.N, kcap);
    w.Hc = hws + ((size_t)blockIdx.x * WPC + wid) * hstride;
    AOp A{c.N, 6 * c.N, c.dyn == 3 ? 1 : 0, c.mu, w.stance, w.hinv};
    bool have = false, exhausted = false;
    int b = 0, trial = 0;
    WInfo info{0, 0.0};
    for (;;) {
        while (!have && !exhausted) {
            if (lane == 0) b = atomicAdd(work_ctr, 1);
            b = __shfl_sync(kFullMask, b, 0);
            if (b >= B) { exhausted = true; break; }
            if (wbegin(c, w, b, B, io, lane)) { have = true; trial = 0; info.nfac = 0; info.flops = c.condense_flops; }
            else if (lane == 0) defer_list[atomicAdd(defer_cnt, 1)] = b;
        }
        if (__syncthreads_and(!have)) break;
        if (have) {
            const int r = wtrial<SLOTS>(c, w, A, kcap, info, lane);
            if (r > 0) { wfinish(c, w, b, B, io, info, lane); have = false; }
            else if (r < 0 || ++trial > c.retries) {
                if (lane == 0) defer_list[atomicAdd(defer_cnt, 1)] = b;
                have = false;
            }
            __syncwarp();
        }
    }
}
#endif  // HMPC_HOST_EMUL

}  // namespace hmpc
