// hmpc_plan.cuh -- the caller side of the hot path on the device: reference rows, footstep rows,
// contact masks and footstep switch steps generated per tick from a handful of per-hopper scalars instead of being
// read from host-built tables.
//
// Reference behaviour (file:line into the reference's src/):
//   path_plan_init   robotrunner.py:182-226  straight-line interpolation start -> goal (np.linspace), the --curve
//                    parabolas (a 3-point not-a-knot CubicSpline IS the parabola through its knots) written into
//                    columns 0 and 5 (column quirk), finite-difference yaw rate and velocities, the
//                    height sine wave, "sit at the goal" rows, footsteps = reference xy at the touch-down indices
//   path_plan_grab   robotrunner.py:228-230  MPC-rate rows k, k + mpc_factor, ...
//   gait_map         robotrunner.py:172-180  contact flags from running float sums (host, shipped as bit masks)
// The arithmetic restates planner.batch_tables (hopper_mpc_inertial_b200/planner.py) operation by operation with
// explicitly rounded multiplies / adds (no FMA contraction), so that the generated rows are BIT-IDENTICAL to the
// numpy tables; everything that depends only on the common clock (sine of the height wave, touch-down index per
// sim step, contact mask and switch step per tick) comes in as small global tables computed once on the host.
#pragma once
#include <stdint.h>

namespace hmpc {

struct PlanConst {
    int N, mpc_factor;
    int N_run;            // sim steps of the run (robotrunner.py:183)
    int t_ref;            // N_run + N * mpc_factor rows of the full-rate reference
    int n_sim;            // entries of sin_tab / pf_idx
    int max_tick;         // entries of cmask / sw_glob
    double dt;            // simulator step
    double amp;           // t_p / 4
    double T;             // (double)N_run
    double curve_psi1, curve_psi2;   // -0.4 sin(45 deg), -sin(45 deg)
    // global tables (device)
    const double* sin_tab;   // [n_sim]   sin(2 pi / t_p (k dt) + 3 pi / 2)
    const int32_t* pf_idx;   // [n_sim]   sim index whose reference xy is the footstep in force at sim step k
    const uint64_t* cmask;   // [max_tick] contact mask of global tick j (bit k = stance at horizon stage k)
    const uint8_t* sw_glob;  // [max_tick] sim step inside tick j at which the footstep index changes (mpc_factor = never)
    // per-hopper parameters (device, SoA)
    const double* x0;        // [12][B] start of the reference
    const double* xf;        // [12][B] goal
    const int32_t* curve;    // [B]
    const int32_t* off;      // [B] tick at which the hopper enters the run
};

#ifdef HMPC_HOST_EMUL
__device__ inline double pmul(double a, double b) { volatile double r = a * b; return r; }
__device__ inline double padd(double a, double b) { volatile double r = a + b; return r; }
__device__ inline double pdiv(double a, double b) { return a / b; }
#else
__device__ __forceinline__ double pmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double padd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double pdiv(double a, double b) { return __ddiv_rn(a, b); }
#endif
__device__ __forceinline__ double psub(double a, double b) { return padd(a, -b); }

// y0 (1 - s)(1 - 2 s) + 4 y1 s (1 - s) + y2 s (2 s - 1), s = k / T, in planner._parabola's evaluation order
__device__ inline double plan_parabola(double y0, double y1, double y2, double T, double k) {
    const double s = pdiv(k, T);
    const double a = pmul(pmul(y0, psub(1.0, s)), psub(1.0, pmul(2.0, s)));
    const double b = pmul(pmul(pmul(y1, 4.0), s), psub(1.0, s));
    const double c = pmul(pmul(y2, s), psub(pmul(2.0, s), 1.0));
    return padd(padd(a, b), c);
}

// Per-hopper scalars of the planner held in registers.
struct PlanHopper {
    double x0[12], xf[12], step[12];
    int curve;
};
__device__ inline void plan_load(const PlanConst& P, int b, int B, PlanHopper& h) {
    const double den = (double)(P.N_run - 1);
#pragma unroll
    for (int q = 0; q < 12; ++q) {
        h.x0[q] = P.x0[(size_t)q * B + b];
        h.xf[q] = P.xf[(size_t)q * B + b];
        h.step[q] = pdiv(psub(h.xf[q], h.x0[q]), den);        // np.linspace step
    }
    h.curve = P.curve[b];
}

// x_ref[i] without the velocity columns 6:9 (planner.batch_tables: ref_rows), i = sim index
__device__ inline void plan_ref_row(const PlanConst& P, const PlanHopper& h, int i, double out[12]) {
    const int ic = i < P.t_ref - 1 ? i : P.t_ref - 1;
    const double k = (double)ic;
    double lin[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) {
        lin[q] = padd(pmul(k, h.step[q]), h.x0[q]);
        if (ic == P.N_run - 1) lin[q] = h.xf[q];                // the endpoint of np.linspace is exact
        out[q] = lin[q];
    }
    if (h.curve) {
        out[0] = plan_parabola(h.x0[1], pmul(0.9, h.xf[1]), h.xf[1], P.T, k);
        out[5] = plan_parabola(0.0, P.curve_psi1, P.curve_psi2, P.T, k);
        if (ic < P.N_run - 1) {
            const double nxt = (ic + 1 == P.N_run - 1) ? h.xf[11] : padd(pmul(k + 1.0, h.step[11]), h.x0[11]);
            out[11] = pdiv(psub(nxt, lin[11]), P.dt);
        }
    }
    if (ic >= P.N_run) {
#pragma unroll
        for (int q = 0; q < 12; ++q) out[q] = h.xf[q];           // sitting at the goal
    }
    out[2] = padd(padd(h.x0[2], P.amp), pmul(P.amp, P.sin_tab[ic < P.n_sim ? ic : P.n_sim - 1]));
}

// MPC-rate row j of hopper b: xref (12, velocities by forward difference) and the footstep (3)
__device__ inline void plan_table_row(const PlanConst& P, const PlanHopper& h, int off, int j, double xr[12], double pf[3]) {
    long long kk = (long long)(off + j) * P.mpc_factor;
    const int k = kk < P.t_ref - 1 ? (int)kk : P.t_ref - 1;
    double r1[12];
    plan_ref_row(P, h, k, xr);
    plan_ref_row(P, h, k + 1, r1);
    if (k == P.t_ref - 1) { xr[6] = h.xf[6]; xr[7] = h.xf[7]; xr[8] = h.xf[8]; }   // the last row keeps the goal's
    else {
#pragma unroll
        for (int q = 0; q < 3; ++q) xr[6 + q] = pdiv(psub(r1[q], xr[q]), P.dt);
    }
    const int sel = P.pf_idx[k < P.n_sim ? k : P.n_sim - 1];
    plan_ref_row(P, h, sel, r1);
    pf[0] = r1[0]; pf[1] = r1[1]; pf[2] = 0.0;
}

#ifndef HMPC_HOST_EMUL
// rows [row0, row0 + nrows) of the MPC-rate tables: xref_out [nrows_x][12][B] (rows < nrows_x only), pf_out [nrows][3][B]
__global__ void __launch_bounds__(128)
plan_rows_kernel(PlanConst P, int B, int row0, int nrows, int nrows_x, double* __restrict__ xref_out, double* __restrict__ pf_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    PlanHopper h;
    plan_load(P, b, B, h);
    const int off = P.off[b];
    for (int r = blockIdx.y; r < nrows; r += gridDim.y) {
        double xr[12], pf[3];
        plan_table_row(P, h, off, row0 + r, xr, pf);
        if (r < nrows_x) {
#pragma unroll
            for (int q = 0; q < 12; ++q) xref_out[((size_t)r * 12 + q) * B + b] = xr[q];
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) pf_out[((size_t)r * 3 + q) * B + b] = pf[q];
    }
}
// contact masks and switch steps of ticks [tick0, tick0 + n_ticks); pf_rows = the footstep rows of the same range
// (row t and t + 1 equal -> the footstep does not change inside tick t)
__global__ void __launch_bounds__(128)
plan_masks_kernel(PlanConst P, int B, int tick0, int n_ticks, const double* __restrict__ pf_rows, uint64_t* __restrict__ C_out,
                  uint8_t* __restrict__ sw_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int off = P.off[b];
    for (int t = blockIdx.y; t < n_ticks; t += gridDim.y) {
        int j = off + tick0 + t;
        if (j >= P.max_tick) j = P.max_tick - 1;
        C_out[(size_t)t * B + b] = P.cmask[j];
        bool same = true;
#pragma unroll
        for (int q = 0; q < 3; ++q) same = same && (pf_rows[((size_t)t * 3 + q) * B + b] == pf_rows[((size_t)(t + 1) * 3 + q) * B + b]);
        sw_out[(size_t)t * B + b] = same ? (uint8_t)P.mpc_factor : P.sw_glob[j];
    }
}
#endif

}  // namespace hmpc
