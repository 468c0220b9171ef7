// oracle/c/hopper_ref.cpp -- *** TEST / BASELINE INFRASTRUCTURE ONLY *** (never linked or loaded by the product).
//
// Compiled C++ restatement of the reference's per-tick recipe, used (a) as the CPU baseline that bench.py times on
// the host cores (cpu_baseline.kind = "port-c++") and (b) as a second, independent implementation the
// numpy oracle is checked against (tests/test_oracle_c.py: same QP data, same OSQP iteration counts, same iterates).
//
// What is restated (file:line into /root/reference/src unless marked [EXT]):
//   convert                  robotrunner.py:19-28        quat2euler: utils.py:54-62 + transforms3d 'rzyx' [EXT]
//   dynamics_ct / rk4        robotrunner.py:126-164      (np.linalg.solve(J, .) -> 3x3 elimination with pivoting)
//   gen_dt_dynamics          mpc_cvx_euler_3f.py:71-94 / mpc_cvx_euler_2f.py:70-94
//   build_qp (cvxpy-shaped)  mpc_cvx_euler_3f.py:96-153 / 2f:96-151, SURVEY App. A: v = [x(0..N); u(0..N-1)],
//                            rows in the order of oracle/hopper_oracle.py:build_qp_full, u_ref aliasing (App. D1)
//   mpcontrol                mpc_cvx_euler_3f.py:41-69   two solves on the first call, time shift afterwards
//   solve                    cp.Problem(...).solve(solver=cp.OSQP): OSQP is a third-party dependency absent from the
//                            reference tree and from this image (README.md:45 pins no version).  Restated from the
//                            published algorithm (Stellato et al., Math. Prog. Comp. 2020) with the 0.6.x defaults
//                            and cvxpy's settings: eps_abs = eps_rel = 1e-5, max_iter 10000, rho 0.1, sigma 1e-6,
//                            alpha 1.6, Ruiz scaling 10, check_termination 25, adaptive rho (fixed interval 50 instead
//                            of OSQP's wall-clock rule, tolerance 5), polish (delta 1e-6, 3 refinements), cold start
//                            [EXT] -- the same statement as oracle/qp_solvers.py:osqp_solve.
// Linear algebra: the quasi-definite KKT matrix is permuted with reverse Cuthill-McKee (the MPC stage chain gives a
// narrow band) and factorised as a banded LDL' without pivoting (OSQP itself: AMD + QDLDL sparse LDL' [EXT]).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

namespace {

const double kInf = 1e30, kInfThresh = 1e26, kMinScaling = 1e-4, kMaxScaling = 1e4;
const double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoTol = 1e-4, kRhoEqFactor = 1e3;

typedef std::vector<double> vec;

// ---------------------------------------------------------------------------------------------------------------
// sparse matrix in coordinate form, rows x cols
// ---------------------------------------------------------------------------------------------------------------
struct Coo {
    int rows = 0, cols = 0;
    std::vector<int> r, c;
    vec v;
    void add(int i, int j, double x) { r.push_back(i); c.push_back(j); v.push_back(x); }
    void mul(const vec& x, vec& y) const {            // y = M x
        y.assign(rows, 0.0);
        for (size_t k = 0; k < v.size(); ++k) y[r[k]] += v[k] * x[c[k]];
    }
    void mulT(const vec& x, vec& y) const {           // y = M' x
        y.assign(cols, 0.0);
        for (size_t k = 0; k < v.size(); ++k) y[c[k]] += v[k] * x[r[k]];
    }
};

// ---------------------------------------------------------------------------------------------------------------
// symmetric banded LDL' after a reverse Cuthill-McKee permutation
// ---------------------------------------------------------------------------------------------------------------
struct BandLdl {
    int n = 0, b = 0;
    std::vector<int> perm, iperm;       // perm[new] = old
    vec band;                           // band[i * w + (b + k - i)] = L(i, k), k = i-b .. i-1;  band[i * w + b] = D(i);  w = b + 1
    vec tmp;

    static void rcm(int n, const std::vector<std::vector<int>>& adj, std::vector<int>& perm) {
        std::vector<int> deg(n), order;
        std::vector<char> seen(n, 0);
        for (int i = 0; i < n; ++i) deg[i] = (int)adj[i].size();
        order.reserve(n);
        auto bfs = [&](int start, std::vector<int>& out, std::vector<char>& mark) {
            size_t head = out.size();
            out.push_back(start); mark[start] = 1;
            std::vector<int> nb;
            while (head < out.size()) {
                const int u = out[head++];
                nb.clear();
                for (int w : adj[u]) if (!mark[w]) { mark[w] = 1; nb.push_back(w); }
                std::sort(nb.begin(), nb.end(), [&](int a, int c2) { return deg[a] != deg[c2] ? deg[a] < deg[c2] : a < c2; });
                for (int w : nb) out.push_back(w);
            }
        };
        for (;;) {
            int start = -1;
            for (int i = 0; i < n; ++i) if (!seen[i] && (start < 0 || deg[i] < deg[start])) start = i;
            if (start < 0) break;
            // pseudo-peripheral node: last node of a BFS from the minimum-degree node, twice
            for (int pass = 0; pass < 2; ++pass) {
                std::vector<int> probe;
                std::vector<char> mark(seen);
                bfs(start, probe, mark);
                start = probe.back();
            }
            bfs(start, order, seen);
        }
        perm.assign(order.rbegin(), order.rend());
    }

    // K given by the entries (i, j, v) of ONE triangle (plus the diagonal); duplicates are summed.
    // Returns false on a zero / non-finite pivot.
    bool factor(int n_, const std::vector<int>& ri, const std::vector<int>& ci, const vec& vi, bool reuse_perm) {
        if (!reuse_perm || (int)perm.size() != n_) {
            n = n_;
            std::vector<std::vector<int>> adj(n);
            for (size_t k = 0; k < vi.size(); ++k)
                if (ri[k] != ci[k]) { adj[ri[k]].push_back(ci[k]); adj[ci[k]].push_back(ri[k]); }
            for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
            rcm(n, adj, perm);
            iperm.assign(n, 0);
            for (int i = 0; i < n; ++i) iperm[perm[i]] = i;
            b = 0;
            for (size_t k = 0; k < vi.size(); ++k) b = std::max(b, abs(iperm[ri[k]] - iperm[ci[k]]));
        }
        const int w = b + 1;
        band.assign((size_t)n * w, 0.0);
        for (size_t k = 0; k < vi.size(); ++k) {
            int i = iperm[ri[k]], j = iperm[ci[k]];
            if (i < j) std::swap(i, j);
            band[(size_t)i * w + (b + j - i)] += vi[k];
        }
        // row-by-row (up-looking) banded LDL': every inner loop runs over contiguous pieces of two rows
        vec dl(w);                                            // L(i, k) D(k) of the current row
        for (int i = 0; i < n; ++i) {
            double* ri_ = &band[(size_t)i * w];               // ri_[b + k - i] = K(i, k) -> L(i, k)
            const int k0 = std::max(0, i - b);
            for (int j = k0; j < i; ++j) {
                const double* rj = &band[(size_t)j * w];      // rj[b + k - j] = L(j, k)
                const int kk0 = std::max(k0, j - b);
                double s = ri_[b + j - i];
                const double* a = dl.data() + (kk0 - k0);
                const double* c2 = rj + (b + kk0 - j);
                for (int k = 0; k < j - kk0; ++k) s -= a[k] * c2[k];      // sum_k L(i,k) D(k) L(j,k)
                dl[j - k0] = s;                                          // = L(i, j) D(j)
                ri_[b + j - i] = s / rj[b];
            }
            double d = ri_[b];
            for (int j = k0; j < i; ++j) d -= dl[j - k0] * ri_[b + j - i];
            if (d == 0.0 || !(fabs(d) < 1e300)) return false;
            ri_[b] = d;
        }
        return true;
    }

    void solve(vec& x) {                     // in place, original ordering
        const int w = b + 1;
        tmp.resize(n);
        for (int i = 0; i < n; ++i) tmp[i] = x[perm[i]];
        for (int i = 0; i < n; ++i) {
            const int k0 = std::max(0, i - b);
            const double* r = &band[(size_t)i * w + (b + k0 - i)];
            const double* t = tmp.data() + k0;
            double s = tmp[i];
            for (int k = 0; k < i - k0; ++k) s -= r[k] * t[k];
            tmp[i] = s;
        }
        for (int i = 0; i < n; ++i) tmp[i] /= band[(size_t)i * w + b];
        for (int i = n - 1; i >= 0; --i) {
            const int k0 = std::max(0, i - b);
            const double* r = &band[(size_t)i * w + (b + k0 - i)];
            double* t = tmp.data() + k0;
            const double xi = tmp[i];
            for (int k = 0; k < i - k0; ++k) t[k] -= r[k] * xi;
        }
        for (int i = 0; i < n; ++i) x[perm[i]] = tmp[i];
    }
};

// ---------------------------------------------------------------------------------------------------------------
// OSQP
// ---------------------------------------------------------------------------------------------------------------
struct OsqpOpts {
    double eps_abs = 1e-5, eps_rel = 1e-5, rho = 0.1, sigma = 1e-6, alpha = 1.6, delta = 1e-6, adapt_tol = 5.0;
    int max_iter = 10000, scaling = 10, check = 25, adapt_interval = 50, polish = 1, polish_refine = 3;
};
struct OsqpResult { int status = 1, iters = 0, polished = 0, n_fac = 0; double rho = 0, pri = 0, dua = 0; };

inline double limit_scaling(double v) { return std::min(v < kMinScaling ? 1.0 : v, kMaxScaling); }

void rho_vec(const vec& l, const vec& u, double rho, vec& rv) {
    const size_t m = l.size();
    rv.assign(m, rho);
    for (size_t i = 0; i < m; ++i) {
        if (l[i] < -kInfThresh && u[i] > kInfThresh) rv[i] = kRhoMin;
        if (u[i] - l[i] < kRhoTol) rv[i] = kRhoEqFactor * rho;
    }
}

// lower triangle of [[P + sigma I, A'], [A, -diag(1/rho)]]
void build_kkt(const Coo& P, const Coo& A, double sigma, const vec& rv, std::vector<int>& ri, std::vector<int>& ci, vec& vi) {
    const int n = P.rows;
    ri.clear(); ci.clear(); vi.clear();
    for (size_t k = 0; k < P.v.size(); ++k)
        if (P.r[k] >= P.c[k]) { ri.push_back(P.r[k]); ci.push_back(P.c[k]); vi.push_back(P.v[k]); }
    for (int i = 0; i < n; ++i) { ri.push_back(i); ci.push_back(i); vi.push_back(sigma); }
    for (size_t k = 0; k < A.v.size(); ++k) { ri.push_back(n + A.r[k]); ci.push_back(A.c[k]); vi.push_back(A.v[k]); }
    for (int i = 0; i < A.rows; ++i) { ri.push_back(n + i); ci.push_back(n + i); vi.push_back(-1.0 / rv[i]); }
}

double inf_norm(const vec& a) { double s = 0; for (double v : a) s = std::max(s, fabs(v)); return s; }

// P: symmetric, BOTH triangles present.  x, y: outputs (unscaled).  Cold start.
OsqpResult osqp_solve(const Coo& P0, const vec& q0, const Coo& A0, const vec& l0, const vec& u0, const OsqpOpts& o,
                      vec& x_out, vec& y_out) {
    const int n = P0.rows, m = A0.rows;
    Coo P = P0, A = A0;
    vec q = q0, l(m), u(m);
    for (int i = 0; i < m; ++i) { l[i] = std::max(l0[i], -kInf); u[i] = std::min(u0[i], kInf); }
    // ---- modified Ruiz equilibration (OSQP scale_data) ----
    vec D(n, 1.0), E(m, 1.0), Dt(n), Et(m), cn(n), rn(m);
    double c = 1.0;
    for (int it = 0; it < o.scaling; ++it) {
        std::fill(cn.begin(), cn.end(), 0.0); std::fill(rn.begin(), rn.end(), 0.0);
        for (size_t k = 0; k < P.v.size(); ++k) cn[P.c[k]] = std::max(cn[P.c[k]], fabs(P.v[k]));
        for (size_t k = 0; k < A.v.size(); ++k) { cn[A.c[k]] = std::max(cn[A.c[k]], fabs(A.v[k])); rn[A.r[k]] = std::max(rn[A.r[k]], fabs(A.v[k])); }
        for (int j = 0; j < n; ++j) Dt[j] = 1.0 / sqrt(limit_scaling(cn[j]));
        for (int i = 0; i < m; ++i) Et[i] = 1.0 / sqrt(limit_scaling(rn[i]));
        for (size_t k = 0; k < P.v.size(); ++k) P.v[k] *= Dt[P.r[k]] * Dt[P.c[k]];
        for (size_t k = 0; k < A.v.size(); ++k) A.v[k] *= Et[A.r[k]] * Dt[A.c[k]];
        for (int j = 0; j < n; ++j) { q[j] *= Dt[j]; D[j] *= Dt[j]; }
        for (int i = 0; i < m; ++i) E[i] *= Et[i];
        std::fill(cn.begin(), cn.end(), 0.0);
        for (size_t k = 0; k < P.v.size(); ++k) cn[P.c[k]] = std::max(cn[P.c[k]], fabs(P.v[k]));
        double mean = 0; for (int j = 0; j < n; ++j) mean += cn[j]; mean /= std::max(1, n);
        const double cP = limit_scaling(mean), cq = limit_scaling(inf_norm(q));
        const double ct = 1.0 / std::max(cP, cq);
        for (double& v : P.v) v *= ct;
        for (double& v : q) v *= ct;
        c *= ct;
    }
    vec ls(m), us(m);
    for (int i = 0; i < m; ++i) { ls[i] = E[i] * l[i]; us[i] = E[i] * u[i]; }
    double rho = o.rho;
    vec rv;
    rho_vec(l, u, rho, rv);
    BandLdl kkt;
    std::vector<int> ri, ci; vec vi;
    build_kkt(P, A, o.sigma, rv, ri, ci, vi);
    OsqpResult res;
    if (!kkt.factor(n + m, ri, ci, vi, false)) { res.status = 3; return res; }
    res.n_fac = 1;
    vec x(n, 0.0), y(m, 0.0), z(m, 0.0), rhs(n + m), Ax, Px, Aty, zt(m);
    double pri = 1e300, dua = 1e300;
    int it = 0;
    res.status = 1;
    for (it = 1; it <= o.max_iter; ++it) {
        for (int j = 0; j < n; ++j) rhs[j] = o.sigma * x[j] - q[j];
        for (int i = 0; i < m; ++i) rhs[n + i] = z[i] - y[i] / rv[i];
        kkt.solve(rhs);
        for (int i = 0; i < m; ++i) zt[i] = z[i] + (rhs[n + i] - y[i]) / rv[i];
        for (int j = 0; j < n; ++j) x[j] = o.alpha * rhs[j] + (1 - o.alpha) * x[j];
        for (int i = 0; i < m; ++i) {
            const double zr = o.alpha * zt[i] + (1 - o.alpha) * z[i];
            const double zn = std::min(std::max(zr + y[i] / rv[i], ls[i]), us[i]);
            y[i] += rv[i] * (zr - zn);
            z[i] = zn;
        }
        const bool check = (it % o.check == 0), adapt = o.adapt_interval > 0 && (it % o.adapt_interval == 0);
        if (!check && !adapt) continue;
        A.mul(x, Ax); P.mul(x, Px); A.mulT(y, Aty);
        double npri = 0, ndua = 0;
        pri = 0; dua = 0;
        for (int i = 0; i < m; ++i) {
            pri = std::max(pri, fabs((Ax[i] - z[i]) / E[i]));
            npri = std::max(npri, std::max(fabs(Ax[i] / E[i]), fabs(z[i] / E[i])));
        }
        for (int j = 0; j < n; ++j) {
            dua = std::max(dua, fabs((Px[j] + q[j] + Aty[j]) / D[j]));
            ndua = std::max(ndua, std::max(fabs(Px[j] / D[j]), std::max(fabs(Aty[j] / D[j]), fabs(q[j] / D[j]))));
        }
        dua /= c; ndua /= c;
        if (check && pri <= o.eps_abs + o.eps_rel * npri && dua <= o.eps_abs + o.eps_rel * ndua) { res.status = 0; break; }
        if (adapt) {
            double rn2 = rho * sqrt((pri / std::max(npri, 1e-10)) / std::max(dua / std::max(ndua, 1e-10), 1e-10));
            rn2 = std::min(std::max(rn2, kRhoMin), kRhoMax);
            if (rn2 > rho * o.adapt_tol || rn2 < rho / o.adapt_tol) {
                rho = rn2;
                rho_vec(l, u, rho, rv);
                build_kkt(P, A, o.sigma, rv, ri, ci, vi);
                if (!kkt.factor(n + m, ri, ci, vi, true)) { res.status = 3; break; }
                ++res.n_fac;
            }
        }
    }
    res.iters = std::min(it, o.max_iter);
    // ---- polish ----
    if (o.polish && res.status == 0) {
        std::vector<int> act;
        vec bnd;
        for (int i = 0; i < m; ++i) {
            const bool low = (z[i] - ls[i]) < -y[i], upp = (us[i] - z[i]) < y[i];
            if (low || upp) { act.push_back(i); bnd.push_back(low ? ls[i] : us[i]); }
        }
        const int na = (int)act.size();
        std::vector<int> rowmap(m, -1);
        for (int k = 0; k < na; ++k) rowmap[act[k]] = k;
        Coo Ar; Ar.rows = na; Ar.cols = n;
        for (size_t k = 0; k < A.v.size(); ++k) if (rowmap[A.r[k]] >= 0) Ar.add(rowmap[A.r[k]], A.c[k], A.v[k]);
        ri.clear(); ci.clear(); vi.clear();
        for (size_t k = 0; k < P.v.size(); ++k) if (P.r[k] >= P.c[k]) { ri.push_back(P.r[k]); ci.push_back(P.c[k]); vi.push_back(P.v[k]); }
        for (int j = 0; j < n; ++j) { ri.push_back(j); ci.push_back(j); vi.push_back(o.delta); }
        for (size_t k = 0; k < Ar.v.size(); ++k) { ri.push_back(n + Ar.r[k]); ci.push_back(Ar.c[k]); vi.push_back(Ar.v[k]); }
        for (int k = 0; k < na; ++k) { ri.push_back(n + k); ci.push_back(n + k); vi.push_back(-o.delta); }
        BandLdl pk;
        if (pk.factor(n + na, ri, ci, vi, false)) {
            vec b(n + na), sol, r(n + na), t1, t2;
            for (int j = 0; j < n; ++j) b[j] = -q[j];
            for (int k = 0; k < na; ++k) b[n + k] = bnd[k];
            sol = b;
            pk.solve(sol);
            for (int rep = 0; rep < o.polish_refine; ++rep) {      // refinement against the unregularised system
                vec xs(sol.begin(), sol.begin() + n), ys(sol.begin() + n, sol.end());
                P.mul(xs, t1); Ar.mulT(ys, t2);
                for (int j = 0; j < n; ++j) r[j] = b[j] - (t1[j] + t2[j]);
                Ar.mul(xs, t1);
                for (int k = 0; k < na; ++k) r[n + k] = b[n + k] - t1[k];
                pk.solve(r);
                for (int k = 0; k < n + na; ++k) sol[k] += r[k];
            }
            vec xp(sol.begin(), sol.begin() + n), yp(m, 0.0), zp(m);
            for (int k = 0; k < na; ++k) yp[act[k]] = sol[n + k];
            A.mul(xp, Ax); P.mul(xp, Px); A.mulT(yp, Aty);
            double pri_p = 0, dua_p = 0;
            for (int i = 0; i < m; ++i) {
                zp[i] = std::min(std::max(Ax[i], ls[i]), us[i]);
                pri_p = std::max(pri_p, fabs((Ax[i] - zp[i]) / E[i]));
            }
            for (int j = 0; j < n; ++j) dua_p = std::max(dua_p, fabs((Px[j] + q[j] + Aty[j]) / D[j]));
            dua_p /= c;
            if ((pri_p < pri && dua_p < dua) || (pri_p < pri && dua < 1e-10) || (dua_p < dua && pri < 1e-10)) {
                x = xp; y = yp; z = zp; res.polished = 1;
            }
        }
    }
    x_out.resize(n); y_out.resize(m);
    for (int j = 0; j < n; ++j) x_out[j] = D[j] * x[j];
    for (int i = 0; i < m; ++i) y_out[i] = E[i] * y[i] / c;
    res.rho = rho; res.pri = pri; res.dua = dua;
    return res;
}

// ---------------------------------------------------------------------------------------------------------------
// hopper model
// ---------------------------------------------------------------------------------------------------------------
struct Params {
    int dyn = 3, N = 10, mpc_factor = 20, uref_aliased = 1;
    double mpc_dt = 0.02, sim_dt = 1e-3, m = 7.5, g = 9.807, mu = 1.0, fz_max = 206.0, z_min = 0.1, kf = 100.0;
    double J[9], Jinv[9], rh[3], Qd[12], Rd[6], tau_max[3];
};

void mat3(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += A[3 * i + k] * B[3 * k + j]; C[3 * i + j] = s; }
}
void mat3T_r(const double* A, const double* B, double* C) {   // C = A B'
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += A[3 * i + k] * B[3 * j + k]; C[3 * i + j] = s; }
}
void mv3(const double* A, const double* x, double* y) { for (int i = 0; i < 3; ++i) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2]; }
void mTv3(const double* A, const double* x, double* y) { for (int i = 0; i < 3; ++i) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2]; }
void cross(const double* a, const double* b, double* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
void inv3(const double* J, double* Ji) {
    const double a = J[0], b = J[1], c = J[2], d = J[3], e = J[4], f = J[5], g = J[6], h = J[7], i = J[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    Ji[0] = (e * i - f * h) / det; Ji[1] = (c * h - b * i) / det; Ji[2] = (b * f - c * e) / det;
    Ji[3] = (f * g - d * i) / det; Ji[4] = (a * i - c * g) / det; Ji[5] = (c * d - a * f) / det;
    Ji[6] = (d * h - e * g) / det; Ji[7] = (b * g - a * h) / det; Ji[8] = (a * e - b * d) / det;
}
void solve3(const double* J, const double* b, double* x) {     // np.linalg.solve: elimination with partial pivoting
    double a[3][4];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) a[i][j] = J[3 * i + j]; a[i][3] = b[i]; }
    for (int k = 0; k < 3; ++k) {
        int p = k;
        for (int i = k + 1; i < 3; ++i) if (fabs(a[i][k]) > fabs(a[p][k])) p = i;
        if (p != k) for (int j = 0; j < 4; ++j) std::swap(a[k][j], a[p][j]);
        for (int i = k + 1; i < 3; ++i) { const double f = a[i][k] / a[k][k]; for (int j = k; j < 4; ++j) a[i][j] -= f * a[k][j]; }
    }
    for (int i = 2; i >= 0; --i) { double s = a[i][3]; for (int j = i + 1; j < 3; ++j) s -= a[i][j] * x[j]; x[i] = s / a[i][i]; }
}
void quat_rotm(const double* q, double* R) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
void quat2euler(const double* q, double* rpy) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double nq = w * w + x * x + y * y + z * z, eps = 2.220446049250313e-16;
    if (nq < eps) { rpy[0] = rpy[1] = rpy[2] = 0; return; }
    const double s = 2.0 / nq, X = x * s, Y = y * s, Z = z * s;
    const double wX = w * X, wY = w * Y, wZ = w * Z, xX = x * X, xY = x * Y, xZ = x * Z, yY = y * Y, yZ = y * Z, zZ = z * Z;
    const double m00 = 1 - (yY + zZ), m10 = xY + wZ, m20 = xZ - wY, m21 = yZ + wX, m22 = 1 - (xX + yY), m11 = 1 - (xX + zZ), m12 = yZ - wX;
    const double cy = sqrt(m00 * m00 + m10 * m10);
    if (cy > 4 * eps) { rpy[0] = atan2(m21, m22); rpy[1] = atan2(-m20, cy); rpy[2] = atan2(m10, m00); }
    else { rpy[0] = atan2(-m12, m11); rpy[1] = atan2(-m20, cy); rpy[2] = 0; }
}
void convert(const double* X, double* x) {
    double R[9];
    quat_rotm(X + 3, R);
    x[0] = X[0]; x[1] = X[1]; x[2] = X[2];
    quat2euler(X + 3, x + 3);
    mv3(R, X + 7, x + 6);
    mv3(R, X + 10, x + 9);
}
void dynamics_ct(const Params& p, const double* X, const double* U, const double* pf, double* dX) {
    const double *pp = X, *q = X + 3, *v = X + 7, *w = X + 10;
    double R[9], Ft[3], Ftb[3], d[3], r[3], Fb[3], tt[3], Jw[3], wJw[3], wv[3], rhs[3];
    quat_rotm(q, R);
    Ft[0] = U[0]; Ft[1] = U[1]; Ft[2] = U[2] - p.g * p.m;
    mTv3(R, Ft, Ftb);
    for (int i = 0; i < 3; ++i) d[i] = pf[i] - pp[i];
    mTv3(R, d, r);
    for (int i = 0; i < 3; ++i) r[i] += p.rh[i];
    mTv3(R, U, Fb);
    cross(r, Fb, tt);
    mv3(R, v, dX);
    const double qw = q[0], *qv = q + 1;
    double qxw[3];
    cross(qv, w, qxw);
    dX[3] = 0.5 * -(qv[0] * w[0] + qv[1] * w[1] + qv[2] * w[2]);
    for (int i = 0; i < 3; ++i) dX[4 + i] = 0.5 * (qw * w[i] + qxw[i]);
    cross(w, v, wv);
    for (int i = 0; i < 3; ++i) dX[7 + i] = Ftb[i] / p.m - wv[i];
    mv3(p.J, w, Jw);
    cross(w, Jw, wJw);
    for (int i = 0; i < 3; ++i) rhs[i] = U[3 + i] + tt[i] - wJw[i];
    solve3(p.J, rhs, dX + 10);
}
void rk4(const Params& p, double* X, const double* U, const double* pf) {
    const double h = p.sim_dt;
    double f1[13], f2[13], f3[13], f4[13], t[13];
    dynamics_ct(p, X, U, pf, f1);
    for (int i = 0; i < 13; ++i) t[i] = X[i] + 0.5 * h * f1[i];
    dynamics_ct(p, t, U, pf, f2);
    for (int i = 0; i < 13; ++i) t[i] = X[i] + 0.5 * h * f2[i];
    dynamics_ct(p, t, U, pf, f3);
    for (int i = 0; i < 13; ++i) t[i] = X[i] + h * f3[i];
    dynamics_ct(p, t, U, pf, f4);
    for (int i = 0; i < 13; ++i) X[i] += (h / 6.0) * (f1[i] + 2 * f2[i] + 2 * f3[i] + f4[i]);
    const double nq = sqrt(X[3] * X[3] + X[4] * X[4] + X[5] * X[5] + X[6] * X[6]);
    for (int i = 3; i < 7; ++i) X[i] /= nq;
}

// Ad [N][12][12], Bd [N][12][6]
void gen_dt_dynamics(const Params& p, const double* x_guess, const double* pf, vec& Ad, vec& Bd) {
    const int N = p.N;
    const double dt = p.mpc_dt;
    Ad.assign((size_t)N * 144, 0.0); Bd.assign((size_t)N * 72, 0.0);
    for (int k = 0; k < N; ++k) {
        const double psi = x_guess[12 * k + 5], cs = cos(psi), sn = sin(psi);
        const double Rz[9] = {cs, sn, 0, -sn, cs, 0, 0, 0, 1};
        double d[3], rf[3], T1[9], Jw[9], JwRzT[9], B9[9];
        for (int i = 0; i < 3; ++i) d[i] = pf[3 * k + i] - x_guess[12 * k + i];
        mv3(Rz, d, rf);
        for (int i = 0; i < 3; ++i) rf[i] += p.rh[i];
        mat3(Rz, p.Jinv, T1); mat3T_r(T1, Rz, Jw); mat3T_r(Jw, Rz, JwRzT);
        double* A = &Ad[(size_t)k * 144]; double* B = &Bd[(size_t)k * 72];
        for (int i = 0; i < 12; ++i) A[12 * i + i] = 1.0;
        for (int i = 0; i < 3; ++i) A[12 * i + 6 + i] += dt;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[12 * (3 + i) + 9 + j] += Rz[3 * i + j] * dt;
        if (p.dyn == 3) {
            double rw[3];
            mTv3(Rz, rf, rw);
            const double H[9] = {0, -rw[2], rw[1], rw[2], 0, -rw[0], -rw[1], rw[0], 0};
            mat3(Jw, H, B9);
            for (int i = 0; i < 3; ++i) B[6 * (6 + i) + i] = (1.0 / p.m) * dt;
        } else {
            const double H[9] = {0, -rf[2], rf[1], rf[2], 0, -rf[0], -rf[1], rf[0], 0};
            mat3(JwRzT, H, B9);
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B[6 * (6 + i) + j] = (Rz[3 * j + i] / p.m) * dt;
        }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { B[6 * (9 + i) + j] = B9[3 * i + j] * dt; B[6 * (9 + i) + 3 + j] = JwRzT[3 * i + j] * dt; }
    }
}

// cvxpy-shaped full QP, rows in the order of oracle/hopper_oracle.py:build_qp_full
void build_qp_full(const Params& p, const double* x_in, const double* x_ref, const vec& Ad, const vec& Bd, const uint8_t* C,
                   Coo& P, vec& q, Coo& A, vec& l, vec& u) {
    const int N = p.N, nx = 12 * (N + 1), nu = 6 * N, nv = nx + nu;
    P = Coo(); P.rows = P.cols = nv; A = Coo(); A.cols = nv;
    q.assign(nv, 0.0); l.clear(); u.clear();
    const double gd = -p.g * p.mpc_dt;
    for (int k = 0; k < N; ++k) {
        const double kf = (k == N - 1) ? p.kf : 1.0, kuf = (k == N - 1) ? 0.0 : 1.0;
        const double ubar = p.uref_aliased ? (C[N - 1] ? 2 * p.m * p.g : 0.0) : (C[k] ? 2 * p.m * p.g : 0.0);
        const int ix = 12 * (k + 1), iu = nx + 6 * k;
        for (int i = 0; i < 12; ++i) { P.add(ix + i, ix + i, 2 * p.Qd[i] * kf); q[ix + i] = -2 * p.Qd[i] * kf * x_ref[12 * k + i]; }
        for (int i = 0; i < 6; ++i) P.add(iu + i, iu + i, 2 * p.Rd[i] * kuf);
        q[iu + 2] = -2 * p.Rd[2] * kuf * ubar;
    }
    int row = 0;
    auto bound = [&](double lo, double hi) { l.push_back(lo); u.push_back(hi); ++row; };
    for (int k = 0; k < N; ++k) {
        const int iu = nx + 6 * k, fx = iu, fy = iu + 1, fz = iu + 2;
        for (int a = 0; a < 3; ++a) { A.add(row, iu + 3 + a, 1.0); bound(-p.tau_max[a], p.tau_max[a]); }
        A.add(row, 12 * k + 2, 1.0); bound(p.z_min, kInf);
        for (int r = 0; r < 12; ++r) {
            A.add(row, 12 * (k + 1) + r, 1.0);
            for (int cc = 0; cc < 12; ++cc) { const double v = Ad[(size_t)k * 144 + 12 * r + cc]; if (v != 0.0) A.add(row, 12 * k + cc, -v); }
            for (int cc = 0; cc < 6; ++cc) { const double v = Bd[(size_t)k * 72 + 6 * r + cc]; if (v != 0.0) A.add(row, iu + cc, -v); }
            const double G = (r == 8) ? gd : 0.0;
            bound(G, G);
        }
        if (p.dyn == 2) { A.add(row, fy, 1.0); bound(0, 0); }
        if (!C[k]) {
            A.add(row, fx, 1.0); bound(0, 0);
            if (p.dyn == 3) { A.add(row, fy, 1.0); bound(0, 0); }
            A.add(row, fz, 1.0); bound(0, 0);
        } else {
            A.add(row, fx, 1.0); A.add(row, fz, -p.mu); bound(-kInf, 0);
            A.add(row, fx, -1.0); A.add(row, fz, -p.mu); bound(-kInf, 0);
            if (p.dyn == 3) {
                A.add(row, fy, 1.0); A.add(row, fz, -p.mu); bound(-kInf, 0);
                A.add(row, fy, -1.0); A.add(row, fz, -p.mu); bound(-kInf, 0);
            }
            A.add(row, fz, 1.0); bound(0, p.fz_max);
        }
    }
    for (int r = 0; r < 12; ++r) { A.add(row, r, 1.0); bound(x_in[r], x_in[r]); }
    A.rows = row;
}

void rollout_linear(const Params& p, const double* x_in, const double* U, const vec& Ad, const vec& Bd, double* X) {
    const int N = p.N;
    for (int i = 0; i < 12; ++i) X[i] = x_in[i];
    for (int k = 0; k < N; ++k)
        for (int r = 0; r < 12; ++r) {
            double s = (r == 8) ? -p.g * p.mpc_dt : 0.0;
            for (int cc = 0; cc < 12; ++cc) s += Ad[(size_t)k * 144 + 12 * r + cc] * X[12 * k + cc];
            for (int cc = 0; cc < 6; ++cc) s += Bd[(size_t)k * 72 + 6 * r + cc] * U[6 * k + cc];
            X[12 * (k + 1) + r] = s;
        }
}

struct Mpc {
    Params p;
    vec xval, uval;      // [(N+1)*12], [N*6]
    int last_iters = 0, last_status = 0, last_polished = 0, inaccurate = 0;
    OsqpOpts opts;
    bool solve(const double* x_in, const double* x_ref, const double* pf, const uint8_t* C, const double* x_guess) {
        vec Ad, Bd, q, l, u, x, y;
        Coo P, A;
        gen_dt_dynamics(p, x_guess, pf, Ad, Bd);
        build_qp_full(p, x_in, x_ref, Ad, Bd, C, P, q, A, l, u);
        const OsqpResult r = osqp_solve(P, q, A, l, u, opts, x, y);
        last_iters += r.iters; last_status = r.status; last_polished = r.polished;
        // max_iter reached: OSQP returns its last iterate and cvxpy keeps it (status "user_limit" carries a solution),
        // so the reference's loop goes on with it; only a failed factorisation has no point to return
        if (r.status == 1) ++inaccurate;
        if (r.status != 0 && r.status != 1) return false;
        const int N = p.N, nx = 12 * (N + 1);
        uval.assign(x.begin() + nx, x.end());
        xval.resize((size_t)(N + 1) * 12);
        rollout_linear(p, x_in, uval.data(), Ad, Bd, xval.data());
        return true;
    }
    bool mpcontrol(const double* x_in, const double* x_ref, const double* pf, const uint8_t* C, bool init) {
        const int N = p.N;
        vec xg((size_t)(N + 1) * 12, 0.0);
        last_iters = 0;
        for (int i = 0; i < 12; ++i) xg[i] = x_in[i];
        if (init) {
            for (int k = 0; k < N; ++k) for (int i = 0; i < 12; ++i) xg[12 * (k + 1) + i] = x_ref[12 * k + i];
            if (!solve(x_in, x_ref, pf, C, xg.data())) return false;
            xg = xval;
        } else {
            for (int k = 1; k < N; ++k) for (int i = 0; i < 12; ++i) xg[12 * k + i] = xval[12 * (k + 1) + i];
            for (int i = 0; i < 12; ++i) xg[12 * N + i] = xval[12 * N + i];
        }
        return solve(x_in, x_ref, pf, C, xg.data());
    }
};

}  // namespace

extern "C" {

// OSQP on dense inputs (P n x n symmetric, A m x n, row-major); zeros are dropped.  Returns status (0 solved,
// 1 max_iter, 3 factorisation failure); info = {iters, polished, n_fac}, dinfo = {rho, pri, dua}.
int ref_osqp_dense(int n, int m, const double* Pd, const double* q, const double* Ad, const double* l, const double* u,
                   double eps, int max_iter, int scaling, int polish, double* x, double* y, int* info, double* dinfo) {
    Coo P, A;
    P.rows = P.cols = n; A.rows = m; A.cols = n;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) if (Pd[(size_t)i * n + j] != 0.0) P.add(i, j, Pd[(size_t)i * n + j]);
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) if (Ad[(size_t)i * n + j] != 0.0) A.add(i, j, Ad[(size_t)i * n + j]);
    OsqpOpts o;
    o.eps_abs = o.eps_rel = eps; o.max_iter = max_iter; o.scaling = scaling; o.polish = polish;
    vec qv(q, q + n), lv(l, l + m), uv(u, u + m), xo, yo;
    const OsqpResult r = osqp_solve(P, qv, A, lv, uv, o, xo, yo);
    if (!xo.empty()) { memcpy(x, xo.data(), n * 8); memcpy(y, yo.data(), m * 8); }
    info[0] = r.iters; info[1] = r.polished; info[2] = r.n_fac;
    dinfo[0] = r.rho; dinfo[1] = r.pri; dinfo[2] = r.dua;
    return r.status;
}

// The reference's full QP for one stage set: returns dense P diag [nv], q [nv], A [m][nv], l, u [m]; m via *m_out
// (A, l, u must hold at least 40 N + 12 rows).
int ref_build_qp(int dyn, int N, const double* Qd, const double* Rd, const double* x_in, const double* x_ref,
                 const double* x_guess, const double* pf, const uint8_t* C, double* Pdiag, double* q, double* A,
                 double* l, double* u, int* m_out) {
    Params p;
    p.dyn = dyn; p.N = N;
    const double J[9] = {76148072.89e-9, 70089.52e-9, 2067970.36e-9, 70089.52e-9, 45477183.53e-9, -87045.58e-9, 2067970.36e-9, -87045.58e-9, 76287220.47e-9};
    memcpy(p.J, J, sizeof(J)); inv3(p.J, p.Jinv);
    p.rh[0] = -0.02663114 / 1000; p.rh[1] = -0.04435752 / 1000; p.rh[2] = -6.61082088 / 1000;
    memcpy(p.Qd, Qd, 96); memcpy(p.Rd, Rd, 48);
    p.tau_max[0] = 7.78; p.tau_max[1] = 7.78; p.tau_max[2] = 4.0;
    vec Ad, Bd, qv, lv, uv;
    Coo P, Am;
    gen_dt_dynamics(p, x_guess, pf, Ad, Bd);
    build_qp_full(p, x_in, x_ref, Ad, Bd, C, P, qv, Am, lv, uv);
    const int nv = 12 * (N + 1) + 6 * N, m = Am.rows;
    memset(Pdiag, 0, nv * 8);
    for (size_t k = 0; k < P.v.size(); ++k) Pdiag[P.r[k]] += P.v[k];
    memcpy(q, qv.data(), nv * 8);
    memset(A, 0, (size_t)m * nv * 8);
    for (size_t k = 0; k < Am.v.size(); ++k) A[(size_t)Am.r[k] * nv + Am.c[k]] += Am.v[k];
    memcpy(l, lv.data(), m * 8); memcpy(u, uv.data(), m * 8);
    *m_out = m;
    return 0;
}

// Closed loop of Runner.run (robotrunner.py:96-113) for one hopper on MPC-rate tables, the reference's per-tick recipe:
// convert -> mpcontrol (fresh full QP, OSQP cold start, eps 1e-5, polish) -> mpc_factor x rk4 with ZOH U[0].
//   X [13] in/out; xref_tab [T+N][12]; pf_tab [T+N+1][3]; C [T][N] (0/1 bytes); pf_switch [T]
//   stops after n_ticks or when budget_s seconds have passed (budget_s <= 0: no limit) or on a solver failure
//   X_log [(n_ticks+1)][13], U_log [n_ticks][6] optional; solve_us [n_ticks] optional: wall time of each mpcontrol;
//   iters [n_ticks] optional: OSQP iterations of the tick; n_inaccurate optional: solves that ended at max_iter (their
//   last iterate is used, as cvxpy does).  Returns the number of ticks done (negative: -1 - ticks when a
//   factorisation failed or the state went non-finite).
int ref_closed_loop(int dyn, int N, const double* Qd, const double* Rd, double* X, const double* xref_tab, const double* pf_tab,
                    const uint8_t* C, const uint8_t* pf_switch, int n_ticks, double budget_s, double eps, double* X_log,
                    double* U_log, double* solve_us, int* iters, int* n_inaccurate) {
    Mpc mpc;
    Params& p = mpc.p;
    p.dyn = dyn; p.N = N;
    const double J[9] = {76148072.89e-9, 70089.52e-9, 2067970.36e-9, 70089.52e-9, 45477183.53e-9, -87045.58e-9, 2067970.36e-9, -87045.58e-9, 76287220.47e-9};
    memcpy(p.J, J, sizeof(J)); inv3(p.J, p.Jinv);
    p.rh[0] = -0.02663114 / 1000; p.rh[1] = -0.04435752 / 1000; p.rh[2] = -6.61082088 / 1000;
    memcpy(p.Qd, Qd, 96); memcpy(p.Rd, Rd, 48);
    p.tau_max[0] = 7.78; p.tau_max[1] = 7.78; p.tau_max[2] = 4.0;
    mpc.opts.eps_abs = mpc.opts.eps_rel = eps;
    using clk = std::chrono::steady_clock;
    const auto t_start = clk::now();
    if (X_log) memcpy(X_log, X, 13 * 8);
    int t = 0;
    for (; t < n_ticks; ++t) {
        double x_in[12];
        convert(X, x_in);
        const auto t0 = clk::now();
        const bool ok = mpc.mpcontrol(x_in, xref_tab + (size_t)t * 12, pf_tab + (size_t)t * 3, C + (size_t)t * N, t == 0);
        const auto t1 = clk::now();
        if (solve_us) solve_us[t] = std::chrono::duration<double, std::micro>(t1 - t0).count();
        if (iters) iters[t] = mpc.last_iters;
        if (!ok) return -1 - t;
        double u0[6];
        memcpy(u0, mpc.uval.data(), 48);
        bool finite = true;
        for (int i = 0; i < 6; ++i) finite = finite && (fabs(u0[i]) < 1e6);
        if (!finite) return -1 - t;
        for (int i = 0; i < p.mpc_factor; ++i) rk4(p, X, u0, (i < pf_switch[t]) ? pf_tab + (size_t)t * 3 : pf_tab + (size_t)(t + 1) * 3);
        if (X_log) memcpy(X_log + (size_t)(t + 1) * 13, X, 13 * 8);
        if (U_log) memcpy(U_log + (size_t)t * 6, u0, 48);
        if (n_inaccurate) *n_inaccurate = mpc.inaccurate;
        if (budget_s > 0 && std::chrono::duration<double>(clk::now() - t_start).count() > budget_s) { ++t; break; }
    }
    if (n_inaccurate) *n_inaccurate = mpc.inaccurate;
    return t;
}

// rk4_normalized x nsteps and convert, for the cross-check against the numpy oracle
void ref_rk4(double* X, const double* U, const double* pf, int nsteps, double* x_out) {
    Params p;
    const double J[9] = {76148072.89e-9, 70089.52e-9, 2067970.36e-9, 70089.52e-9, 45477183.53e-9, -87045.58e-9, 2067970.36e-9, -87045.58e-9, 76287220.47e-9};
    memcpy(p.J, J, sizeof(J)); inv3(p.J, p.Jinv);
    p.rh[0] = -0.02663114 / 1000; p.rh[1] = -0.04435752 / 1000; p.rh[2] = -6.61082088 / 1000;
    for (int i = 0; i < nsteps; ++i) rk4(p, X, U, pf);
    if (x_out) convert(X, x_out);
}

}  // extern "C"
