"""QP solvers of the CPU oracle.  *** TEST INFRASTRUCTURE ONLY *** (see hopper_oracle.py header)

Three independent pieces:
  * ``osqp_solve``      -- restatement of the OSQP algorithm (the solver the reference reaches through
                           cvxpy at mpc_cvx_euler_3f.py:156-157).  OSQP itself is a third-party
                           dependency absent from /root/reference and from this image; the reference
                           pins no version (README.md:45).  Restated from the published algorithm
                           (Stellato et al., "OSQP: an operator splitting solver for quadratic
                           programs", Math. Prog. Comp. 2020) with the 0.6.x defaults and the settings
                           cvxpy passes (eps_abs=eps_rel=1e-5, max_iter=10000, polish on) [EXT].
                           One deliberate deviation: ``adaptive_rho_interval`` is a fixed iteration
                           count (OSQP derives it from measured wall time, which is not reproducible).
  * ``exact_qp``        -- solver-independent exact optimum: Mehrotra primal-dual interior point on
                           the inequality form, then an active-set KKT solve with verification.
  * ``kkt_certificate`` -- residuals of the KKT conditions for  min 1/2 x'Px+q'x, l<=Ax<=u.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

INF = 1e30
_INF_THRESH = INF * 1e-4       # OSQP: OSQP_INFTY * MIN_SCALING
_MIN_SCALING, _MAX_SCALING = 1e-4, 1e4
_RHO_MIN, _RHO_MAX = 1e-6, 1e6
_RHO_TOL = 1e-4
_RHO_EQ_FACTOR = 1e3


# ---------------------------------------------------------------------------------------------
def kkt_certificate(P, q, A, l, u, x, y):
    """Solver-independent optimality certificate.

    Returns dict(stat, prim, comp, sign): inf-norms of stationarity P x + q + A'y, primal
    infeasibility, complementarity (y+ * (u - Ax), y- * (Ax - l)) and dual sign violations
    (y>0 on a row with u infinite, y<0 with l infinite)."""
    Ax = A @ x
    stat = np.max(np.abs(P @ x + q + A.T @ y)) if len(x) else 0.0
    prim = max(0.0, np.max(np.maximum(l - Ax, 0.0)), np.max(np.maximum(Ax - u, 0.0)))
    yp, ym = np.maximum(y, 0.0), np.maximum(-y, 0.0)
    fin_u, fin_l = u < _INF_THRESH, l > -_INF_THRESH
    comp = 0.0
    if fin_u.any():
        comp = max(comp, np.max(np.abs(yp[fin_u] * (u[fin_u] - Ax[fin_u]))))
    if fin_l.any():
        comp = max(comp, np.max(np.abs(ym[fin_l] * (Ax[fin_l] - l[fin_l]))))
    sign = 0.0
    if (~fin_u).any():
        sign = max(sign, np.max(yp[~fin_u]))
    if (~fin_l).any():
        sign = max(sign, np.max(ym[~fin_l]))
    return dict(stat=float(stat), prim=float(prim), comp=float(comp), sign=float(sign))


# ---------------------------------------------------------------------------------------------
def _limit_scaling(v):
    v = np.where(v < _MIN_SCALING, 1.0, v)
    return np.minimum(v, _MAX_SCALING)


def _ruiz(P, q, A, l, u, iters=10):
    """OSQP scale_data(): modified Ruiz equilibration of [[P, A'],[A, 0]] + cost normalisation."""
    n, m = P.shape[0], A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    P, A, q = P.copy(), A.copy(), q.copy()
    for _ in range(iters):
        colP = np.abs(P).max(axis=0) if n else np.zeros(0)
        colA = np.abs(A).max(axis=0) if m else np.zeros(n)
        Dt = 1.0 / np.sqrt(_limit_scaling(np.maximum(colP, colA)))
        Et = 1.0 / np.sqrt(_limit_scaling(np.abs(A).max(axis=1))) if m else np.ones(0)
        P = Dt[:, None] * P * Dt[None, :]
        A = Et[:, None] * A * Dt[None, :]
        q = Dt * q
        D *= Dt
        E *= Et
        cP = _limit_scaling(np.array([np.abs(P).max(axis=0).mean()]))[0]
        cq = _limit_scaling(np.array([np.abs(q).max()]))[0]
        ct = 1.0 / max(cP, cq)
        P *= ct
        q *= ct
        c *= ct
    return P, q, A, E * l, E * u, D, E, c


def _rho_vec(l, u, rho):
    r = np.full(l.shape, rho)
    r[(l < -_INF_THRESH) & (u > _INF_THRESH)] = _RHO_MIN
    r[(u - l) < _RHO_TOL] = _RHO_EQ_FACTOR * rho
    return r


class _KKT:
    """Quasi-definite KKT [[P+sigma I, A'],[A, -diag(1/rho)]] factorised with sparse LU."""

    def __init__(self, P, A, sigma, rho_vec):
        n, m = P.shape[0], A.shape[0]
        K = sp.bmat([[sp.csc_matrix(P) + sigma * sp.identity(n), sp.csc_matrix(A).T],
                     [sp.csc_matrix(A), -sp.diags(1.0 / rho_vec)]], format="csc")
        self.lu = spla.splu(K)
        self.n = n

    def solve(self, rx, rz):
        s = self.lu.solve(np.concatenate((rx, rz)))
        return s[:self.n], s[self.n:]


def osqp_solve(P, q, A, l, u, eps_abs=1e-5, eps_rel=1e-5, max_iter=10000, rho=0.1, sigma=1e-6,
               alpha=1.6, scaling=10, check_termination=25, adaptive_rho=True,
               adaptive_rho_interval=50, adaptive_rho_tolerance=5.0, polish=True, delta=1e-6,
               polish_refine_iter=3, eps_prim_inf=1e-4, x0=None, y0=None):
    """OSQP algorithm (see module header).  Returns dict(x, y, status, iters, rho, polished, n_fac)."""
    P = np.asarray(P, float); A = np.asarray(A, float)
    n, m = P.shape[0], A.shape[0]
    l = np.maximum(np.asarray(l, float), -INF); u = np.minimum(np.asarray(u, float), INF)
    if scaling:
        Ps, qs, As, ls, us, D, E, c = _ruiz(P, q, A, l, u, scaling)
    else:
        Ps, qs, As, ls, us, D, E, c = P, np.asarray(q, float), A, l, u, np.ones(n), np.ones(m), 1.0
    Dinv, Einv = 1.0 / D, 1.0 / E
    rv = _rho_vec(l, u, rho)
    kkt = _KKT(Ps, As, sigma, rv)
    n_fac = 1
    x = np.zeros(n) if x0 is None else np.asarray(x0, float) / D
    y = np.zeros(m) if y0 is None else np.asarray(y0, float) * c / E
    z = As @ x
    status = "max_iter"
    it = 0
    pri = dua = np.inf
    for it in range(1, max_iter + 1):
        xt, nu = kkt.solve(sigma * x - qs, z - y / rv)
        zt = z + (nu - y) / rv
        x = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        z_new = np.clip(zr + y / rv, ls, us)
        y = y + rv * (zr - z_new)
        z = z_new
        check = (it % check_termination == 0)
        adapt = adaptive_rho and adaptive_rho_interval and (it % adaptive_rho_interval == 0)
        if not (check or adapt):
            continue
        Ax, Px, Aty = As @ x, Ps @ x, As.T @ y
        pri = np.max(np.abs(Einv * (Ax - z))) if m else 0.0
        dua = np.max(np.abs(Dinv * (Px + qs + Aty))) / c
        npri = max(np.max(np.abs(Einv * Ax)), np.max(np.abs(Einv * z))) if m else 0.0
        ndua = max(np.max(np.abs(Dinv * Px)), np.max(np.abs(Dinv * Aty)), np.max(np.abs(Dinv * qs))) / c
        if check and pri <= eps_abs + eps_rel * npri and dua <= eps_abs + eps_rel * ndua:
            status = "solved"
            break
        if check and m:
            # primal infeasibility certificate on delta_y is omitted: the hot path flags the only
            # infeasible case (height rows k=0,1, App. D2) before solving.
            pass
        if adapt:
            rho_new = rho * np.sqrt((pri / max(npri, 1e-10)) / max(dua / max(ndua, 1e-10), 1e-10))
            rho_new = min(max(rho_new, _RHO_MIN), _RHO_MAX)
            if rho_new > rho * adaptive_rho_tolerance or rho_new < rho / adaptive_rho_tolerance:
                rho = rho_new
                rv = _rho_vec(l, u, rho)
                kkt = _KKT(Ps, As, sigma, rv)
                n_fac += 1
    polished = False
    if polish and status == "solved":
        low = (z - ls) < -y
        upp = (us - z) < y
        act = low | upp
        Ar = As[act]
        na = Ar.shape[0]
        b = np.where(low, ls, us)[act]
        Kreg = np.block([[Ps + delta * np.eye(n), Ar.T], [Ar, -delta * np.eye(na)]])
        K0 = np.block([[Ps, Ar.T], [Ar, np.zeros((na, na))]])
        rhs = np.concatenate((-qs, b))
        lu = sla.lu_factor(Kreg)
        sol = sla.lu_solve(lu, rhs)
        for _ in range(polish_refine_iter):
            sol = sol + sla.lu_solve(lu, rhs - K0 @ sol)
        xp = sol[:n]
        yp = np.zeros(m); yp[act] = sol[n:]
        zp = np.clip(As @ xp, ls, us)
        pri_p = np.max(np.abs(Einv * (As @ xp - zp))) if m else 0.0
        dua_p = np.max(np.abs(Dinv * (Ps @ xp + qs + As.T @ yp))) / c
        if (pri_p < pri and dua_p < dua) or (pri_p < pri and dua < 1e-10) or (dua_p < dua and pri < 1e-10):
            x, y, z, polished = xp, yp, zp, True
    return dict(x=D * x, y=E * y / c, status=status, iters=it, rho=rho, polished=polished,
                n_fac=n_fac, pri=pri, dua=dua)


# ---------------------------------------------------------------------------------------------
def _active_set_solve(H, g, A, b, fixed_tol=0.0):
    """min 1/2 x'Hx + g'x s.t. A x = b  (H PD).  Range-space solve with refinement."""
    n, na = H.shape[0], A.shape[0]
    if na == 0:
        return -sla.cho_solve(sla.cho_factor(H), g), np.zeros(0)
    K = np.block([[H, A.T], [A, np.zeros((na, na))]])
    rhs = np.concatenate((-g, b))
    try:
        lu = sla.lu_factor(K + np.diag(np.concatenate((np.zeros(n), -1e-12 * np.ones(na)))))
    except Exception:
        return None, None
    sol = sla.lu_solve(lu, rhs)
    for _ in range(4):
        sol = sol + sla.lu_solve(lu, rhs - K @ sol)
    return sol[:n], sol[n:]


def _ipm(H, g, G, h, tol=1e-11, max_iter=100):
    """Mehrotra predictor-corrector primal-dual interior point for  min 1/2 x'Hx + g'x, G x <= h
    (H positive definite).  Returns (x, lam, s, iters)."""
    n, ni = H.shape[0], G.shape[0]
    if ni == 0:
        return -sla.cho_solve(sla.cho_factor(H), g), np.zeros(0), np.zeros(0), 0
    x = sla.cho_solve(sla.cho_factor(H + G.T @ G), -g + G.T @ h)
    s = h - G @ x
    s = np.maximum(s + max(-1.5 * s.min(), 0.0), 1e-2)
    lam = np.ones(ni)
    gs, hs = max(1.0, np.abs(g).max()), max(1.0, np.abs(h).max())

    def step_len(v, dv):
        neg = dv < 0
        return min(1.0, float(np.min(-v[neg] / dv[neg]))) if neg.any() else 1.0

    it = 0
    for it in range(max_iter):
        rd = H @ x + g + G.T @ lam
        rp = G @ x + s - h
        mu_ = (s @ lam) / ni
        if np.abs(rd).max() < tol * gs and np.abs(rp).max() < tol * hs and mu_ < tol:
            break
        W = lam / s
        Hs = H + G.T @ (W[:, None] * G)
        try:
            cf = sla.cho_factor(Hs)
        except sla.LinAlgError:
            break

        def newton(rc):
            r1 = -rd - G.T @ ((lam * rp - rc) / s)
            dx = sla.cho_solve(cf, r1)
            dx += sla.cho_solve(cf, r1 - Hs @ dx)
            ds = -rp - G @ dx
            return dx, ds, -(rc + lam * ds) / s

        dx, ds, dlam = newton(s * lam)
        a_aff = min(step_len(s, ds), step_len(lam, dlam))
        mu_aff = ((s + a_aff * ds) @ (lam + a_aff * dlam)) / ni
        sig = (mu_aff / mu_) ** 3 if mu_ > 0 else 0.0
        dx, ds, dlam = newton(s * lam + ds * dlam - sig * mu_)
        a = min(1.0, 0.99 * min(step_len(s, ds), step_len(lam, dlam)))
        x = x + a * dx; s = s + a * ds; lam = lam + a * dlam
    return x, lam, s, it


def exact_qp(H, g, A, l, u, tol=1e-11, max_iter=100):
    """Exact optimum of the strictly convex QP  min 1/2 x'Hx + g'x, l <= Ax <= u  whose first n rows of A
    are the identity (the condensed layout of hopper_oracle.build_qp_condensed).

    Variables with l == u on their identity row are eliminated, a Mehrotra interior point runs on the
    remaining inequality-only problem, then an active-set KKT solve is accepted only if it improves the
    KKT certificate.  Independent of the ADMM code path on purpose.  Returns dict(x, y, cert, ok)."""
    H = np.asarray(H, float); A = np.asarray(A, float)
    l = np.asarray(l, float); u = np.asarray(u, float); g = np.asarray(g, float)
    n, m = H.shape[0], A.shape[0]
    assert m >= n and np.array_equal(A[:n], np.eye(n)), "identity box rows expected first"
    fixed = (u[:n] - l[:n]) < 1e-12
    fr = ~fixed
    xfix = np.where(fixed, l[:n], 0.0)
    Hr = H[np.ix_(fr, fr)]
    gr = g[fr] + H[np.ix_(fr, fixed)] @ xfix[fixed]
    off = A[:, fixed] @ xfix[fixed]
    Ar_all, lr_all, ur_all = A[:, fr], l - off, u - off
    rows = np.ones(m, bool); rows[:n] = fr
    rows &= np.abs(Ar_all).sum(axis=1) > 0
    ridx = np.where(rows)[0]
    Ar, lr, ur = Ar_all[rows], lr_all[rows], ur_all[rows]
    # rows without support but with violated constant bounds -> infeasible
    dead = ~rows; dead[:n] = False
    infeasible_const = bool(np.any(lr_all[dead] > 1e-12) or np.any(ur_all[dead] < -1e-12))
    fu, fl = ur < _INF_THRESH, lr > -_INF_THRESH
    G = np.vstack((Ar[fu], -Ar[fl]))
    h = np.concatenate((ur[fu], -lr[fl]))
    x_r, lam, s, it = _ipm(Hr, gr, G, h, tol, max_iter)
    nu_u = int(fu.sum())
    yr = np.zeros(Ar.shape[0])
    yr[np.where(fu)[0]] += lam[:nu_u]
    yr[np.where(fl)[0]] -= lam[nu_u:]

    def full(xr, yrow):
        x = xfix.copy(); x[fr] = xr
        y = np.zeros(m); y[ridx] = yrow
        grad = H @ x + g + A.T @ y
        y[:n][fixed] = -grad[fixed]      # multiplier of the eliminated variable's equality row
        return x, y

    cands = [full(x_r, yr)]
    # active-set refinement on the reduced problem
    act_u = np.zeros(Ar.shape[0], bool); act_l = np.zeros(Ar.shape[0], bool)
    act_u[np.where(fu)[0]] = lam[:nu_u] > s[:nu_u]
    act_l[np.where(fl)[0]] = lam[nu_u:] > s[nu_u:]
    act = act_u | act_l
    if act.any():
        Aa = Ar[act]; b = np.where(act_u, ur, lr)[act]
        Qr, Rr, piv = sla.qr(Aa.T, mode="economic", pivoting=True)
        rank = int(np.sum(np.abs(np.diag(Rr)) > 1e-10 * max(1.0, abs(Rr[0, 0]))))
        keep = np.sort(piv[:rank])
        xa, ya = _active_set_solve(Hr, gr, Aa[keep], b[keep])
        if xa is not None:
            y2 = np.zeros(Ar.shape[0])
            y2[np.where(act)[0][keep]] = ya
            if np.all(y2[act_u] >= -1e-9) and np.all(y2[act_l] <= 1e-9):
                cands.append(full(xa, y2))
    else:
        xa, _ = _active_set_solve(Hr, gr, Ar[:0], lr[:0])
        cands.append(full(xa, np.zeros(Ar.shape[0])))
    best, best_cert = None, None
    for x, y in cands:
        c = kkt_certificate(H, g, A, l, u, x, y)
        score = max(c["stat"] / max(1.0, np.abs(g).max()), c["prim"], c["sign"], c["comp"])
        if best is None or score < best[0]:
            best, best_cert = (score, x, y), c
    gs = max(1.0, float(np.abs(g).max()))
    ok = (not infeasible_const) and best_cert["stat"] < 1e-9 * gs and best_cert["prim"] < 1e-9 \
        and best_cert["sign"] < 1e-7 and best_cert["comp"] < 1e-6 * gs
    return dict(x=best[1], y=best[2], cert=best_cert, ok=bool(ok), ipm_iters=it)
