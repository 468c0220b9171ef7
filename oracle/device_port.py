"""numpy port of the PRODUCT's device algorithms (csrc/hmpc_qp.cuh).  *** TEST INFRASTRUCTURE ONLY ***

Same arithmetic as the CUDA kernels, dense numpy instead of per-CTA shared memory.  It exists to
  * develop / tune the device algorithms on the CPU (this container has no GPU),
  * pin their host-visible behaviour (status codes, iteration counts, ADMM iterates at a fixed
    iteration count) in the test-suite: the GPU ADMM iterate is compared against ``admm_solve`` here
    (same schedule -> same iterate up to rounding), which is what "OSQP at matched eps" can mean when
    the OSQP binary is not installable.
The *checker* for optimality is not this file but ``qp_solvers.exact_qp`` + ``kkt_certificate``.

Problem form:  min 1/2 x'Hx + g'x,  lo <= A x <= hi,  rows = [I (n) ; friction slots (4N) ; height slots (N)].
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

INF_T = 1e26
RHO_MIN, RHO_MAX = 1e-6, 1e6
ST_SOLVED, ST_MAX_ITER, ST_INFEASIBLE, ST_NON_FINITE, ST_INEXACT = 0, 1, 2, 3, 4


def _factor(H, A, wts, dadd, pin):
    """Cholesky of K = H + dadd I + A' diag(wts) A with pinned variables removed (identity rows)."""
    n = H.shape[0]
    K = H + dadd * np.eye(n) + A[n:].T @ (wts[n:, None] * A[n:]) + np.diag(wts[:n])
    K[pin, :] = 0.0
    K[:, pin] = 0.0
    K[pin, pin] = 1.0
    return sla.cho_factor(K)


def shift_warm_start(x_prev, code_prev, N):
    """Warm start from the previous tick (device: mpc_hopper): stage k takes the previous stage k+1, except
    the last two stages, which keep their own previous solution / active-set pattern."""
    n = 6 * N
    x0 = np.array(x_prev, float)
    c = np.array(code_prev)
    for k in range(N - 2):
        x0[6 * k:6 * k + 6] = x_prev[6 * (k + 1):6 * (k + 1) + 6]
        c[6 * k:6 * k + 6] = code_prev[6 * (k + 1):6 * (k + 1) + 6]
        c[n + 4 * k:n + 4 * k + 4] = code_prev[n + 4 * (k + 1):n + 4 * (k + 1) + 4]
        c[n + 4 * N + k] = code_prev[n + 4 * N + k + 1]
    return x0, c


def _kkt_factor(H, A, pin, rows, eps):
    """Quasi-definite KKT of the equality-constrained QP on the unpinned variables F and the active
    general rows: [[H_FF, G'], [G, -eps I]] (device: signed LDL' in shared memory)."""
    F = np.where(~pin)[0]
    G = A[np.ix_(rows, F)]
    nF, ng = len(F), len(rows)
    K = np.zeros((nF + ng, nF + ng))
    K[:nF, :nF] = H[np.ix_(F, F)]
    K[nF:, :nF] = G
    K[:nF, nF:] = G.T
    # relative regularisation of the row block (device: LinSys::factor): the diagonal of the Schur
    # complement -G H_FF^-1 G' is scaled by (1 + eps); rows without an unpinned variable are decoupled
    if ng:
        S = G @ np.linalg.solve(H[np.ix_(F, F)], G.T) if nF else np.zeros((ng, ng))
        d = np.where(np.abs(G).sum(axis=1) > 0, eps * np.diag(S), 1.0)
        K[nF:, nF:] = -np.diag(d)
    return sla.lu_factor(K), F


def polish_verified(H, g, A, lo, hi, x, code, eps=1e-9, tol=1e-9, retries=8, max_refine=6):
    """Verified primal-dual active-set refinement (device: polish_verified).

    ``code`` (+1 upper active, -1 lower active, 0 inactive) is the active-set guess.  Box-active and
    a-priori fixed variables are pinned exactly; active friction/height rows enter a quasi-definite KKT
    system [[H_FF, G'],[G, -eps I]] that is factorised once per trial and applied in correction form
    (iterative refinement removes the eps perturbation) until the residuals stagnate.  The KKT conditions
    of the ORIGINAL QP are then checked; on failure wrong-signed rows are released and violated rows
    activated, at most ``retries`` times.  Returns ((x, y, code), nfac) or (None, nfac)."""
    n, m = H.shape[0], A.shape[0]
    fixed = (hi[:n] - lo[:n]) < 1e-12
    apriori = np.concatenate((fixed, np.zeros(m - n, bool)))
    code = np.where((code > 0) & (hi > INF_T), 0, code)
    code = np.where((code < 0) & (lo < -INF_T), 0, code)
    code[:n][fixed] = 0
    xp = np.array(x, float)
    nfac = 0
    for _trial in range(retries + 1):
        bnd = np.where(code < 0, lo, hi)
        pin = fixed | (code[:n] != 0)
        xp = np.where(pin, np.where(fixed, lo[:n], bnd[:n]), xp)
        ga = code != 0
        ga[:n] = False
        rows = np.where(ga)[0]
        nF, ng = int((~pin).sum()), len(rows)
        if nF + ng > (8 * n) // 6:      # device capacity: 8N unknowns
            return None, nfac
        lu, F = _kkt_factor(H, A, pin, rows, eps)
        nfac += 1
        mul = np.zeros(m)
        prev = np.inf
        for k in range(max_refine):
            Hx_, Aty_, Ax_ = H @ xp, A.T @ mul, A @ xp
            rd = -(Hx_ + g + Aty_)[F]
            rp = (bnd - Ax_)[rows]
            res = max(np.abs(rd).max(initial=0.0), np.abs(rp).max(initial=0.0))
            rel = max((np.abs(rd) / ((np.abs(Hx_) + np.abs(g) + np.abs(Aty_))[F] + 1e-300)).max(initial=0.0),
                      (np.abs(rp) / ((np.abs(bnd) + np.abs(Ax_))[rows] + 1e-300)).max(initial=0.0))
            # stop when every row's residual sits at its rounding level or the residual has stopped contracting
            if k == 1 and ng == 0 and res <= 1e-7 * prev:      # exact system, already at working accuracy
                break
            if k >= 1 and (rel <= 1e-12 or (k >= 2 and res > 0.25 * prev)):
                break
            prev = res
            sol = sla.lu_solve(lu, np.concatenate((rd, rp)))
            xp[F] += sol[:nF]
            mul[rows] += sol[nF:]
        Hx = H @ xp
        Aty = A.T @ mul
        G = Hx + g + Aty
        stat = np.abs(G[~pin]).max() if (~pin).any() else 0.0
        scale = max(1.0, np.abs(Hx).max(), np.abs(g).max(), np.abs(Aty).max())
        lam = mul.copy()
        lam[:n] = np.where(pin, -G, 0.0)
        stol = tol * max(scale, np.abs(lam).max())
        ax = A @ xp
        act = code != 0
        bad = (not np.isfinite(stat)) or stat > 1e-10 * scale
        wrong = act & (((code > 0) & (lam < -stol)) | ((code < 0) & (lam > stol)))
        vl = (~act) & ~apriori & (lo - ax > tol * (1 + np.abs(lo)))
        vu = (~act) & ~apriori & (ax - hi > tol * (1 + np.abs(hi)))
        eqbad = ga & (np.abs(ax - bnd) > tol * (1 + np.abs(bnd)))
        if not (bad or wrong.any() or vl.any() or vu.any() or eqbad.any()):
            y = np.where(act | apriori, lam, 0.0)
            return (xp, y, code), nfac
        if not np.isfinite(stat):
            return None, nfac
        # release wrong-signed rows first; violated rows are only activated by a trial without wrong-signed rows
        ncode = code.copy()
        ncode[wrong] = 0
        if not wrong.any():
            ncode[vl] = -1
            ncode[vu] = 1
        if (ncode == code).all():
            return None, nfac
        code = ncode
    return None, nfac


def ipm_solve(H, g, A, lo, hi, tol=1e-9, max_iter=40, s0=0.1, support=None):
    """Mehrotra predictor-corrector interior point in the device's row layout (device: ipm_solve).

    Every row keeps an upper and a lower slack/multiplier pair; sides at +-inf and rows of a-priori
    fixed variables are masked.  Start: one Newton step of the quadratic penalty towards the row
    mid-points.  Returns (x, y, code, iters, converged)."""
    n, m = H.shape[0], A.shape[0]
    fixed = (hi[:n] - lo[:n]) < 1e-12
    fu = hi < INF_T
    fl = lo > -INF_T
    fu[:n] &= ~fixed
    fl[:n] &= ~fixed
    if support is not None:
        fu &= support
        fl &= support
    ni = int(fu.sum() + fl.sum())
    x = np.where(fixed, lo[:n], 0.0)
    if ni == 0:
        cf = _factor(H, A, np.zeros(m), 0.0, fixed)
        r = -(H @ x + g); r[fixed] = 0.0
        return x + sla.cho_solve(cf, r), np.zeros(m), np.zeros(m, int), 0, True
    w0 = np.where(fu | fl, 1.0, 0.0)
    mid = np.where(fu & fl, 0.5 * (lo + hi), np.where(fu, hi - 1.0, np.where(fl, lo + 1.0, 0.0)))
    cf = _factor(H, A, w0, 0.0, fixed)
    r = -(H @ x + g) - A.T @ (w0 * (A @ x - mid)); r[fixed] = 0.0
    x = x + sla.cho_solve(cf, r)
    ax = A @ x
    su = np.where(fu, hi - ax, 1.0)
    sl = np.where(fl, ax - lo, 1.0)
    smin = min(su[fu].min() if fu.any() else 1.0, sl[fl].min() if fl.any() else 1.0)
    shift = max(0.0, -1.5 * smin)
    su = np.maximum(su + shift, s0)
    sl = np.maximum(sl + shift, s0)
    lu = np.where(fu, s0, 0.0)
    ll = np.where(fl, s0, 0.0)
    gs = max(1.0, np.abs(g).max())
    conv = False
    it = 0

    def ratio(v, dv, mask):
        neg = mask & (dv < 0)
        return min(1.0, float(np.min(-v[neg] / dv[neg]))) if neg.any() else 1.0

    for it in range(max_iter + 1):
        ax = A @ x
        rd = H @ x + g + A.T @ (lu - ll)
        rd[fixed] = 0.0
        rpu = np.where(fu, ax + su - hi, 0.0)
        rpl = np.where(fl, -ax + sl + lo, 0.0)
        mu = (su @ lu + sl @ ll) / ni
        if np.abs(rd).max() < tol * gs and max(np.abs(rpu).max(), np.abs(rpl).max()) < tol and mu < tol:
            conv = True
            break
        if it == max_iter or not np.isfinite(mu):
            break
        w = np.where(fu, lu / su, 0.0) + np.where(fl, ll / sl, 0.0)
        try:
            cf = _factor(H, A, w, 0.0, fixed)
        except np.linalg.LinAlgError:
            break

        def newton(rcu, rcl):
            tu = np.where(fu, (lu * rpu - rcu) / su, 0.0)
            tl = np.where(fl, (ll * rpl - rcl) / sl, 0.0)
            r1 = -rd - A.T @ (tu - tl)
            r1[fixed] = 0.0
            dx = sla.cho_solve(cf, r1)
            adx = A @ dx
            dsu = -rpu - adx
            dsl = -rpl + adx
            dlu = np.where(fu, -(rcu + lu * dsu) / su, 0.0)
            dll = np.where(fl, -(rcl + ll * dsl) / sl, 0.0)
            return dx, dsu, dsl, dlu, dll

        dx, dsu, dsl, dlu, dll = newton(su * lu, sl * ll)
        a = min(ratio(su, dsu, fu), ratio(sl, dsl, fl), ratio(lu, dlu, fu), ratio(ll, dll, fl))
        mu_aff = (np.where(fu, (su + a * dsu) * (lu + a * dlu), 0.0).sum()
                  + np.where(fl, (sl + a * dsl) * (ll + a * dll), 0.0).sum()) / ni
        sig = (mu_aff / mu) ** 3
        dx, dsu, dsl, dlu, dll = newton(su * lu + dsu * dlu - sig * mu, sl * ll + dsl * dll - sig * mu)
        a = min(1.0, 0.99 * min(ratio(su, dsu, fu), ratio(sl, dsl, fl), ratio(lu, dlu, fu), ratio(ll, dll, fl)))
        x = x + a * dx
        su = su + a * dsu; sl = sl + a * dsl; lu = lu + a * dlu; ll = ll + a * dll
    code = np.where(fu & (lu > su), 1, np.where(fl & (ll > sl), -1, 0))
    return x, lu - ll, code, it, conv


def solve_exact(H, g, A, lo, hi, warm=None, tol=1e-9, eps=1e-9, retries=8, ipm_tol=1e-9, support=None):
    """The product's default solver (device: solve_exact): warm-started verified active-set refinement,
    interior-point fallback, verified polish.  Returns (x, y, code, info)."""
    info = dict(status=ST_MAX_ITER, path="", nfac=0, ipm_iters=0)
    if warm is not None:
        r, nf = polish_verified(H, g, A, lo, hi, warm[0], warm[1], eps, tol, retries)
        info["nfac"] += nf
        if r is not None:
            info.update(status=ST_SOLVED, path="warm")
            return r[0], r[1], r[2], info
    x, y, code, it, conv = ipm_solve(H, g, A, lo, hi, ipm_tol, support=support)
    info["ipm_iters"] = it
    info["nfac"] += it + 1
    if not np.all(np.isfinite(x)):
        info.update(status=ST_NON_FINITE, path="ipm")
        return x, y, code, info
    r, nf = polish_verified(H, g, A, lo, hi, x, code, eps, tol, retries)
    info["nfac"] += nf
    if r is not None:
        info.update(status=ST_SOLVED, path="ipm+polish")
        return r[0], r[1], r[2], info
    info.update(status=ST_INEXACT if conv else ST_MAX_ITER, path="ipm")
    return x, y, code, info


# ---------------------------------------------------------------------------------------------
def admm_solve(H, g, A, lo, hi, rho0=0.1, sigma=1e-6, alpha=1.6, max_iter=4000, check=25, first_check=25,
               adaptive_rho=True, eps_abs=1e-5, eps_rel=1e-5, fixed_iter=False, x0=None, y0=None):
    """OSQP iteration in the dense condensed form (device: admm_solve) without Ruiz
    scaling: x~ = K^-1 (sigma x - g + A'(rho z - y)), K = H + sigma I + A' diag(rho) A.
    ``fixed_iter``: run exactly max_iter iterations with no checks (and no rho updates)."""
    n, m = H.shape[0], A.shape[0]
    fixed = (hi[:n] - lo[:n]) < 1e-12

    def rho_vec(r):
        v = np.full(m, r)
        v[(lo < -INF_T) & (hi > INF_T)] = RHO_MIN
        v[(hi - lo) < 1e-4] = min(1e3 * r, RHO_MAX)
        return v

    rho = rho0
    rv = rho_vec(rho)
    x = np.zeros(n) if x0 is None else np.array(x0, float)
    y = np.zeros(m) if y0 is None else np.array(y0, float)
    x = np.clip(np.where(fixed, lo[:n], x), lo[:n], hi[:n])
    z = A @ x                      # OSQP: z = A x at a (warm or cold) start, no projection
    cf = _factor(H, A, rv, sigma, fixed)
    info = dict(status=ST_MAX_ITER, iters=0, nfac=1, rho=rho, pri=np.inf, dua=np.inf)
    next_check = max_iter if fixed_iter else min(first_check, max_iter)
    for it in range(1, max_iter + 1):
        wv = rv * z - y
        rhs = np.where(fixed, 0.0, sigma * x - g + A.T @ wv)
        xt = sla.cho_solve(cf, rhs)
        xt = np.where(fixed, lo[:n], xt)
        zt = A @ xt
        zr = alpha * zt + (1 - alpha) * z
        zn = np.clip(zr + y / rv, lo, hi)
        y = y + rv * (zr - zn)
        z = zn
        x = alpha * xt + (1 - alpha) * x
        info["iters"] = it
        if it != next_check and it != max_iter:
            continue
        next_check = it + check
        Hx, Ax, Aty = H @ x, A @ x, A.T @ y
        fr = ~fixed
        pri = np.abs(Ax - z).max()
        npri = max(np.abs(Ax).max(), np.abs(z).max())
        dua = np.abs((Hx + g + Aty)[fr]).max()
        ndua = max(np.abs(Hx[fr]).max(), np.abs(Aty[fr]).max(), np.abs(g[fr]).max())
        info.update(pri=pri, dua=dua)
        if not np.isfinite(pri + dua):
            info["status"] = ST_NON_FINITE
            break
        if fixed_iter:
            break
        if pri <= eps_abs + eps_rel * npri and dua <= eps_abs + eps_rel * ndua:
            info["status"] = ST_INEXACT
            break
        if it == max_iter:
            break
        if adaptive_rho:
            rn = rho * np.sqrt((pri / max(npri, 1e-10)) / max(dua / max(ndua, 1e-10), 1e-10))
            rn = min(max(rn, RHO_MIN), RHO_MAX)
            if rn > 5 * rho or rn < 0.2 * rho:
                rho = rn
                rv = rho_vec(rho)
                info["nfac"] += 1
                cf = _factor(H, A, rv, sigma, fixed)
    info["rho"] = rho
    code = np.where((z - lo) < -y, -1, np.where((hi - z) < y, 1, 0))
    return x, y, code, info
