"""Oracle MPC driver and closed loop.  *** TEST INFRASTRUCTURE ONLY *** (see hopper_oracle.py header)

``OracleMpc.mpcontrol`` follows Mpc.mpcontrol (mpc_cvx_euler_3f.py:41-69): first call = two solves with
x_guess[1:] = x_ref, later calls = time shift of the previous x.value.  Each QP is solved to its exact
(unique) optimum on the condensed form and certified by ``kkt_certificate``; ``solver='osqp'`` runs the
restated OSQP algorithm on the reference's full form instead (the reference's own settings).
``closed_loop`` follows Runner.run (robotrunner.py:96-113) on MPC-rate tables.
"""
from __future__ import annotations

import numpy as np

from . import hopper_oracle as ho
from . import qp_solvers as qs


class QPFailed(Exception):
    pass


class OracleMpc:
    def __init__(self, prm: ho.Params, solver="exact", osqp_opts=None, sqp_sweeps=1):
        self.prm, self.solver = prm, solver
        self.sqp_sweeps = max(1, int(sqp_sweeps))   # 1 = the reference; k > 1: product extension (hmpc_config.sqp_sweeps)
        self.osqp_opts = dict(osqp_opts or {})
        self.xval = None
        self.uval = None
        self.last = None

    def _solve(self, x_in, x_ref, pf, C, x_guess):
        prm = self.prm
        Ad, Bd, Gd = ho.gen_dt_dynamics(x_guess, pf, prm)
        if self.solver == "exact":
            qp = ho.build_qp_condensed(x_in, x_ref, Ad, Bd, Gd, C, prm)
            if qp["infeasible"]:
                raise QPFailed("infeasible: height rows z_k >= z_min cannot be met (min slack %.4g)" % float(np.min(qp["zmax"]) - prm.z_min))
            res = qs.exact_qp(qp["H"], qp["g"], qp["A"], qp["l"], qp["u"])
            if not res["ok"]:
                raise QPFailed(f"exact_qp not certified: {res['cert']}")
            U = res["x"].reshape(prm.N, 6)
            self.last = dict(qp=qp, res=res)
        else:
            qp = ho.build_qp_full(x_in, x_ref, Ad, Bd, Gd, C, prm)
            res = qs.osqp_solve(qp["P"], qp["q"], qp["A"], qp["l"], qp["u"], **self.osqp_opts)
            if res["status"] != "solved":
                raise QPFailed(res["status"])
            U = res["x"][qp["nx"]:].reshape(prm.N, 6)
            self.last = dict(qp=qp, res=res)
        X = ho.rollout_linear(x_in, U, Ad, Bd, Gd, prm)
        self.xval, self.uval = X, U
        return U

    def mpcontrol(self, x_in, x_ref_in, pf, C, init):
        N = self.prm.N
        x_guess = np.zeros((N + 1, 12))
        if init:
            x_guess[0] = x_in
            x_guess[1:] = x_ref_in
            self._solve(x_in, x_ref_in, pf, C, x_guess)
            x_guess = self.xval.copy()
        else:
            x_guess[0] = x_in
            x_guess[1:-1] = self.xval[2:]
            x_guess[-1] = self.xval[-1]
            for _ in range(self.sqp_sweeps - 1):
                self._solve(x_in, x_ref_in, pf, C, x_guess)
                x_guess = self.xval.copy()
        return self._solve(x_in, x_ref_in, pf, C, x_guess)


def leg_reaches(X, pf, prm, leg_max):
    """HMPC_GATE_DETECT: the leg vector of dynamics_ct, r = rh + R(q)'(pf - p) (robotrunner.py:143), is no longer than
    leg_max."""
    Rm = ho.quat_rotm(X[3:7])
    r = prm.rh + Rm.T @ (np.asarray(pf, float) - X[0:3])
    return float(r @ r) <= leg_max * leg_max


def closed_loop(prm: ho.Params, X0, xref_tab, pf_tab, C, pf_switch, n_ticks, solver="exact", osqp_opts=None,
                u_perturb=None, sqp_sweeps=1, gate=None, leg_max=None):
    """X0 (13,), xref_tab (T+N,12), pf_tab (T+N+1,3), C (T,N), pf_switch (T,).  Returns X_log
    (n_ticks+1,13), U_log (n_ticks,6).  ``u_perturb(t)`` optionally adds a perturbation to U[0] (used by
    the sensitivity study that justifies the closed-loop tolerance).  ``gate`` (T, mpc_factor) 0/1: the
    scheduled contact s at every simulator step, applied as the reference's commented-out factor; ``leg_max``:
    contact detected from the leg reach instead (product extension, include/hmpc.h HMPC_GATE_*)."""
    mpc = OracleMpc(prm, solver, osqp_opts, sqp_sweeps)
    X = np.array(X0, float)
    N = prm.N
    X_log = np.zeros((n_ticks + 1, 13)); U_log = np.zeros((n_ticks, 6))
    X_log[0] = X
    for t in range(n_ticks):
        x_in = ho.convert(X)
        U = mpc.mpcontrol(x_in, xref_tab[t:t + N], pf_tab[t:t + N], C[t], init=(t == 0))
        u0 = U[0].copy()
        if u_perturb is not None:
            u0 = u0 + u_perturb(t)
        for i in range(prm.mpc_factor):
            pf = pf_tab[t] if i < pf_switch[t] else pf_tab[t + 1]
            s = 1.0
            if gate is not None:                    # `f_hist[k, :] = U[0, :] * s` (robotrunner.py:111, commented out there)
                s = float(gate[t, i] != 0)
            elif leg_max is not None:
                s = 1.0 if leg_reaches(X, pf, prm, leg_max) else 0.0
            X = ho.rk4_normalized(X, u0 * s, pf, prm)
        X_log[t + 1] = X
        U_log[t] = u0
    return X_log, U_log
