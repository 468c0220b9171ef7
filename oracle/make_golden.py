"""Freeze reference outputs into tests/golden/.  *** TEST INFRASTRUCTURE ONLY ***

Run in the build container (needs /root/reference):   python -m oracle.make_golden

Every fixture below is produced by the REFERENCE's own code (bbokser/hopper-mpc-inertial, imported
unchanged through oracle/refshim.py with stand-ins for cvxpy / transforms3d / matplotlib) on seeded
inputs; the tests then check the oracle restatement (tests -m "not gpu") and the CUDA path
(tests -m gpu, where /root/reference does not exist) against these files.

  sim.npz        dynamics_ct, rk4_normalized, convert          robotrunner.py:19-28,126-164
  gait.npz       gait_scheduler / gait_map                     robotrunner.py:166-180
  planner.npz    path_plan_init (straight and --curve)         robotrunner.py:182-226
  lin_{2f,3f}.npz   Mpc.gen_dt_dynamics                        mpc_cvx_euler_*.py:70-94
  qp_{2f,3f}.npz    Mpc.build_qp executed through the mini-cvxpy shim -> (P, q, A, l, u) in a
                    canonical one-sided row form                mpc_cvx_euler_*.py:96-153
Two more fixtures are ORACLE outputs (exact optimum per tick), not reference outputs -- the reference's
cvxpy/OSQP back end cannot run here.  They save the GPU box from minutes of numpy closed loop:
  loop_2f.npz / loop_3f_curve.npz   closed loop of run.py 2f / 3f --curve over 2000 ms, N = 60
  loop_3f_curve_5s.npz              run.py 3f --curve with the default 5 s run time
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def canonical_rows(A, l, u, tol=1e-12):
    """l <= A v <= u  ->  sorted equality rows (a, b) and one-sided rows a v <= b, each scaled so that its
    largest |coefficient| is 1 (sign kept for inequalities, first non-zero made positive for equalities)."""
    eq, ineq = [], []
    for a, lo, hi in zip(A, l, u):
        s = np.abs(a).max()
        if s == 0:
            continue
        if hi - lo < tol:
            j = np.flatnonzero(a)[0]
            sg = np.sign(a[j])
            eq.append(np.concatenate((a * sg / s, [lo * sg / s])))
        else:
            if hi < 1e20:
                ineq.append(np.concatenate((a / s, [hi / s])))
            if lo > -1e20:
                ineq.append(np.concatenate((-a / s, [-lo / s])))

    def srt(rows, nv):
        if not rows:
            return np.zeros((0, nv + 1))
        R = np.array(rows)
        key = np.round(R, 9)
        order = np.lexsort(key.T[::-1])
        return R[order]
    nv = A.shape[1]
    return srt(eq, nv), srt(ineq, nv)


def _ref_inputs(rng, N, curve_yaw=False):
    x_in = np.array([0.1, -0.05, 0.33, 0.02, -0.03, 0.1, 0.3, -0.1, 0.2, 0.05, -0.02, 0.01]) + rng.normal(size=12) * 0.01
    x_ref = np.zeros((N, 12))
    x_ref[:, 0] = x_in[0] + 0.008 * np.arange(N)
    x_ref[:, 1] = x_in[1] + 0.002 * np.arange(N)
    x_ref[:, 2] = 0.35 + 0.05 * np.sin(0.3 * np.arange(N))
    x_ref[:, 5] = 0.1 - (0.01 * np.arange(N) if curve_yaw else 0.0)
    x_ref[:, 6] = 0.4
    pf = np.zeros((N, 3))
    pf[:, 0] = x_in[0] + 0.03
    pf[:, 1] = x_in[1] - 0.01
    return x_in, x_ref, pf


def main():
    sys.path.insert(0, ROOT)
    from oracle import refshim, hopper_oracle as ho
    from oracle.closed_loop import closed_loop
    ref = refshim.load()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)

    # ---- simulator ----
    runner = ref.robotrunner.Runner(dt=1e-3, dyn="3f", curve=False, N_run=200)
    K = 24
    X = np.zeros((K, 13)); U = np.zeros((K, 6)); PF = np.zeros((K, 3))
    dX = np.zeros((K, 13)); Xn = np.zeros((K, 13)); xc = np.zeros((K, 12)); X20 = np.zeros((K, 13))
    for i in range(K):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        if i < 4:
            q = np.array([1.0, 0, 0, 0]) if i % 2 == 0 else q
        X[i] = np.concatenate((rng.normal(size=3) * 0.3 + [0, 0, 0.4], q, rng.normal(size=3) * 0.5, rng.normal(size=3)))
        U[i] = np.concatenate((rng.normal(size=3) * 30 + [0, 0, 70], rng.normal(size=3) * 3))
        PF[i] = X[i, :3] * [1, 1, 0] + rng.normal(size=3) * [0.05, 0.05, 0]
        dX[i] = runner.dynamics_ct(X[i], U[i], PF[i])
        Xn[i] = runner.rk4_normalized(X[i], U[i], PF[i])
        xc[i] = ref.robotrunner.convert(X[i])
        Xs = X[i].copy()
        for _ in range(20):
            Xs = runner.rk4_normalized(Xs, U[i], PF[i])
        X20[i] = Xs
    np.savez_compressed(os.path.join(OUT, "sim.npz"), X=X, U=U, pf=PF, dX=dX, Xn=Xn, x=xc, X20=X20)

    # ---- gait ----
    ts = np.concatenate((np.linspace(0.0, 3.0, 301), 0.2 + 1e-3 * np.arange(1, 400)))
    sched = np.array([runner.gait_scheduler(t, 0) for t in ts])
    maps = np.array([runner.gait_map(60, 0.02, t, 0) for t in ts[::7]])
    np.savez_compressed(os.path.join(OUT, "gait.npz"), ts=ts, sched=sched, map_ts=ts[::7], maps=maps)

    # ---- planner ----
    plan = {}
    for curve in (False, True):
        r = ref.robotrunner.Runner(dt=1e-3, dyn="3f", curve=curve, N_run=400)
        x_ref, pf_ref = r.path_plan_init(x_in=ref.robotrunner.convert(r.X_0), xf=ref.robotrunner.convert(r.X_f))
        tag = "curve" if curve else "straight"
        plan[f"x_ref_{tag}"] = x_ref
        plan[f"pf_ref_{tag}"] = pf_ref
        plan[f"grab_{tag}"] = r.path_plan_grab(x_ref, 40)
    np.savez_compressed(os.path.join(OUT, "planner.npz"), **plan)

    # ---- linearisation + QP data ----
    N = 10
    Jinv = np.linalg.inv(ho.J_REF)
    for dyn, mod in (("3f", ref.mpc3f), ("2f", ref.mpc2f)):
        mpc = mod.Mpc(t=0.02, N=N, m=7.5, g=9.807, mu=1, Jinv=Jinv, rh=ho.RH_REF)
        lin, qp = {}, {}
        for ci, (Cpat, cy) in enumerate((([1] * 10, False), ([1, 1, 1, 0, 0, 0, 0, 1, 1, 1], True),
                                         ([0, 0, 0, 0, 1, 1, 1, 1, 1, 0], False))):
            x_in, x_ref, pf = _ref_inputs(rng, N, cy)
            C = np.array(Cpat, float)
            x_guess = np.vstack((x_in, x_ref))
            x_guess[1:, 0:3] += rng.normal(size=(N, 3)) * 0.01
            mpc.gen_dt_dynamics(x_guess, pf)
            Ad, Bd = np.array(mpc.Ad), np.array(mpc.Bd)
            lin[f"x_guess{ci}"] = x_guess; lin[f"pf{ci}"] = pf; lin[f"Ad{ci}"] = Ad; lin[f"Bd{ci}"] = Bd
            cost, constr = mpc.build_qp(x_in=x_in, x_ref=x_ref, Ad=mpc.Ad, Bd=mpc.Bd, Gd=mpc.Gd, C=C)
            prob = ref.minicvx.Problem(ref.minicvx.Minimize(cost), constr)
            from oracle.minicvx import _collect_variables
            can = prob.canonicalize(_collect_variables(prob))
            eq, ineq = canonical_rows(can["A"], can["l"], can["u"])
            qp.update({f"x_in{ci}": x_in, f"x_ref{ci}": x_ref, f"pf{ci}": pf, f"C{ci}": C, f"x_guess{ci}": x_guess,
                       f"Pdiag{ci}": np.diag(can["P"]).copy(), f"Poff{ci}": np.abs(can["P"] - np.diag(np.diag(can["P"]))).max(),
                       f"q{ci}": can["q"], f"const{ci}": can["const"], f"eq{ci}": eq, f"ineq{ci}": ineq})
        np.savez_compressed(os.path.join(OUT, f"lin_{dyn}.npz"), **lin)
        np.savez_compressed(os.path.join(OUT, f"qp_{dyn}.npz"), **qp)

    # ---- oracle closed loops of the two reference runs (N = 60, 2000 ms) ----
    sys.path.insert(0, ROOT)
    from hopper_mpc_inertial_b200 import planner
    for tag, dyn, curve, N_run in (("loop_2f", "2f", False, 2000), ("loop_3f_curve", "3f", True, 2000),
                                   ("loop_3f_curve_5s", "3f", True, 5000)):
        N60 = 60
        r = ref.robotrunner.Runner(dt=1e-3, dyn=dyn, curve=curve, N_run=N_run)
        x_ref, pf_ref = r.path_plan_init(x_in=ref.robotrunner.convert(r.X_0), xf=ref.robotrunner.convert(r.X_f))
        n_ticks = N_run // 20
        xt, pt, C, sw = planner.mpc_tables(x_ref, pf_ref, n_ticks, N60, 20, 1e-3, 0.02, r.t_start)
        prm = ho.Params(dyn=dyn, N=N60)
        Xl, Ul = closed_loop(prm, r.X_0, xt, pt, C, sw, n_ticks)
        np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), X_log=Xl, U_log=Ul, xref_tab=xt, pf_tab=pt, C=C,
                            pf_switch=sw, X0=r.X_0, N=N60, n_ticks=n_ticks)
        print(tag, "final", Xl[-1][:3])
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
