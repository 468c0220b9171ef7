"""Freeze the REFERENCE'S OWN closed loop into tests/golden/loop_ref_*.npz.  *** TEST INFRASTRUCTURE ONLY ***

Runs the unmodified ``Runner.run`` of /root/reference/src/robotrunner.py (robotrunner.py:81-124) through
oracle/refshim.py: the reference's planner, gait map, convert, mpcontrol (its own build_qp through the mini-cvxpy
shim), solve_qp and RK4, with the one thing that cannot be installed -- the OSQP binary behind cvxpy -- replaced by the
restated OSQP algorithm at cvxpy's settings (eps_abs = eps_rel = 1e-5, polish, cold start; oracle/qp_solvers.py).
The run's own plot calls (replaced by hooks, they would block on plt.show()) hand over X_traj and f_hist.

    python -m oracle.make_loop_ref          (build container only: needs /root/reference)
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def run_reference(dyn, curve, N_run):
    from oracle import refshim
    ref = refshim.load()
    cap = {}
    ref.plots.fplot = lambda N_run, p_hist, f_hist, s_hist: cap.update(f_hist=np.array(f_hist), s_hist=np.array(s_hist))
    ref.plots.posplot_animate_cube = lambda p_ref, X_hist: cap.update(X_ticks=np.array(X_hist))
    iters = []
    solve0 = ref.minicvx.Problem.solve

    def solve(self, *a, **k):
        r = solve0(self, *a, **k)
        iters.append(ref.minicvx.LAST["res"]["iters"])
        return r
    ref.minicvx.Problem.solve = solve
    try:
        runner = ref.robotrunner.Runner(dt=1e-3, dyn=dyn, curve=curve, N_run=N_run)
        t0 = time.time()
        runner.run()
        secs = time.time() - t0
    finally:
        ref.minicvx.Problem.solve = solve0
    x_ref, pf_ref = runner.path_plan_init(x_in=ref.robotrunner.convert(runner.X_0), xf=ref.robotrunner.convert(runner.X_f))
    return runner, cap, np.array(iters), x_ref, pf_ref, secs


def main():
    sys.path.insert(0, ROOT)
    from hopper_mpc_inertial_b200 import planner
    os.makedirs(OUT, exist_ok=True)
    for tag, dyn, curve in (("loop_ref_2f", "2f", False), ("loop_ref_3f_curve", "3f", True)):
        N_run, N = 240, 60
        runner, cap, iters, x_ref, pf_ref, secs = run_reference(dyn, curve, N_run)
        n_ticks = N_run // 20
        X_log = cap["X_ticks"][:n_ticks + 1]                   # X_traj[::20]: state at every tick boundary
        U_log = cap["f_hist"][0:N_run:20]                      # f_hist[20 t] = U[0] of tick t (zero-order hold)
        xt, pt, C, sw = planner.mpc_tables(x_ref, pf_ref, n_ticks, N, 20, 1e-3, 0.02, runner.t_start)
        np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), X_log=X_log, U_log=U_log, xref_tab=xt, pf_tab=pt, C=C,
                            pf_switch=sw, X0=runner.X_0, N=N, n_ticks=n_ticks, osqp_iters=iters,
                            osqp_settings=np.array([1e-5, 1e-5, 10000, 1]))   # eps_abs, eps_rel, max_iter, polish
        print(f"{tag}: {n_ticks} ticks of the reference's Runner.run in {secs:.1f} s, OSQP iterations per solve {iters.tolist()}")


if __name__ == "__main__":
    main()
