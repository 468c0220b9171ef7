"""T1/T2: the oracle restatement against the reference's OWN code (imported through
oracle/refshim.py, build container only) and against the golden fixtures frozen from it
(tests/golden, oracle/make_golden.py) -- the fixtures make the same checks run on the GPU box."""
import numpy as np
import pytest

from oracle import hopper_oracle as ho
from oracle.make_golden import canonical_rows
from tests.conftest import golden

PRM = ho.Params()


def test_sim_golden():
    g = golden("sim.npz")
    for i in range(g["X"].shape[0]):
        X, U, pf = g["X"][i], g["U"][i], g["pf"][i]
        np.testing.assert_allclose(ho.dynamics_ct(X, U, pf, PRM), g["dX"][i], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(ho.rk4_normalized(X, U, pf, PRM), g["Xn"][i], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(ho.convert(X), g["x"][i], rtol=1e-13, atol=1e-13)
        Xs = X.copy()
        for _ in range(20):
            Xs = ho.rk4_normalized(Xs, U, pf, PRM)
        np.testing.assert_allclose(Xs, g["X20"][i], rtol=1e-12, atol=1e-12)


def test_gait_golden_bit_exact():
    g = golden("gait.npz")
    sched = np.array([ho.gait_scheduler(t, 0, PRM) for t in g["ts"]])
    assert np.array_equal(sched, g["sched"])
    maps = np.array([ho.gait_map(60, 0.02, t, 0, PRM) for t in g["map_ts"]])
    assert np.array_equal(maps, g["maps"])


@pytest.mark.parametrize("curve", [False, True])
def test_planner_golden(curve):
    g = golden("planner.npz")
    tag = "curve" if curve else "straight"
    prm = ho.Params(N=60)
    X0 = np.array([0, 0, 0.27, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
    Xf = X0.copy(); Xf[0] = 0.4 * 400 * 1e-3
    x_ref, pf_ref = ho.path_plan_init(ho.convert(X0), ho.convert(Xf), prm, 400, curve, 0.5 * 0.8 * 0.5)
    np.testing.assert_allclose(x_ref, g[f"x_ref_{tag}"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(pf_ref, g[f"pf_ref_{tag}"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(ho.path_plan_grab(x_ref, 40, prm), g[f"grab_{tag}"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("dyn", ["3f", "2f"])
def test_linearisation_golden(dyn):
    g = golden(f"lin_{dyn}.npz")
    prm = ho.Params(dyn=dyn, N=10)
    for ci in range(3):
        Ad, Bd, Gd = ho.gen_dt_dynamics(g[f"x_guess{ci}"], g[f"pf{ci}"], prm)
        np.testing.assert_allclose(Ad, g[f"Ad{ci}"], rtol=0, atol=1e-15)
        np.testing.assert_allclose(Bd, g[f"Bd{ci}"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("dyn", ["3f", "2f"])
def test_qp_data_match_reference_build_qp(dyn):
    """The oracle's direct assembler reproduces the QP the reference's own build_qp emits through the
    mini-cvxpy shim: same P, q, constant and the same constraint set (row order / scaling canonicalised).
    This includes the u_ref aliasing of SURVEY App. D1."""
    g = golden(f"qp_{dyn}.npz")
    prm = ho.Params(dyn=dyn, N=10)
    for ci in range(3):
        Ad, Bd, Gd = ho.gen_dt_dynamics(g[f"x_guess{ci}"], g[f"pf{ci}"], prm)
        qp = ho.build_qp_full(g[f"x_in{ci}"], g[f"x_ref{ci}"], Ad, Bd, Gd, g[f"C{ci}"], prm)
        assert float(g[f"Poff{ci}"]) == 0.0
        np.testing.assert_allclose(np.diag(qp["P"]), g[f"Pdiag{ci}"], rtol=1e-14, atol=0)
        np.testing.assert_allclose(qp["q"], g[f"q{ci}"], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(qp["const"], float(g[f"const{ci}"]), rtol=1e-12)
        eq, ineq = canonical_rows(qp["A"], qp["l"], qp["u"])
        assert eq.shape == g[f"eq{ci}"].shape and ineq.shape == g[f"ineq{ci}"].shape
        np.testing.assert_allclose(eq, g[f"eq{ci}"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(ineq, g[f"ineq{ci}"], rtol=0, atol=1e-12)


# ---- live comparison against the reference modules (build container only) ----------------------
def test_live_reference_sim_and_gait(reference):
    rng = np.random.default_rng(5)
    runner = reference.robotrunner.Runner(dt=1e-3, dyn="2f", curve=False, N_run=100)
    for _ in range(10):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        X = np.concatenate((rng.normal(size=3), q, rng.normal(size=6)))
        U = rng.normal(size=6) * 20
        pf = rng.normal(size=3) * 0.2
        np.testing.assert_allclose(ho.rk4_normalized(X, U, pf, PRM), runner.rk4_normalized(X, U, pf), rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(ho.convert(X), reference.robotrunner.convert(X), rtol=1e-13, atol=1e-13)
        t = rng.uniform(0, 5)
        assert np.array_equal(ho.gait_map(60, 0.02, t, 0, PRM), runner.gait_map(60, 0.02, t, 0))


@pytest.mark.parametrize("dyn", ["3f", "2f"])
def test_live_reference_mpcontrol_through_shim(reference, dyn):
    """The reference's own mpcontrol (two solves on the first call), run through the mini-cvxpy shim with
    the restated OSQP at tight tolerances, lands on the oracle's exact optimum."""
    from oracle import minicvx
    from oracle.closed_loop import OracleMpc
    mod = reference.mpc3f if dyn == "3f" else reference.mpc2f
    N = 6
    prm = ho.Params(dyn=dyn, N=N)
    mpc = mod.Mpc(t=0.02, N=N, m=7.5, g=9.807, mu=1, Jinv=prm.Jinv, rh=prm.rh)
    x_in = np.array([0.0, 0.0, 0.3, 0.01, -0.02, 0.05, 0.2, 0.0, 0.1, 0.0, 0.0, 0.0])
    x_ref = np.tile(x_in, (N, 1)); x_ref[:, 0] += 0.01 * np.arange(N); x_ref[:, 2] = 0.32
    pf = np.zeros((N, 3)); pf[:, 0] = 0.02
    C = np.array([1, 1, 1, 0, 0, 1.0])
    old = dict(minicvx.SOLVER_OPTS)
    minicvx.SOLVER_OPTS.update(eps_abs=1e-10, eps_rel=1e-10, max_iter=200000, adaptive_rho_interval=100)
    try:
        U_ref = mpc.mpcontrol(x_in=x_in, x_ref_in=x_ref, pf=pf, C=C, init=True)
    finally:
        minicvx.SOLVER_OPTS.clear(); minicvx.SOLVER_OPTS.update(old)
    U_or = OracleMpc(prm).mpcontrol(x_in, x_ref, pf, C, True)
    np.testing.assert_allclose(U_ref, U_or, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("tag,dyn", [("loop_ref_2f", "2f"), ("loop_ref_3f_curve", "3f")])
def test_oracle_loop_matches_the_references_own_runner_golden(tag, dyn):
    """tests/golden/loop_ref_*.npz: 12 ticks of the reference's unmodified Runner.run at N = 60 (oracle/make_loop_ref.py,
    OSQP restated at eps 1e-5 + polish).  The oracle's exact-optimum loop -- what the GPU path is compared with --
    agrees with it to 1e-3 in the applied controls and 1e-5 in the state (measured 2.6e-4 / 2.3e-6)."""
    from oracle import hopper_oracle as ho
    from oracle.closed_loop import closed_loop
    from tests.conftest import golden
    g = golden(f"{tag}.npz")
    prm = ho.Params(dyn=dyn, N=int(g["N"]))
    n = int(g["n_ticks"])
    Xo, Uo = closed_loop(prm, g["X0"], g["xref_tab"], g["pf_tab"], g["C"], g["pf_switch"], n)
    assert np.abs(Uo - g["U_log"]).max() < 1e-3
    np.testing.assert_allclose(Xo, g["X_log"], rtol=0, atol=1e-5)
