"""T2/T3: the oracle's QP forms and solvers agree with each other and with the KKT certificate."""
import numpy as np
import pytest

from oracle import device_port as dp
from oracle import hopper_oracle as ho
from oracle import qp_solvers as qs
from tests.conftest import golden, normalised_oracle_qp


def _case(dyn, ci, N=10, **kw):
    g = golden(f"qp_{dyn}.npz")
    prm = ho.Params(dyn=dyn, N=N, **kw)
    Ad, Bd, Gd = ho.gen_dt_dynamics(g[f"x_guess{ci}"], g[f"pf{ci}"], prm)
    return g, prm, Ad, Bd, Gd


@pytest.mark.parametrize("dyn", ["3f", "2f"])
@pytest.mark.parametrize("ci", [0, 1, 2])
def test_condensed_and_full_form_share_the_optimum(dyn, ci):
    g, prm, Ad, Bd, Gd = _case(dyn, ci)
    x_in, x_ref, C = g[f"x_in{ci}"], g[f"x_ref{ci}"], g[f"C{ci}"]
    qc = ho.build_qp_condensed(x_in, x_ref, Ad, Bd, Gd, C, prm)
    assert not qc["infeasible"]
    r = qs.exact_qp(qc["H"], qc["g"], qc["A"], qc["l"], qc["u"])
    assert r["ok"], r["cert"]
    U = r["x"].reshape(prm.N, 6)
    X = ho.rollout_linear(x_in, U, Ad, Bd, Gd, prm)
    qf = ho.build_qp_full(x_in, x_ref, Ad, Bd, Gd, C, prm)
    v = np.concatenate((X.reshape(-1), U.reshape(-1)))
    Av = qf["A"] @ v
    assert np.all(Av >= qf["l"] - 1e-8) and np.all(Av <= qf["u"] + 1e-8)          # feasible in the cvxpy-shaped form
    f_full = 0.5 * v @ qf["P"] @ v + qf["q"] @ v
    f_cond = 0.5 * r["x"] @ qc["H"] @ r["x"] + qc["g"] @ r["x"]
    # restated OSQP on the full form, tight tolerances: same optimum
    res = qs.osqp_solve(qf["P"], qf["q"], qf["A"], qf["l"], qf["u"], eps_abs=1e-9, eps_rel=1e-9, max_iter=100000,
                        adaptive_rho_interval=100)
    assert res["status"] == "solved"
    Uo = res["x"][qf["nx"]:].reshape(prm.N, 6)
    # an eps-terminated ADMM iterate is not the optimum: weakly determined input
    # directions (curvature 2R = 2e-3) are still ~1e-4 off at eps = 1e-9, hence the looser bound here
    np.testing.assert_allclose(Uo, U, rtol=1e-3, atol=2e-3)
    f_osqp = 0.5 * res["x"] @ qf["P"] @ res["x"] + qf["q"] @ res["x"]
    assert abs(f_osqp - f_full) <= 1e-6 * max(1.0, abs(f_full))
    # the two forms' objectives differ by a constant only: compare differences between two points
    U2 = np.clip(U * 0.9, None, None)
    X2 = ho.rollout_linear(x_in, U2, Ad, Bd, Gd, prm)
    v2 = np.concatenate((X2.reshape(-1), U2.reshape(-1)))
    d_full = (0.5 * v2 @ qf["P"] @ v2 + qf["q"] @ v2) - f_full
    d_cond = (0.5 * U2.reshape(-1) @ qc["H"] @ U2.reshape(-1) + qc["g"] @ U2.reshape(-1)) - f_cond
    assert abs(d_full - d_cond) <= 1e-8 * max(1.0, abs(d_full))


def test_uref_aliasing_changes_the_solution():
    """the reference's in-place u_ref makes every stage see the last-written value."""
    g, prm_a, Ad, Bd, Gd = _case("3f", 1)
    x_in, x_ref, C = g["x_in1"], g["x_ref1"], np.array([1, 1, 1, 1, 1, 0, 0, 0, 0, 0.0])
    prm_p = ho.Params(dyn="3f", N=10, uref_mode="per_stage")
    assert np.all(ho.uref_z(C, prm_a) == 0.0)                       # C[N-1] == 0 -> all stages see 0
    assert np.allclose(ho.uref_z(C, prm_p)[:5], 2 * 7.5 * 9.807)
    qa = ho.build_qp_condensed(x_in, x_ref, Ad, Bd, Gd, C, prm_a)
    qp = ho.build_qp_condensed(x_in, x_ref, Ad, Bd, Gd, C, prm_p)
    ra = qs.exact_qp(qa["H"], qa["g"], qa["A"], qa["l"], qa["u"])
    rp = qs.exact_qp(qp["H"], qp["g"], qp["A"], qp["l"], qp["u"])
    assert ra["ok"] and rp["ok"]
    assert np.abs(ra["x"] - rp["x"]).max() > 0.1


def test_height_infeasibility_is_flagged_exactly():
    g, prm, Ad, Bd, Gd = _case("3f", 0)
    x_in = g["x_in0"].copy()
    C = np.zeros(10)                      # all swing: ballistic, height rows are u-independent
    x_in[2] = 0.30; x_in[8] = -1.5        # falling fast -> z_k < 0.1 inside the horizon
    q = ho.build_qp_condensed(x_in, g["x_ref0"], Ad, Bd, Gd, C, prm)
    assert q["infeasible"] and not q["infeasible_const"]
    x_in[2] = 0.05                        # below z_min at k = 0 (App. D2)
    q = ho.build_qp_condensed(x_in, g["x_ref0"], Ad, Bd, Gd, C, prm)
    assert q["infeasible"] and q["infeasible_const"]
    x_in[2] = 0.4; x_in[8] = 0.0
    q = ho.build_qp_condensed(x_in, g["x_ref0"], Ad, Bd, Gd, np.ones(10), prm)
    assert not q["infeasible"]


@pytest.mark.parametrize("dyn", ["3f", "2f"])
def test_device_port_matches_exact_optimum(dyn):
    """The numpy statement of the device algorithms (warm/cold, IPM, verified polish, ADMM+polish)."""
    for ci in range(3):
        g, prm, Ad, Bd, Gd = _case(dyn, ci)
        qc = ho.build_qp_condensed(g[f"x_in{ci}"], g[f"x_ref{ci}"], Ad, Bd, Gd, g[f"C{ci}"], prm)
        ref = qs.exact_qp(qc["H"], qc["g"], qc["A"], qc["l"], qc["u"])
        A, lo, hi = normalised_oracle_qp(qc, prm)
        x, y, code, info = dp.solve_exact(qc["H"], qc["g"], A, lo, hi)
        assert info["status"] == dp.ST_SOLVED and info["path"] == "ipm+polish"
        np.testing.assert_allclose(x, ref["x"], rtol=1e-7, atol=1e-7)
        # warm start from a perturbed copy of the solution's active set
        x2, y2, code2, info2 = dp.solve_exact(qc["H"], qc["g"], A, lo, hi, warm=(x + 0.1, code))
        assert info2["status"] == dp.ST_SOLVED and info2["path"] == "warm" and info2["nfac"] == 1
        np.testing.assert_allclose(x2, ref["x"], rtol=1e-7, atol=1e-7)
        # multipliers satisfy the certificate of the normalised problem
        cert = qs.kkt_certificate(qc["H"], qc["g"], A, lo, hi, x, y)
        assert cert["stat"] < 1e-8 * max(1.0, np.abs(qc["g"]).max()) and cert["prim"] < 1e-8 and cert["sign"] < 1e-8
        # ADMM reaches OSQP's residual test and its iterate is near (not at) the optimum
        xa, ya, ca, ia = dp.admm_solve(qc["H"], qc["g"], A, lo, hi, max_iter=4000)
        assert ia["status"] == dp.ST_INEXACT
        assert np.abs(xa - ref["x"]).max() < 5.0
