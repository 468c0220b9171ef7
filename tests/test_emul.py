"""The product's DEVICE source (csrc/*.cuh) compiled for the host and run serially (tests/emul) against
the oracle: the same arithmetic the CUDA kernels execute, checked without a GPU.  The -m gpu tests run
the real kernels through the C ABI against the same oracle."""
import numpy as np
import pytest

from hopper_mpc_inertial_b200 import scenarios
from oracle import device_port as dp
from oracle import hopper_oracle as ho
from oracle.closed_loop import OracleMpc, QPFailed
from tests.conftest import golden, normalised_oracle_qp, u_tol
from tests.emul import EmulMpc


def _x_in(sc):
    return np.stack([ho.convert(sc["X0"][:, b]) for b in range(sc["X0"].shape[1])], 1)


def test_rk4_and_convert_golden():
    g = golden("sim.npz")
    K = g["X"].shape[0]
    em = EmulMpc(K)
    X1 = em.rk4(g["X"].T, g["U"].T, g["pf"].T, 1)
    np.testing.assert_allclose(X1.T, g["Xn"], rtol=1e-13, atol=1e-13)
    X20, x20 = em.rk4(g["X"].T, g["U"].T, g["pf"].T, 20, convert=True)
    np.testing.assert_allclose(X20.T, g["X20"], rtol=1e-12, atol=1e-12)
    _, x0 = em.rk4(g["X"].T, g["U"].T, g["pf"].T, 0, convert=True)
    np.testing.assert_allclose(x0.T, g["x"], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("dyn", ["3f", "2f"])
def test_condense_matches_reference_qp_data(dyn):
    g = golden(f"qp_{dyn}.npz")
    N = 10
    prm = ho.Params(dyn=dyn, N=N)
    em = EmulMpc(1, dyn=dyn, N=N)
    for ci in range(3):
        x_guess = g[f"x_guess{ci}"]
        cb = np.array([sum(1 << k for k in range(N) if g[f"C{ci}"][k] != 0)], np.uint64)
        H, gg, lo, hi, inf = em.condense(g[f"x_in{ci}"][:, None], x_guess[:, :, None], g[f"x_ref{ci}"][:, :, None],
                                         g[f"pf{ci}"][:, :, None], cb)
        Ad, Bd, Gd = ho.gen_dt_dynamics(x_guess, g[f"pf{ci}"], prm)
        qc = ho.build_qp_condensed(g[f"x_in{ci}"], g[f"x_ref{ci}"], Ad, Bd, Gd, g[f"C{ci}"], prm)
        A, l, u = normalised_oracle_qp(qc, prm)
        np.testing.assert_allclose(H[..., 0], qc["H"], rtol=0, atol=1e-12 * np.abs(qc["H"]).max())
        np.testing.assert_allclose(gg[:, 0], qc["g"], rtol=0, atol=1e-12 * np.abs(qc["g"]).max())
        np.testing.assert_allclose(lo[:, 0], l, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(hi[:, 0], u, rtol=1e-12, atol=1e-12)
        assert inf[0] == 0


@pytest.mark.parametrize("dyn,N,precision,sweeps", [("3f", 10, 0, 1), ("2f", 10, 0, 1), ("3f", 20, 0, 1), ("3f", 10, 1, 1),
                                                    ("2f", 10, 1, 1), ("3f", 10, 0, 2), ("2f", 10, 0, 3)])
def test_closed_loop_matches_oracle(dyn, N, precision, sweeps):
    """precision 1 = FP32 factorisation / substitution with FP64 data and refinement: same parity bound.
    sweeps > 1 = hmpc_config.sqp_sweeps: relinearise about the tick's own solution and solve again."""
    B, n_ticks = (6, 25) if N == 10 else (2, 8)
    sc = scenarios.make_batch(B, N=N, n_ticks=n_ticks, seed=21, dyn=dyn)
    em = EmulMpc(B, dyn=dyn, N=N, precision=precision, sqp_sweeps=sweeps)
    em.set_gains(sc["Qdiag"], sc["Rdiag"])
    mpcs = [OracleMpc(ho.Params(dyn=dyn, N=N, Qdiag=sc["Qdiag"][:, b].copy(), Rdiag=sc["Rdiag"][:, b].copy()), sqp_sweeps=sweeps)
            for b in range(B)]
    X = sc["X0"].copy()
    paths = np.zeros(5, int)
    for t in range(n_ticks):
        x_in = np.stack([ho.convert(X[:, b]) for b in range(B)], 1)
        U, Xs, st, it, nf, pa = em.solve(x_in, sc["xref_tab"][t:t + N], sc["pf_tab"][t:t + N], sc["C_tab"][t], t == 0)
        if t == 0:
            assert np.all(it > 0)                 # cold start: the interior-point path ran
        for b in range(B):
            Uo = mpcs[b].mpcontrol(x_in[:, b], sc["xref_tab"][t:t + N, :, b], sc["pf_tab"][t:t + N, :, b], sc["C"][t, b], t == 0)
            assert st[b] == 0, (t, b, st[b], pa[b])
            assert np.all(np.abs(U[:, :, b] - Uo) <= u_tol(Uo)), (t, b, np.abs(U[:, :, b] - Uo).max())
            np.testing.assert_allclose(Xs[:, :, b], mpcs[b].xval, rtol=1e-6, atol=1e-7)
            paths[pa[b]] += 1
            for i in range(20):
                pf = sc["pf_tab"][t, :, b] if i < sc["pf_switch"][t, b] else sc["pf_tab"][t + 1, :, b]
                X[:, b] = ho.rk4_normalized(X[:, b], Uo[0], pf, mpcs[b].prm)
    assert paths[dp.ST_SOLVED + 1] > 0            # the warm path was exercised


@pytest.mark.parametrize("N", [3, 5])
def test_short_horizons_match_oracle(N):
    """Short horizons (the reference only ever runs N = 60): same parity bound; when the oracle finds the QP
    infeasible (the reference would raise) the device flags the hopper instead of returning a point."""
    B, n_ticks = 3, 10
    sc = scenarios.make_batch(B, N=N, n_ticks=n_ticks, seed=9)
    em = EmulMpc(B, N=N)
    em.set_gains(sc["Qdiag"], sc["Rdiag"])
    mpcs = [OracleMpc(ho.Params(N=N, Qdiag=sc["Qdiag"][:, b].copy(), Rdiag=sc["Rdiag"][:, b].copy())) for b in range(B)]
    X = sc["X0"].copy()
    alive = [True] * B
    checked = 0
    for t in range(n_ticks):
        x_in = np.stack([ho.convert(X[:, b]) for b in range(B)], 1)
        U, Xs, st, it, nf, pa = em.solve(x_in, sc["xref_tab"][t:t + N], sc["pf_tab"][t:t + N], sc["C_tab"][t], t == 0)
        for b in range(B):
            if not alive[b]:
                continue
            try:
                Uo = mpcs[b].mpcontrol(x_in[:, b], sc["xref_tab"][t:t + N, :, b], sc["pf_tab"][t:t + N, :, b], sc["C"][t, b], t == 0)
            except QPFailed:
                assert st[b] == dp.ST_INFEASIBLE
                alive[b] = False
                continue
            assert st[b] == 0
            assert np.all(np.abs(U[:, :, b] - Uo) <= u_tol(Uo))
            checked += 1
            for i in range(20):
                pf = sc["pf_tab"][t, :, b] if i < sc["pf_switch"][t, b] else sc["pf_tab"][t + 1, :, b]
                X[:, b] = ho.rk4_normalized(X[:, b], Uo[0], pf, mpcs[b].prm)
    assert checked >= 20


def test_infeasible_hopper_is_flagged_and_gets_zero_input():
    N = 10
    sc = scenarios.make_batch(2, N=N, n_ticks=2, seed=3)
    em = EmulMpc(2, N=N)
    x_in = _x_in(sc)
    x_in[2, 1] = 0.05          # below z_min at k = 0
    U, Xs, st, it, nf, pa = em.solve(x_in, sc["xref_tab"][:N], sc["pf_tab"][:N], sc["C_tab"][0], True)
    assert st[0] == 0 and st[1] == 2
    assert np.all(U[:, :, 1] == 0.0)
    with pytest.raises(QPFailed):
        OracleMpc(ho.Params(N=N)).mpcontrol(x_in[:, 1], sc["xref_tab"][:N, :, 1], sc["pf_tab"][:N, :, 1], sc["C"][0, 1], True)


@pytest.mark.parametrize("dyn", ["3f", "2f"])
def test_admm_mode_matches_numpy_port(dyn):
    """OSQP-style ADMM: fixed-iteration iterate equals the numpy statement to rounding; early exit stops
    at the same iteration with OSQP's residual test met; with polish the exact optimum is returned."""
    N, B = 10, 4
    sc = scenarios.make_batch(B, N=N, n_ticks=2, seed=5, dyn=dyn)
    x_in = _x_in(sc)
    xref, pfw = sc["xref_tab"][:N], sc["pf_tab"][:N]
    x_guess = np.concatenate((x_in[None], xref), 0)
    qps = []
    for b in range(B):
        p = ho.Params(dyn=dyn, N=N, Qdiag=sc["Qdiag"][:, b].copy(), Rdiag=sc["Rdiag"][:, b].copy())
        Ad, Bd, Gd = ho.gen_dt_dynamics(x_guess[:, :, b], pfw[:, :, b], p)
        qc = ho.build_qp_condensed(x_in[:, b], xref[:, :, b], Ad, Bd, Gd, sc["C"][0, b], p)
        qps.append((qc,) + normalised_oracle_qp(qc, p))
    for kw, fixed in ((dict(mode=1, max_iter=40, polish=0), True), (dict(mode=0, max_iter=4000, polish=0), False),
                      (dict(mode=0, max_iter=4000, polish=1), False)):
        em = EmulMpc(B, dyn=dyn, N=N, solver=1, warm_start=0, **kw)
        em.set_gains(sc["Qdiag"], sc["Rdiag"])
        em.Xsol[1:] = x_guess[:-1]              # time shift reproduces x_guess
        U, Xs, st, it, nf, pa = em.solve(x_in, xref, pfw, sc["C_tab"][0], False)
        for b in range(B):
            qc, A, lo, hi = qps[b]
            x, y, code, info = dp.admm_solve(qc["H"], qc["g"], A, lo, hi, max_iter=kw["max_iter"], fixed_iter=fixed)
            assert it[b] == info["iters"]
            if kw["polish"]:
                from oracle import qp_solvers as qs
                ref = qs.exact_qp(qc["H"], qc["g"], qc["A"], qc["l"], qc["u"])
                assert st[b] == 0
                np.testing.assert_allclose(U[:, :, b].reshape(-1), ref["x"], rtol=1e-6, atol=1e-6)
            else:
                assert st[b] == info["status"]
                np.testing.assert_allclose(U[:, :, b].reshape(-1), x, rtol=1e-7, atol=1e-7)


def test_contact_gate_sim_tick_matches_oracle():
    """hmpc_sim.cuh sim_tick with the contact gate (include/hmpc.h HMPC_GATE_*): the scheduled gate is the reference's
    commented-out `U[0, :] * s` (robotrunner.py:111) step by step; the detected gate tests the leg vector of
    dynamics_ct (robotrunner.py:143) against leg_max before every step."""
    from oracle.closed_loop import leg_reaches
    B, N = 5, 10
    sc = scenarios.make_batch(B, N=N, n_ticks=4, seed=3)
    em = EmulMpc(B, N=N)
    rng = np.random.default_rng(0)
    U = rng.normal(size=(6, B)) * 10
    U[2] += 70
    bits = np.array([0xfffff, 0x003ff, 0xffc00, 0x0, 0x5a5a5], np.uint32)
    sw = np.array([20, 5, 20, 12, 0], np.uint8)
    pfa, pfb = sc["pf_tab"][0], sc["pf_tab"][1]
    prm = ho.Params(N=N)
    X = sc["X0"].copy()
    X[2] = np.array([0.30, 0.36, 0.395, 0.41, 0.45])
    toggled = 0
    for mode, lm in ((0, 0.0), (1, 0.0), (2, 0.40)):
        Xn = em.sim_tick(X, U, pfa, pfb, sw, mode, bits, lm)
        for b in range(B):
            Xo, ons = X[:, b].copy(), []
            for i in range(20):
                pf = pfa[:, b] if i < sw[b] else pfb[:, b]
                s = 1.0 if mode == 0 else (float((int(bits[b]) >> i) & 1) if mode == 1 else float(leg_reaches(Xo, pf, prm, lm)))
                ons.append(s)
                Xo = ho.rk4_normalized(Xo, U[:, b] * s, pf, prm)
            np.testing.assert_allclose(Xn[:, b], Xo, rtol=1e-13, atol=1e-13)
            toggled += mode == 2 and 0 < sum(ons) < 20
    assert toggled >= 1            # the detected gate switched inside a tick for at least one hopper


def test_closed_loop_with_scheduled_gate_matches_oracle():
    """Closed loop with HMPC_GATE_SCHEDULE: emulated solver + gated simulator tick vs the oracle loop with the same gate."""
    from oracle.closed_loop import closed_loop
    B, N, n_ticks = 3, 10, 14
    sc = scenarios.make_batch(B, N=N, n_ticks=n_ticks, seed=12)
    em = EmulMpc(B, N=N)
    em.set_gains(sc["Qdiag"], sc["Rdiag"])
    X = sc["X0"].copy()
    Xl = [X.copy()]
    for t in range(n_ticks):
        _, x_in = em.rk4(X, np.zeros((6, B)), np.zeros((3, B)), 0, convert=True)
        U, Xs, st, it, nf, pa = em.solve(x_in, sc["xref_tab"][t:t + N], sc["pf_tab"][t:t + N], sc["C_tab"][t], t == 0)
        assert np.all(st == 0)
        X = em.sim_tick(X, U[0], sc["pf_tab"][t], sc["pf_tab"][t + 1], sc["pf_switch"][t], 1, sc["gate_tab"][t])
        Xl.append(X.copy())
    Xl = np.stack(Xl)
    gated = 0
    for b in range(B):
        prm = ho.Params(N=N, Qdiag=sc["Qdiag"][:, b].copy(), Rdiag=sc["Rdiag"][:, b].copy())
        gate = (sc["gate_tab"][:n_ticks, b, None].astype(np.int64) >> np.arange(20)) & 1
        gated += int((gate == 0).sum())
        Xo, Uo = closed_loop(prm, sc["X0"][:, b], sc["xref_tab"][:, :, b], sc["pf_tab"][:, :, b], sc["C"][:, b],
                             sc["pf_switch"][:, b], n_ticks, gate=gate)
        np.testing.assert_allclose(Xl[:, :, b], Xo, rtol=0, atol=2e-9)
    assert gated > 0


@pytest.mark.parametrize("polish", [1, 0])
def test_warp_admm_equals_cta_admm(polish):
    """solver = ADMM on warm ticks runs one warp per hopper (hmpc_warp.cuh: wadmm, tensor-core tiled factor and
    substitutions, shuffle reductions for OSQP's residual test and the rho update).  Same iteration as the CTA statement
    (hmpc_qp.cuh: admm_solve): equal iteration counts, factorisations and status, iterates equal to rounding; with
    polish the verified optimum."""
    B, N, n_ticks = 3, 10, 4
    sc = scenarios.make_batch(B, N=N, n_ticks=n_ticks, seed=21)
    ems = [EmulMpc(B, N=N, solver=1, polish=polish, hot_path=hp, max_iter=4000) for hp in (0, 1)]
    for em in ems:
        em.set_gains(sc["Qdiag"], sc["Rdiag"])
    X = sc["X0"].copy()
    warp_ticks = 0
    for t in range(n_ticks):
        _, x_in = ems[0].rk4(X, np.zeros((6, B)), np.zeros((3, B)), 0, convert=True)
        outs = [em.solve(x_in, sc["xref_tab"][t:t + N], sc["pf_tab"][t:t + N], sc["C_tab"][t], t == 0) for em in ems]
        (U0, X0s, st0, it0, nf0, pa0), (U1, X1s, st1, it1, nf1, pa1) = outs
        warp_ticks += ems[0].warp_done
        assert ems[1].warp_done == 0
        assert np.array_equal(st0, st1) and np.array_equal(it0, it1) and np.array_equal(nf0, nf1)
        assert np.all(st0 == (0 if polish else 4)) and np.all(pa0 == 4)
        np.testing.assert_allclose(U0, U1, rtol=1e-7, atol=1e-7)
        np.testing.assert_allclose(X0s, X1s, rtol=1e-7, atol=1e-8)
        for em in ems:                              # keep both on the same closed loop: the CTA statement's control
            em.Usol[:] = ems[1].Usol; em.Xsol[:] = ems[1].Xsol; em.code[:] = ems[1].code
            em.warm_ok[:] = 0                       # ... which supersedes the warp kernel's own warm blocks
        X = ems[1].sim_tick(X, U1[0], sc["pf_tab"][t], sc["pf_tab"][t + 1], sc["pf_switch"][t])
    assert warp_ticks == B * (n_ticks - 1)          # every warm tick went through the warp kernel
