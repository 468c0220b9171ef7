"""First-light diagnostics on a GPU box: every stage of the hot path vs the CPU oracle, printing
errors instead of asserting.  Test infrastructure (imports oracle/)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hopper_mpc_inertial_b200 import scenarios
from hopper_mpc_inertial_b200.batch import BatchMpc, cbits_from_C
from oracle import hopper_oracle as ho, qp_solvers as qs
from oracle.closed_loop import OracleMpc, closed_loop


def T(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


def stage_checks(dyn, N, B=16, seed=0):
    print(f"=== {dyn} N={N} B={B}")
    sc = scenarios.make_batch(B, N=N, n_ticks=4, seed=seed, dyn=dyn)
    bm = BatchMpc(B, dyn=dyn, N=N)
    dev = bm.device
    bm.set_gains(T(sc["Qdiag"], dev), T(sc["Rdiag"], dev))
    X0 = sc["X0"]
    # convert
    xg = bm.convert(T(X0, dev)).cpu().numpy()
    xo = np.stack([ho.convert(X0[:, b]) for b in range(B)], 1)
    print("convert err", np.abs(xg - xo).max())
    # rk4
    rng = np.random.default_rng(1)
    U = rng.normal(size=(6, B)) * 20; pf = rng.normal(size=(3, B)) * 0.2
    Xd = T(X0, dev).clone()
    Xs = bm.rk4(Xd, T(U, dev), T(pf, dev), 20, log_steps=True).cpu().numpy()
    prm = ho.Params(dyn=dyn, N=N)
    err = 0
    for b in range(B):
        X = X0[:, b].copy()
        for k in range(20):
            X = ho.rk4_normalized(X, U[:, b], pf[:, b], prm)
            err = max(err, np.abs(X - Xs[k, :, b]).max())
    print("rk4 20 steps err", err)
    # linearize / condense
    x_in = xo
    xref = sc["xref_tab"][0:N]; pfw = sc["pf_tab"][0:N]
    x_guess = np.concatenate((x_in[None], xref), 0)
    Ad, Bd = bm.linearize(T(x_guess, dev), T(pfw, dev))
    Ad, Bd = Ad.cpu().numpy(), Bd.cpu().numpy()
    cb = sc["C_tab"][0]
    H, g, lo, hi, inf = [a.cpu().numpy() for a in bm.condense(T(x_in, dev), T(x_guess, dev), T(xref, dev), T(pfw, dev), T(cb.view(np.int64), dev))]
    eA = eB = eH = eg = el = 0
    for b in range(B):
        p = ho.Params(dyn=dyn, N=N, Qdiag=sc["Qdiag"][:, b].copy(), Rdiag=sc["Rdiag"][:, b].copy())
        Ao, Bo, Gd = ho.gen_dt_dynamics(x_guess[:, :, b], pfw[:, :, b], p)
        eA = max(eA, np.abs(Ao - Ad[..., b]).max()); eB = max(eB, np.abs(Bo - Bd[..., b]).max())
        qp = ho.build_qp_condensed(x_in[:, b], xref[:, :, b], Ao, Bo, Gd, sc["C"][0, b], p)
        eH = max(eH, np.abs(qp["H"] - H[..., b]).max() / np.abs(qp["H"]).max())
        eg = max(eg, np.abs(qp["g"] - g[..., b]).max() / np.abs(qp["g"]).max())
        lq = np.clip(qp["l"], -1e30, 1e30); n_ = 6 * N
        for k in range(2, N):
            if lq[n_ + 4 * N + k] > -1e26:
                lq[n_ + 4 * N + k] /= p.mpc_dt ** 2 * (k - 1) / p.m
        el = max(el, (np.abs(lq - lo[:, b]) / (1 + np.abs(lq))).max(), np.abs(np.clip(qp["u"], -1e30, 1e30) - hi[:, b]).max())
    print("linearize err Ad", eA, "Bd", eB, "| condense rel err H", eH, "g", eg, "bounds", el)
    # solve (init) vs oracle mpcontrol
    t0 = time.time()
    Ug, Xg, st, it = bm.solve(T(x_in, dev), T(xref, dev), T(pfw, dev), T(cb.view(np.int64), dev), True)
    torch.cuda.synchronize()
    print("solve time", time.time() - t0, "status", st.cpu().numpy(), "iters", it.cpu().numpy())
    Ug, Xg = Ug.cpu().numpy(), Xg.cpu().numpy()
    eu = ex = 0
    for b in range(B):
        p = ho.Params(dyn=dyn, N=N, Qdiag=sc["Qdiag"][:, b].copy(), Rdiag=sc["Rdiag"][:, b].copy())
        om = OracleMpc(p)
        try:
            Uo = om.mpcontrol(x_in[:, b], xref[:, :, b], pfw[:, :, b], sc["C"][0, b], True)
        except Exception as e:
            print("oracle failed", b, e); continue
        d = np.abs(Uo - Ug[..., b]); tol = 1e-5 + 1e-4 * np.abs(Uo)
        eu = max(eu, d.max()); ex = max(ex, np.abs(om.xval - Xg[..., b]).max())
        if (d > tol).any():
            print("  hopper", b, "U mismatch max", d.max(), "C", sc["C"][0, b].astype(int))
    print("solve: max |U-Uo|", eu, "max |X-Xo|", ex)
    return bm, sc


def loop_check(dyn, N, B=8, n_ticks=10):
    sc = scenarios.make_batch(B, N=N, n_ticks=n_ticks, seed=7, dyn=dyn)
    bm = BatchMpc(B, dyn=dyn, N=N)
    dev = bm.device
    bm.set_gains(T(sc["Qdiag"], dev), T(sc["Rdiag"], dev))
    X = T(sc["X0"], dev).clone()
    t0 = time.time()
    out = bm.rollout(X, T(sc["xref_tab"], dev), T(sc["pf_tab"], dev), T(sc["C_tab"].view(np.int64), dev), T(sc["pf_switch"], dev), 0, n_ticks, True, log=True)
    torch.cuda.synchronize()
    print(f"rollout {dyn} N={N} B={B} ticks={n_ticks}: {time.time()-t0:.3f}s status", out["status"].cpu().numpy(), "iters", out["iters"].cpu().numpy())
    Xg, Ug = out["X_log"].cpu().numpy(), out["U_log"].cpu().numpy()
    for b in range(B):
        p = ho.Params(dyn=dyn, N=N, Qdiag=sc["Qdiag"][:, b].copy(), Rdiag=sc["Rdiag"][:, b].copy())
        try:
            Xo, Uo = closed_loop(p, sc["X0"][:, b], sc["xref_tab"][:, :, b], sc["pf_tab"][:, :, b], sc["C"][:, b], sc["pf_switch"][:, b], n_ticks)
        except Exception as e:
            print("  oracle failed", b, e); continue
        print(f"  hopper {b}: max|dU| {np.abs(Uo-Ug[:,:,b]).max():.2e} max|dX| {np.abs(Xo-Xg[:,:,b]).max():.2e}")


def throughput(dyn, N, B, nt, **kw):
    sc = scenarios.make_batch(B, N=N, n_ticks=nt, dyn=dyn)
    bm = BatchMpc(B, dyn=dyn, N=N, **kw)
    dev = bm.device
    bm.set_gains(T(sc["Qdiag"], dev), T(sc["Rdiag"], dev))
    args = (T(sc["xref_tab"], dev), T(sc["pf_tab"], dev), T(sc["C_tab"].view(np.int64), dev), T(sc["pf_switch"], dev))
    X = T(sc["X0"], dev).clone()
    out = bm.rollout(X, *args, 0, 2, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = bm.rollout(X, *args, 2, nt - 2, False)
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    nf, pa, ni = [a.cpu().numpy() for a in bm.solve_stats()]
    print(f"{dyn} N={N} B={B} {kw} {nt-2} ticks: {dt:.3f}s -> {B*(nt-2)/dt:.0f} steps/s; status counts", np.bincount(st, minlength=5),
          "iters/tick", it.mean() / (nt - 2), "nfac/tick", nf.mean() / (nt - 2), "infeasible ticks", ni.sum(), "paths", np.bincount(pa, minlength=5))


if __name__ == "__main__":
    for dyn in ("3f", "2f"):
        stage_checks(dyn, 10)
    stage_checks("3f", 20, B=4)
    loop_check("3f", 10, n_ticks=30)
    loop_check("2f", 10, n_ticks=30)
    throughput("3f", 10, 4096, 42)
    throughput("3f", 10, 4096, 42, warm_start=0)
    throughput("3f", 10, 32768, 22)
    throughput("2f", 10, 4096, 42)
    throughput("3f", 10, 4096, 22, solver="admm", mode="fixed_iter", max_iter=50, polish=0)
    throughput("3f", 20, 1024, 12)
    bm = BatchMpc(1, dyn="3f", N=10)
    print("fp64 peak TFLOP/s", bm.measure_fp64_peak())
