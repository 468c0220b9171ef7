#!/usr/bin/env python
"""bench.py -- closed-loop hopper-MPC throughput on B200 (BASELINE.json metric) + reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # product arm (1 GPU by default)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # CPU arm: the reference's algorithm

One "step" is one closed-loop MPC tick of EVERY hopper of the batch: time shift + linearise + condense +
QP solve (mpcontrol, mpc_cvx_euler_3f.py:41-69) followed by mpc_factor = 20 RK4 simulator steps with the
zero-order-held control and the state conversion for the next tick (robotrunner.py:101-113).
Metric: hopper-MPC steps/s = hoppers x ticks / time, whole job over all ranks (weak scaling: the per-GPU
batch is fixed, hoppers are sharded by global index, no collective on the data path).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "hopper_mpc_steps_per_s"
UNIT = "closed-loop MPC steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--batch", type=int, default=131072, help="hoppers per GPU (1M over 8 GPUs = BASELINE config 4)")
    ap.add_argument("--horizon", type=int, default=10)
    ap.add_argument("--dyn", choices=["2f", "3f"], default="3f")
    ap.add_argument("--solver", choices=["exact", "admm"], default="exact")
    ap.add_argument("--precision", choices=["fp64", "fp32"], default="fp64",
                    help="fp32: FP32 factorisation / substitution with FP64 data and refinement (mixed precision)")
    ap.add_argument("--sqp-sweeps", type=int, default=1, help="relinearisation sweeps per tick (1 = the reference)")
    ap.add_argument("--hot-path", choices=["auto", "cta"], default="auto",
                    help="auto: warp-per-hopper kernel + CTA fallback (default); cta: round-1 CTA-per-hopper kernel only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the 4096-hopper and hard-workload records")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="wall-clock budget of the CPU baseline sample")
    return ap.parse_args()


def workload_name(args, world):
    return (f"{args.dyn} batch, {args.batch} hoppers/GPU x {world} GPU(s) = {args.batch * world} hoppers "
            f"(BASELINE configs[3]: 1M hoppers sharded over 8 B200 -> 131072 per GPU), horizon {args.horizon}, "
            f"randomised initial states, gains and references, FP64")


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                parts = [p.strip() for p in out.split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's per-tick recipe on the host cores, one process per core.
#   port-c++ : oracle/c/hopper_ref.cpp (g++ -O3): fresh cvxpy-shaped full QP every tick, restated OSQP (sparse banded
#              LDL', Ruiz scaling, eps 1e-5, adaptive rho, polish, cold start), 20 RK4 steps -- the compiled baseline
#   port     : the same recipe in numpy (oracle/closed_loop.py solver='osqp'), kept as a second figure
# Neither is the cvxpy/OSQP binary (not installable here: no network); both restate its published algorithm.
# ------------------------------------------------------------------------------------------------
def _cpu_worker_cpp(args):
    """Closed loops of hoppers idx, idx + stride, ... with the compiled restatement until the budget is used.
    Returns (ticks done, seconds, per-tick mpcontrol micro-seconds, OSQP iterations per tick, hoppers started, failed)."""
    idx, stride, dyn, N, seconds, max_ticks = args
    from hopper_mpc_inertial_b200 import scenarios
    from oracle import cref
    t0 = time.perf_counter()
    done, us, its, started, failed, inacc = 0, [], [], 0, 0, 0
    while time.perf_counter() - t0 < seconds:
        sc = scenarios.make_batch(1, idx0=idx, N=N, n_ticks=max_ticks, dyn=dyn)
        left = seconds - (time.perf_counter() - t0)
        r = cref.closed_loop(int(dyn[0]), N, sc["Qdiag"][:, 0], sc["Rdiag"][:, 0], sc["X0"][:, 0], sc["xref_tab"][:, :, 0],
                             sc["pf_tab"][:, :, 0], sc["C"][:, 0], sc["pf_switch"][:, 0], max_ticks, budget_s=max(left, 0.05))
        done += r["ticks"]; us += list(r["solve_us"]); its += list(r["iters"]); started += 1; failed += int(r["failed"])
        inacc += r["inaccurate"]
        idx += stride
    return done, time.perf_counter() - t0, us, its, started, failed, inacc


def _cpu_worker_py(args):
    """The same recipe in numpy (oracle/closed_loop.py solver='osqp').  Returns (ticks done, seconds)."""
    idx, dyn, N, seconds, max_ticks = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from hopper_mpc_inertial_b200 import scenarios
    from oracle import hopper_oracle as ho
    from oracle.closed_loop import OracleMpc, QPFailed
    sc = scenarios.make_batch(1, idx0=idx, N=N, n_ticks=max_ticks, dyn=dyn)
    prm = ho.Params(dyn=dyn, N=N, Qdiag=sc["Qdiag"][:, 0].copy(), Rdiag=sc["Rdiag"][:, 0].copy())
    mpc = OracleMpc(prm, solver="osqp")
    X = sc["X0"][:, 0].copy()
    t0 = time.perf_counter()
    done = 0
    for t in range(max_ticks):
        x_in = ho.convert(X)
        try:
            U = mpc.mpcontrol(x_in, sc["xref_tab"][t:t + N, :, 0], sc["pf_tab"][t:t + N, :, 0], sc["C"][t, 0], t == 0)
        except QPFailed:
            break
        for i in range(prm.mpc_factor):
            pf = sc["pf_tab"][t, :, 0] if i < sc["pf_switch"][t, 0] else sc["pf_tab"][t + 1, :, 0]
            X = ho.rk4_normalized(X, U[0], pf, prm)
        done += 1
        if time.perf_counter() - t0 > seconds:
            break
    return done, time.perf_counter() - t0


def cpu_reference_sample(dyn, N, seconds, cores=None, python_seconds=None):
    import multiprocessing as mp
    from oracle import cref
    cref.build()                                      # before the pool: one build, not one per process
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker_cpp, [(i, cores, dyn, N, seconds, 200) for i in range(cores)])
        py = pool.map(_cpu_worker_py, [(i, dyn, N, python_seconds, 400) for i in range(cores)]) if python_seconds else None
    wall = time.perf_counter() - t0
    ticks = sum(r[0] for r in res)
    value = sum(r[0] / r[1] for r in res if r[1] > 0)    # per-core rates add up: every process timed itself
    us = np.array([u for r in res for u in r[2]]) if ticks else np.zeros(1)
    its = np.array([i for r in res for i in r[3]]) if ticks else np.zeros(1)
    out = {"value": value, "unit": UNIT, "cores": cores, "kind": "port-c++", "wall_s": wall, "same_config": False,
           "p50_qp_solve_us": float(np.median(us)), "p99_qp_solve_us": float(np.percentile(us, 99)),
           "mean_osqp_iters_per_tick": float(its.mean()),
           "hoppers_started": int(sum(r[4] for r in res)), "hoppers_failed": int(sum(r[5] for r in res)),
           "solves_at_max_iter": int(sum(r[6] for r in res)),
           "sample": (f"{cores} processes (one per core), {ticks} closed-loop ticks in {seconds:.0f} s each; compiled C++ "
                      f"restatement (g++ -O3 -march=x86-64-v3) of the reference's per-tick recipe: cvxpy-shaped full QP rebuilt "
                      f"every tick, OSQP algorithm (banded LDL', Ruiz scaling, eps_abs=eps_rel=1e-5, adaptive rho, polish), "
                      f"cold start, + 20 RK4 steps; horizon {N}, {dyn}; same scenario generator as the GPU batch but only "
                      f"these hoppers (same_config=false).  NOT the cvxpy/OSQP binary (not installable: no network)")}
    if py:
        out["python_port"] = {"value": sum(r[0] / r[1] for r in py if r[1] > 0), "unit": UNIT, "kind": "port",
                              "sample": f"numpy restatement of the same recipe, {python_seconds:.0f} s per process"}
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cores = os.cpu_count() or 1
    per = []
    for s in range(args.warmup + args.steps):
        budget = max(1.5, min(args.cpu_seconds, 100.0 / max(1, args.warmup + args.steps)))
        r = cpu_reference_sample(args.dyn, args.horizon, budget, cores)
        if s >= args.warmup:
            per.append(r)
    value = float(np.mean([r["value"] for r in per]))
    last = per[-1]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean([r["wall_s"] for r in per])) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "note": "CPU arm: each step is a bounded sample of the same workload"},
            "cpu_baseline": {k: last[k] for k in ("unit", "cores", "kind", "sample", "same_config", "p50_qp_solve_us",
                                                   "mean_osqp_iters_per_tick")} | {"value": value},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
def plan_on_device(bm, sc, dev):
    """Hand the scenario's per-hopper planner scalars to the device-side planner (hmpc_plan_set)."""
    import torch
    from hopper_mpc_inertial_b200 import planner
    p = sc["plan"]
    gt = planner.global_tables(**p["global_args"])
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    bm.plan_set(T(p["x0"]), T(p["xf"]), T(p["curve"]), T(p["tick_offset"]), gt)


def quick_run(args, dev, local, B, W, K, tag, **scenario_kw):
    """A secondary configuration measured in the same process: W warm-up ticks (first = init), K timed ticks with the
    tables resident in HBM.  Returns a small record (steps/s, ms per tick, solver statistics)."""
    import torch
    from hopper_mpc_inertial_b200 import scenarios
    from hopper_mpc_inertial_b200.batch import BatchMpc
    N = args.horizon
    sc = scenarios.make_batch(B, N=N, n_ticks=W + K + 1, dyn=args.dyn, tables=False, **scenario_kw)
    bm = BatchMpc(B, dyn=args.dyn, N=N, device=local, solver=args.solver, precision=args.precision,
                  on_infeasible="respawn", sqp_sweeps=args.sqp_sweeps, hot_path=args.hot_path)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    bm.set_gains(T(sc["Qdiag"]), T(sc["Rdiag"]))
    plan_on_device(bm, sc, dev)
    tabs = bm.plan_tables(0, W + K + 1)               # the MPC-rate tables, generated on the device
    xr, pf, Cd, sw = tabs["xref_tab"], tabs["pf_tab"], tabs["C_tab"], tabs["pf_switch"]
    X = T(sc["X0"]).clone()
    bm.rollout(X, xr, pf, Cd, sw, 0, W, True)
    torch.cuda.synchronize()
    bm.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = bm.rollout(X, xr, pf, Cd, sw, W, K, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    mpc_t, _ = bm.tick_times()
    st = out["status"].cpu().numpy()
    nf, pa, ni = [a.cpu().numpy() for a in bm.solve_stats()]
    hot = bm.hot_path_info()
    rec = {"workload": tag, "batch": B, "steps": K, "warmup": W, "value": B * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K,
           "p50_qp_solve_us_amortised": float(np.median(mpc_t)) * 1e3 / B, "p50_qp_batch_latency_us": float(np.median(mpc_t)) * 1e3,
           "solved_exact_frac": float(np.mean(st == 0)), "infeasible_ticks": int(ni.sum()),
           "factorisations_per_tick": float(nf.mean() / K), "deferred_frac": hot["deferred"] / float(B * K)}
    bm.close()
    return rec


def run_b200(args):
    import torch
    from hopper_mpc_inertial_b200 import scenarios, sharding
    from hopper_mpc_inertial_b200.batch import BatchMpc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    rank, world, local = sharding.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, N, K, W = args.batch, args.horizon, args.steps, max(args.warmup, 3)
    n_ticks = W + K
    idx0 = rank * B                                   # contiguous global hopper range of this rank

    # ---- synthetic scenario for this shard: per-hopper scalars on the host (numpy, keyed by the global hopper index),
    # the MPC-rate tables generated by the device-side planner (bit-identical to planner.batch_tables), resident in HBM;
    # the e2e leg uploads the same tables from pinned host memory tick by tick ----
    sc = scenarios.make_batch(B, idx0=idx0, N=N, n_ticks=n_ticks + 1, dyn=args.dyn, tables=False)
    bm = BatchMpc(B, dyn=args.dyn, N=N, device=local, solver=args.solver, precision=args.precision,
                  on_infeasible="respawn", sqp_sweeps=args.sqp_sweeps, hot_path=args.hot_path)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    bm.set_gains(T(sc["Qdiag"]), T(sc["Rdiag"]))
    plan_on_device(bm, sc, dev)
    tabs = bm.plan_tables(0, n_ticks + 1)
    xref_d, pf_d, C_d, sw_d = tabs["xref_tab"], tabs["pf_tab"], tabs["C_tab"], tabs["pf_switch"]
    pin = lambda t: torch.empty(t.shape, dtype=t.dtype).pin_memory().copy_(t)
    xref_h, pf_h, C_h, sw_h = pin(xref_d), pin(pf_d), pin(C_d), pin(sw_d)
    X = T(sc["X0"]).clone()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (first tick is the init tick with two solves) ----
    bm.rollout(X, xref_d, pf_d, C_d, sw_d, 0, W, True)
    barrier()

    # ---- timed region 1: K ticks, inputs resident in HBM ----
    bm.set_timing(True)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = bm.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    out = bm.rollout(X, xref_d, pf_d, C_d, sw_d, W, K, False)
    e1.record()
    barrier()
    launches = bm.launch_count() - l0
    ms = e0.elapsed_time(e1)
    mpc_ms, sim_ms, nt = bm.kernel_times()
    mpc_tick, sim_tick = bm.tick_times()
    bm.set_timing(False)
    st = out["status"].cpu().numpy()
    it = out["iters"].cpu().numpy()
    nf, pa, ni = [a.cpu().numpy() for a in bm.solve_stats()]
    hot = bm.hot_path_info()
    flops = float(bm.solve_flops().sum().item())
    ms_max = sharding.max_over_ranks(ms, dev)

    # ---- timed region 2 (e2e): host buffers, per-step H2D of the step's inputs and D2H of its results ----
    # The reference-facing call per tick is mpcontrol(x_in, x_ref window, pf window, C) -> U followed by the
    # simulator; here every step uploads its reference rows from pinned host memory, runs one tick through
    # the same C ABI (hmpc_rollout over the 1-tick staging tables) and downloads the applied control, the
    # new state and the status.
    # Double-buffered: while tick t computes, a copy stream uploads tick t+1's rows and downloads tick t-1's
    # results, so the copies overlap the kernels; every copy still happens inside the timed region.
    cs = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    stage = [dict(x=torch.empty(N + 1, 12, B, dtype=torch.float64, device=dev),
                  p=torch.empty(N + 2, 3, B, dtype=torch.float64, device=dev),
                  c=torch.empty(1, B, dtype=torch.int64, device=dev),
                  s=torch.empty(1, B, dtype=torch.uint8, device=dev),
                  up=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
    res = [dict(u=torch.empty(6, B, dtype=torch.float64, device=dev), X=torch.empty(13, B, dtype=torch.float64, device=dev),
                st=torch.empty(B, dtype=torch.int32, device=dev), done=torch.cuda.Event(), read=torch.cuda.Event())
           for _ in range(2)]
    u_h = torch.empty(6, B, dtype=torch.float64).pin_memory()
    x_h = torch.empty(13, B, dtype=torch.float64).pin_memory()
    s_h = torch.empty(B, dtype=torch.int32).pin_memory()
    o2 = dict(status=bm.empty(B, dtype=torch.int32), iters=bm.empty(B, dtype=torch.int32),
              X_log=bm.empty(2, 13, B), U_log=bm.empty(1, 6, B))
    h2d = (stage[0]["x"][:N].numel() + stage[0]["p"][:N + 1].numel()) * 8 + stage[0]["c"].numel() * 8 + stage[0]["s"].numel()
    d2h = (u_h.numel() + x_h.numel()) * 8 + s_h.numel() * 4

    def upload(t, sb):
        with torch.cuda.stream(cs):
            cs.wait_event(sb["free"])                       # the tick that last used this buffer has finished
            sb["x"][:N].copy_(xref_h[t:t + N], non_blocking=True)
            sb["p"][:N + 1].copy_(pf_h[t:t + N + 1], non_blocking=True)
            sb["c"].copy_(C_h[t:t + 1], non_blocking=True)
            sb["s"].copy_(sw_h[t:t + 1], non_blocking=True)
            sb["up"].record(cs)

    def download(rb):
        with torch.cuda.stream(cs):
            cs.wait_event(rb["done"])
            u_h.copy_(rb["u"], non_blocking=True)
            x_h.copy_(rb["X"], non_blocking=True)
            s_h.copy_(rb["st"], non_blocking=True)
            rb["read"].record(cs)

    def e2e_run(t0, count):
        for sb in stage:
            sb["free"].record(main)
        for rb in res:
            rb["read"].record(main)
        upload(t0, stage[0])
        for i in range(count):
            sb, rb = stage[i % 2], res[i % 2]
            if i + 1 < count:
                upload(t0 + i + 1, stage[(i + 1) % 2])
            main.wait_event(sb["up"])
            bm.rollout(X, sb["x"], sb["p"], sb["c"], sb["s"], 0, 1, False, log=True, out=o2)
            sb["free"].record(main)
            main.wait_event(rb["read"])                     # the previous download from this result buffer is done
            rb["u"].copy_(o2["U_log"][0], non_blocking=True)
            rb["X"].copy_(X, non_blocking=True)
            rb["st"].copy_(o2["status"], non_blocking=True)
            rb["done"].record(main)
            download(rb)
        main.wait_stream(cs)

    # Same hoppers, same ticks as timed region 1 (the per-tick cost drifts with the tick index): a second
    # handle replays the warm-up from the initial states, then the e2e region covers ticks W .. W+K-1.
    bm_res, X_res = bm, X
    bm = BatchMpc(B, dyn=args.dyn, N=N, device=local, solver=args.solver, precision=args.precision, on_infeasible="respawn",
                  sqp_sweeps=args.sqp_sweeps, hot_path=args.hot_path)
    bm.set_gains(T(sc["Qdiag"]), T(sc["Rdiag"]))
    X = T(sc["X0"]).clone()
    bm.rollout(X, xref_d, pf_d, C_d, sw_d, 0, W - 2, True)
    e2e_run(W - 2, 2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_run(W, K)
    f1.record()
    barrier()
    e2e_state_matches = bool(torch.equal(X, X_res))      # the two paths computed the same closed loop
    e2e_ms = sharding.max_over_ranks(f0.elapsed_time(f1), dev)
    # ---- timed region 3 (e2e, device planner): nothing to upload -- every tick generates its own reference window on
    # the GPU from the per-hopper scalars set once (hmpc_rollout_planned); results are downloaded as above ----
    bm_e2e = bm
    bm = BatchMpc(B, dyn=args.dyn, N=N, device=local, solver=args.solver, precision=args.precision, on_infeasible="respawn",
                  sqp_sweeps=args.sqp_sweeps, hot_path=args.hot_path)
    bm.set_gains(T(sc["Qdiag"]), T(sc["Rdiag"]))
    plan_on_device(bm, sc, dev)
    Xp = T(sc["X0"]).clone()
    bm.rollout_planned(Xp, 0, W, True)

    def planned_run(t0, count):
        for rb in res:
            rb["read"].record(main)
        for i in range(count):
            rb = res[i % 2]
            bm.rollout_planned(Xp, t0 + i, 1, False, log=True, out=o2)
            main.wait_event(rb["read"])
            rb["u"].copy_(o2["U_log"][0], non_blocking=True)
            rb["X"].copy_(Xp, non_blocking=True)
            rb["st"].copy_(o2["status"], non_blocking=True)
            rb["done"].record(main)
            download(rb)
        main.wait_stream(cs)

    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    planned_run(W, K)
    g1.record()
    barrier()
    planned_state_matches = bool(torch.equal(Xp, X_res))
    planned_ms = sharding.max_over_ranks(g0.elapsed_time(g1), dev)
    bm.close()
    bm = bm_e2e
    clocks = sampler.stop()
    # NCCL is used for exactly one thing: gathering logged results after the run (here the final states)
    Xg = sharding.gather_hoppers(X, B * world, dst=0)
    gather_info = None
    if rank == 0:
        gather_info = {"backend": "nccl" if world > 1 else "none (1 rank)", "bytes": int(Xg.numel() * 8),
                       "mean_height_m": float(Xg[2].mean().item()), "finite": bool(torch.isfinite(Xg).all().item())}

    # ---- aggregate over ranks ----
    tot = lambda v: sharding.sum_over_ranks(v, dev)
    n_hop = tot(B)
    value = n_hop * K / (ms_max * 1e-3)
    e2e_value = n_hop * K / (e2e_ms * 1e-3)
    solved = tot(float(np.sum(st == 0))) / n_hop
    inf_ticks = tot(float(ni.sum()))
    launches_all = tot(launches)
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (mpc_kernel: K1 condense + K2 solve), this rank ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    m_rows = 11 * N
    # algorithmic HBM bytes per hopper-tick of mpc_kernel (DESIGN.md "Kernels"): x_in 12, gains 18, reference
    # window 15 N, contact mask 1, previous trajectory (p, yaw of N stages) 4 N read; trajectory 12 (N+1),
    # inputs 6 N + 6 written; warm start: inputs 6 N read, active-set codes 11 N bytes read + written; 4 int32 stats
    bytes_tick = 8 * (12 + 18 + 15 * N + 1 + 4 * N + 12 * (N + 1) + 6 * N + 6 + 6 * N) + 2 * m_rows + 6 * 4
    if hot["warps_per_sm"]:
        # phase split: the prep kernel writes and the solve kernel reads the per-hopper QP record (compact Hessian over
        # the nf = 6N - 3 N_swing non-fixed variables, N_swing ~ N/2, rows padded to 8; gradient; height bounds; x_in)
        nfree = 6 * N - 3 * (N // 2) if args.dyn == "3f" else 5 * N - 2 * (N // 2)
        bytes_tick += 2 * 8 * (nfree * ((nfree + 7) // 8 * 8) + 6 * N + N + 12)
    mpc_s = mpc_ms * 1e-3 / max(nt, 1)
    ach_gbs = bytes_tick * B / mpc_s / 1e9
    fp64_peak = bm.measure_fp64_peak()
    ach_tf = flops / max(nt, 1) / mpc_s / 1e12
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
    except Exception:
        pass
    # The solver kernels' arithmetic runs on the FP64 pipe (DMMA.8x8x4 tensor-core tiles + DFMA): that is the roofline
    # the kernel is measured against.  It is far from it: what binds is latency -- dependent DMMA / shuffle chains, L2 reads
    # of the Hessian record and the lock-step barriers at 12 resident warps per SM (ncu, profiles/README.md) -- not FP64
    # throughput and not HBM.
    kname = "mpc_prep_kernel + mpc_warp_rounds_kernel + mpc_kernel (deferred hoppers)" if hot["warps_per_sm"] else "mpc_kernel"
    roofline = {"bound": "tensor", "pipe": "FP64 (DMMA.8x8x4 tensor-core tiles and DFMA share the FP64 pipe)", "kernel": kname,
                "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak if fp64_peak else None,
                "traffic": (prof["mpc_kernel_dram_bytes_per_hopper"] * B) if "mpc_kernel_dram_bytes_per_hopper" in prof else None,
                "traffic_source": ("ncu dram__bytes_read+write per hopper at batch %d (profiles/%s) x this batch" % (prof.get("capture_batch", 0), prof.get("source", "?"))) if prof else None,
                "peak_source": "measured in this run: dependent-chain-free FP64 FMA microbenchmark (hmpc_measure_fp64_peak); "
                               "tools/micro/dmma.cu measures the same 37 TFLOP/s through DMMA",
                "algorithmic_flops_per_launch": flops / max(nt, 1), "algorithmic_flops_per_hopper_tick": flops / max(nt, 1) / B,
                "avg_launch_ms": mpc_s * 1e3, "share_of_step": mpc_ms / ms,
                "launches_per_step": "prep + solve + deferred-list kernel (timed together with CUDA events around the three launches)",
                "binding_resource": "dependent-issue latency at 12 resident warps per SM (shared memory: 18.9 KB of KKT factor and vectors per "
                                    "hopper): ncu of the same command (profiles/r2ar_mpc_kernel_ncu.txt) shows issue slots 31 % busy, FP64 pipe "
                                    "6.7 %, DRAM 6.1 %; stalls: fixed-latency wait 25 %, long scoreboard (L2) 23 %, lock-step barriers 18 %.  "
                                    "Neither FP64 nor HBM binds"}
    roofline_hbm = {"bound": "hbm", "kernel": kname, "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                    "peak_source": hbm_src, "algorithmic_bytes_per_launch": bytes_tick * B,
                    "note": "reported for completeness: arithmetic intensity >> machine balance, HBM does not bind"}
    # sim_kernel (K3): one thread per hopper, 20 RK4 steps in registers -- the one kernel whose bound is FP64 arithmetic.
    # Algorithmic FLOPs: SURVEY 8(d) F_rk4 = mpc_factor x 1300 per hopper-tick.
    sim_s = sim_ms * 1e-3 / max(nt, 1)
    sim_flops = 20 * 1300.0 * B
    roofline_sim = {"bound": "fp64", "kernel": "sim_kernel", "achieved": sim_flops / sim_s / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": (sim_flops / sim_s / 1e12 / fp64_peak) if fp64_peak else None, "avg_launch_ms": sim_s * 1e3,
                    "algorithmic_flops_per_hopper_tick": 20 * 1300.0, "share_of_step": sim_ms / ms,
                    "ncu": "profiles/r2ab_sim_kernel_ncu.txt (FP64 pipe active cycles and DFMA issue rate of the same kernel)"}
    working_set_mb = B * (bytes_tick + 13 * 16 + 15 * 8 * 2) / 1e6
    cache_note = ("inputs larger than L2 (per-tick working set %.0f MB per GPU > 126 MB L2)" % working_set_mb) if working_set_mb > 126 \
        else ("per-tick working set %.0f MB per GPU fits the 126 MB L2; every tick reads new reference rows and rewrites the "
              "whole per-hopper state, no flush between ticks" % working_set_mb)
    p50 = {"amortised_us_per_solve": float(np.median(mpc_tick)) * 1e3 / B, "batch_latency_us": float(np.median(mpc_tick)) * 1e3,
           "p99_batch_latency_us": float(np.percentile(mpc_tick, 99)) * 1e3,
           "definition": "median over the timed ticks of the solver kernels' device time (time shift + linearise + condense + QP "
                         "solve for the whole batch of one tick, CUDA events); amortised = that / hoppers per GPU"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == "fp64" else "f64 data and residuals, f32 factorisation (mixed precision)",
            "data": "synthetic",
            "config": {"workload": workload_name(args, world), "precision": args.precision, "dyn": args.dyn, "horizon": N, "batch_per_gpu": B,
                       "solver": args.solver, "sqp_sweeps": args.sqp_sweeps, "mpc_factor": 20, "parallelism": f"shard-by-hopper x{world}, no data-path collective",
                       "cache": cache_note, "hot_path": args.hot_path,
                       "on_infeasible": "respawn"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / K, "same_ticks_as_value": True, "final_state_equals_resident_run": e2e_state_matches,
                    "overlap": "double-buffered: uploads of tick t+1 / downloads of tick t-1 on a copy stream"},
            "e2e_device_planner": {"value": n_hop * K / (planned_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(d2h),
                                   "ms_per_step": planned_ms / K, "final_state_equals_resident_run": planned_state_matches,
                                   "note": "hmpc_rollout_planned: reference rows, footsteps, contact masks generated per tick on the GPU "
                                           "from per-hopper planner scalars uploaded once (SURVEY 8 row f1); the headline e2e above "
                                           "uploads host-built windows every tick"},
            "gpu_launches": int(launches_all),
            "clocks": clocks, "log_gather": gather_info,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_sim_kernel": roofline_sim, "p50_qp_solve_us": p50,
            "solver_stats": {"solved_exact_frac": solved, "infeasible_ticks": int(inf_ticks),
                             "ipm_iters_per_tick": float(it.mean() / K), "factorisations_per_tick": float(nf.mean() / K),
                             "warm_path_frac_last_tick": float(np.mean(pa == 1)),
                             "hot_path": dict(hot, mode=args.hot_path, deferred_frac=hot["deferred"] / float(B * K)),
                             "mpc_kernel_ms_per_tick": mpc_ms / max(nt, 1), "sim_kernel_ms_per_tick": sim_ms / max(nt, 1)}}
    if world == 1 and not args.no_extra_configs:
        # BASELINE.json configs[2] (4096 hoppers on one B200) and a harder synthetic workload (gains x logU(0.5, 2),
        # full-size initial perturbations: SURVEY 8(d)), measured in this same process
        line["configs"] = [quick_run(args, dev, local, 4096, W, K, "BASELINE configs[2]: 3f batch of 4096 hoppers, 1 B200" if args.dyn == "3f" else "4096 hoppers"),
                           quick_run(args, dev, local, 32768, W, K, "hard: 32768 hoppers, gain_spread 2.0 (logU(0.5,2)), perturb 1.0",
                                     gain_spread=2.0, perturb=1.0)]
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_sample(args.dyn, N, args.cpu_seconds, python_seconds=4.0)
        line["p50_qp_solve_us"]["cpu_us_per_solve"] = line["cpu_baseline"]["p50_qp_solve_us"]
    if world > 1:
        torch.distributed.destroy_process_group()
    _emit(line)


def _emit(line):
    """Write the JSON line to the REAL stdout (everything else that lands on fd 1 -- e.g. NCCL's version
    banner -- has been diverted to stderr by main())."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # stray prints of libraries go to stderr; stdout carries one JSON line
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
