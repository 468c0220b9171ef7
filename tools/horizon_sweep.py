"""BASELINE configs[4]: horizon sweep N = 10/20/40 at 65536 hoppers (2f and 3f): closed-loop steps/s, QP
iterations, factorisations, us per solve and the FP64 roofline fraction of the solver kernel.
Run on the GPU box:  python tools/horizon_sweep.py [--batch 65536] > gpurun_out/horizon_sweep.json"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hopper_mpc_inertial_b200 import planner, scenarios          # noqa: E402
from hopper_mpc_inertial_b200.batch import BatchMpc     # noqa: E402


def T(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


def run(dyn, N, B, warm=3, ticks=8, precision="fp64"):
    sc = scenarios.make_batch(B, N=N, n_ticks=warm + ticks, dyn=dyn, tables=False)
    bm = BatchMpc(B, dyn=dyn, N=N, on_infeasible="respawn", precision=precision)
    dev = bm.device
    bm.set_gains(T(sc["Qdiag"], dev), T(sc["Rdiag"], dev))
    p = sc["plan"]                                       # tables from the device-side planner
    bm.plan_set(T(p["x0"], dev), T(p["xf"], dev), T(p["curve"], dev), T(p["tick_offset"], dev), planner.global_tables(**p["global_args"]))
    tabs = bm.plan_tables(0, warm + ticks)
    args = (tabs["xref_tab"], tabs["pf_tab"], tabs["C_tab"], tabs["pf_switch"])
    X = T(sc["X0"], dev).clone()
    bm.rollout(X, *args, 0, warm, True)
    torch.cuda.synchronize()
    bm.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = bm.rollout(X, *args, warm, ticks, False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    mpc_ms, sim_ms, nt = bm.kernel_times()
    st = out["status"].cpu().numpy()
    nf, pa, ni = [a.cpu().numpy() for a in bm.solve_stats()]
    flops = float(bm.solve_flops().sum().item())
    peak = bm.measure_fp64_peak()
    res = dict(dyn=dyn, N=N, batch=B, ticks=ticks, steps_per_s=B * ticks / (ms * 1e-3),
               us_per_solve_amortised=mpc_ms * 1e3 / ticks / B, mpc_kernel_ms_per_tick=mpc_ms / ticks,
               sim_kernel_ms_per_tick=sim_ms / ticks, ipm_iters_per_tick=float(out["iters"].float().mean().item()) / ticks,
               factorisations_per_tick=float(nf.mean()) / ticks, solved_exact_frac=float(np.mean(st == 0)),
               inexact_frac=float(np.mean(st == 4)), infeasible_ticks=int(ni.sum()),
               flops_per_solve=flops / ticks / B, fp64_tflops=flops / (mpc_ms * 1e-3) / 1e12,
               fp64_peak_tflops=peak, fp64_frac=flops / (mpc_ms * 1e-3) / 1e12 / peak)
    bm.close()
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--horizons", type=int, nargs="+", default=[10, 20, 40])
    ap.add_argument("--dyns", nargs="+", default=["3f", "2f"])
    a = ap.parse_args()
    rows = []
    for N in a.horizons:
        for dyn in a.dyns:
            r = run(dyn, N, a.batch)
            rows.append(r)
            print(json.dumps(r), flush=True)
