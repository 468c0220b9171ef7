"""Does the order in which the persistent warps take the hoppers matter?  (GPU box helper)
python tools/order_probe.py [batch]: the same synthetic batch in its natural (random gait phase) order, sorted by gait
phase, and sorted by gait phase within blocks of 4096 -- solver time per tick over ticks 5..24."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hopper_mpc_inertial_b200 import planner, scenarios   # noqa: E402
from hopper_mpc_inertial_b200.batch import BatchMpc       # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
N, W, K = 10, 5, 20
dev = torch.device("cuda:0")
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
sc = scenarios.make_batch(B, N=N, n_ticks=W + K + 2, tables=False)
p = sc["plan"]
period = int(round(sc["t_p"] / 0.02)) if "t_p" in sc else 10
phase = np.asarray(p["tick_offset"]) % period
orders = {"natural": np.arange(B), "by phase": np.argsort(phase, kind="stable"),
          "by phase within 4096": np.concatenate([o + np.argsort(phase[o:o + 4096], kind="stable") for o in range(0, B, 4096)])}
for name, od in orders.items():
    bm = BatchMpc(B, dyn="3f", N=N, on_infeasible="respawn")
    bm.set_gains(T(sc["Qdiag"][:, od]), T(sc["Rdiag"][:, od]))
    bm.plan_set(T(p["x0"][:, od]), T(p["xf"][:, od]), T(np.asarray(p["curve"])[od]), T(np.asarray(p["tick_offset"])[od]), planner.global_tables(**p["global_args"]))
    X = T(sc["X0"][:, od]).clone()
    bm.rollout_planned(X, 0, W, True)
    bm.set_timing(True)
    out = bm.rollout_planned(X, W, K, False)
    torch.cuda.synchronize()
    mpc_ms, sim_ms, nt = bm.kernel_times()
    nf = bm.solve_stats()[0].double().mean().item() / K
    inv = np.argsort(od)
    chk = float(X[:, inv].double().sum().item())
    print(f"{name:22s}: solver {mpc_ms / nt:6.2f} ms/tick, {B / ((mpc_ms + sim_ms) / nt) / 1e3:6.2f} M steps/s, fac/tick {nf:.3f}, deferred {bm.hot_path_info()['deferred']}, state checksum {chk:.9e}", flush=True)
    bm.close()
