"""Steady-state probe (GPU box helper): python tools/soak_probe.py [batch] [ticks] [key=value ...]
Runs the planned closed loop for `ticks` ticks and prints, per block of 10 ticks, the solver time per tick, the
hoppers handed to the CTA kernel, interior-point iterations and factorisations per tick.  key=value pairs override
hmpc_config fields (e.g. polish_retries=16)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hopper_mpc_inertial_b200 import planner, scenarios   # noqa: E402
from hopper_mpc_inertial_b200.batch import BatchMpc       # noqa: E402

pos = [a for a in sys.argv[1:] if "=" not in a]
kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[1:] if "=" in a)}
B = int(pos[0]) if pos else 131072
K = int(pos[1]) if len(pos) > 1 else 100
N, blk = 10, 10
dev = torch.device("cuda:0")
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
sc = scenarios.make_batch(B, N=N, n_ticks=K + 2, tables=False)
bm = BatchMpc(B, dyn="3f", N=N, on_infeasible="respawn", **kw)
bm.set_gains(T(sc["Qdiag"]), T(sc["Rdiag"]))
p = sc["plan"]
bm.plan_set(T(p["x0"]), T(p["xf"]), T(p["curve"]), T(p["tick_offset"]), planner.global_tables(**p["global_args"]))
X = T(sc["X0"]).clone()
bm.set_timing(True)
tot = 0.0
for t0 in range(0, K, blk):
    out = bm.rollout_planned(X, t0, blk, t0 == 0)
    torch.cuda.synchronize()
    mpc_ms, sim_ms, nt = bm.kernel_times()
    nf, pa, ni = [a.cpu().numpy() for a in bm.solve_stats()]
    it = out["iters"].cpu().numpy()
    hp = bm.hot_path_info()
    tot += mpc_ms + sim_ms
    print(f"ticks {t0:4d}-{t0 + blk - 1:4d}: solver {mpc_ms / nt:6.2f} ms/tick, {B / ((mpc_ms + sim_ms) / nt) / 1e3:6.2f} M steps/s, deferred/tick {hp['deferred'] / blk:8.1f}, "
          f"ipm it/tick {it.mean() / blk:.4f}, fac/tick {nf.mean() / blk:.3f}, infeasible ticks {int(ni.sum())}, ipm-path hoppers (last tick) {(pa >= 2).sum()}", flush=True)
