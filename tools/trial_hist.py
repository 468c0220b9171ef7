"""Histogram of factorisations per hopper-tick by solver path (GPU box helper): python tools/trial_hist.py [batch]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hopper_mpc_inertial_b200 import planner, scenarios   # noqa: E402
from hopper_mpc_inertial_b200.batch import BatchMpc       # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
N, W = 10, 12
dev = torch.device("cuda:0")
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
sc = scenarios.make_batch(B, N=N, n_ticks=W + 2, tables=False)
bm = BatchMpc(B, dyn="3f", N=N, on_infeasible="respawn")
bm.set_gains(T(sc["Qdiag"]), T(sc["Rdiag"]))
p = sc["plan"]
bm.plan_set(T(p["x0"]), T(p["xf"]), T(p["curve"]), T(p["tick_offset"]), planner.global_tables(**p["global_args"]))
X = T(sc["X0"]).clone()
bm.rollout_planned(X, 0, W, True)
out = bm.rollout_planned(X, W, 1, False)
torch.cuda.synchronize()
nf, pa, ni = [a.cpu().numpy() for a in bm.solve_stats()]
st = out["status"].cpu().numpy()
names = {0: "none", 1: "warm", 2: "ipm+polish", 3: "ipm", 4: "admm"}
print("deferred", bm.hot_path_info()["deferred"], "of", B)
for path in np.unique(pa):
    sel = pa == path
    h = np.bincount(nf[sel], minlength=12)
    print(f"path {names.get(int(path), path)}: {sel.sum()} hoppers ({100 * sel.mean():.2f} %), factorisations histogram {h.tolist()}")
print("status counts", np.bincount(st, minlength=5).tolist(), "mean factorisations", nf.mean())
