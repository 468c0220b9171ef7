"""Small closed-loop run for compute-sanitizer (racecheck / memcheck / synccheck) on the GPU box:
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
Covers the init tick (interior point + polish), warm ticks, the ADMM mode and the simulator kernel."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hopper_mpc_inertial_b200 import scenarios          # noqa: E402
from hopper_mpc_inertial_b200.batch import BatchMpc     # noqa: E402

T = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
for dyn, kw in (("3f", {}), ("2f", {}), ("3f", dict(solver="admm", max_iter=60, polish=1))):
    B, N, nt = 12, 10, 4
    sc = scenarios.make_batch(B, N=N, n_ticks=nt, dyn=dyn, seed=2)
    bm = BatchMpc(B, dyn=dyn, N=N, **kw)
    bm.set_gains(T(sc["Qdiag"]), T(sc["Rdiag"]))
    X = T(sc["X0"]).clone()
    out = bm.rollout(X, T(sc["xref_tab"]), T(sc["pf_tab"]), T(sc["C_tab"].view(np.int64)), T(sc["pf_switch"]), 0, nt, True)
    torch.cuda.synchronize()
    print(dyn, kw, "status", out["status"].cpu().numpy())
    bm.close()
