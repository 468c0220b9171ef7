"""Throughput of the stand-alone batch condense kernel (parity entry point hmpc_condense: linearise + condense
for every hopper, H / g / bounds written to HBM) -- a data point for the phase-split pipeline discussed in
DESIGN.md 5.4: how fast can one phase run when its state goes through HBM?   python tools/condense_throughput.py [B]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hopper_mpc_inertial_b200 import scenarios   # noqa: E402
from hopper_mpc_inertial_b200.batch import BatchMpc   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
N = 10
n, m = 6 * N, 11 * N
sc = scenarios.make_batch(B, N=N, n_ticks=2)
dev = torch.device("cuda:0")
T = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
bm = BatchMpc(B, dyn="3f", N=N, device=0)
bm.set_gains(T(sc["Qdiag"]), T(sc["Rdiag"]))
x_in = bm.convert(T(sc["X0"]))
xref, pf, cb = T(sc["xref_tab"][:N]), T(sc["pf_tab"][:N]), T(np.ascontiguousarray(sc["C_tab"][0]).view(np.int64))
x_guess = torch.cat((x_in[None], xref), 0).contiguous()
for _ in range(3):
    out = bm.condense(x_in, x_guess, xref, pf, cb)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
K = 10
ev[0].record()
for _ in range(K):
    out = bm.condense(x_in, x_guess, xref, pf, cb)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / K
wr = 8.0 * B * (n * n + n + 2 * m) + 4.0 * B            # H (full square), g, lo, hi, infeasible flag
rd = 8.0 * B * (12 + 12 * (N + 1) + 12 * N + 3 * N + 18) + 8.0 * B
print(json.dumps({"kernel": "condense_kernel", "batch": B, "ms": ms, "hoppers_per_s": B / ms * 1e3,
                  "bytes_written": wr, "bytes_read": rd, "GBps": (wr + rd) / ms / 1e6,
                  "note": "includes torch.empty of the 4 GB output per call"}))
