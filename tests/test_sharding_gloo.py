"""N > 1 host logic on CPU: two gloo ranks each own a contiguous shard of the batch (scenario keyed by the
global hopper index), run the (emulated) hot path on it with no data-path collective, and gather the logged
results in global order -- the result equals the unsharded run bit for bit."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_shard(B, idx0, N, n_ticks):
    from hopper_mpc_inertial_b200 import scenarios
    from tests.emul import EmulMpc
    sc = scenarios.make_batch(B, idx0=idx0, N=N, n_ticks=n_ticks, seed=42)
    em = EmulMpc(B, N=N)
    em.set_gains(sc["Qdiag"], sc["Rdiag"])
    X = sc["X0"].copy()
    U_log = np.zeros((n_ticks, 6, B))
    for t in range(n_ticks):
        _, x_in = em.rk4(X, np.zeros((6, B)), np.zeros((3, B)), 0, convert=True)
        U, Xs, st, it, nf, pa = em.solve(x_in, sc["xref_tab"][t:t + N], sc["pf_tab"][t:t + N], sc["C_tab"][t], t == 0)
        U_log[t] = U[0]
        for i in range(20):
            pf = np.where(i < sc["pf_switch"][t][None, :], sc["pf_tab"][t], sc["pf_tab"][t + 1])
            X = em.rk4(X, U[0], pf, 1)
    return X, U_log


def _worker(rank, world, port, B_total, N, n_ticks, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from hopper_mpc_inertial_b200 import sharding
    r, w, _ = sharding.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = sharding.shard_range(B_total, rank, world)
    X, U_log = _run_shard(hi - lo, lo, N, n_ticks)
    Xg = sharding.gather_hoppers(torch.from_numpy(X), B_total)
    Ug = sharding.gather_hoppers(torch.from_numpy(U_log), B_total, dst=0)
    tmax = sharding.max_over_ranks(1.0 + rank, torch.device("cpu"))
    assert tmax == float(world)
    assert sharding.sum_over_ranks(hi - lo, torch.device("cpu")) == B_total
    if rank == 0:
        np.savez(out_path, X=Xg.numpy(), U=Ug.numpy())
    else:
        assert Ug is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_run_equals_unsharded(tmp_path):
    B_total, N, n_ticks = 5, 10, 3          # odd batch: shard sizes differ (3 + 2)
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), B_total, N, n_ticks, out), nprocs=2, join=True)
    got = np.load(out)
    X, U_log = _run_shard(B_total, 0, N, n_ticks)
    np.testing.assert_array_equal(got["X"], X)
    np.testing.assert_array_equal(got["U"], U_log)
