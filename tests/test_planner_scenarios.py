"""Host-side callers of the hot path: planner / gait tables and synthetic scenarios."""
import numpy as np
import pytest

from hopper_mpc_inertial_b200 import planner, scenarios, sharding
from hopper_mpc_inertial_b200.batch import cbits_from_C
from oracle import hopper_oracle as ho
from tests.conftest import golden


@pytest.mark.parametrize("curve", [False, True])
def test_planner_matches_reference_golden(curve):
    g = golden("planner.npz")
    tag = "curve" if curve else "straight"
    x0 = np.zeros(12); x0[2] = 0.27
    xf = x0.copy(); xf[0] = 0.4 * 400 * 1e-3
    x_ref, pf_ref = planner.path_plan_init(x0, xf, 400, 60, 20, 1e-3, curve, 0.5 * 0.8 * 0.5)
    np.testing.assert_allclose(x_ref, g[f"x_ref_{tag}"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(pf_ref, g[f"pf_ref_{tag}"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(planner.path_plan_grab(x_ref, 40, 60, 20), g[f"grab_{tag}"], rtol=0, atol=1e-12)


def test_gait_map_bit_exact():
    g = golden("gait.npz")
    assert np.array_equal(planner.gait_scheduler(g["ts"]), g["sched"])
    assert np.array_equal(planner.gait_map(60, 0.02, g["map_ts"]), g["maps"])


def test_mpc_tables_reproduce_full_rate_tables():
    x0 = np.zeros(12); x0[2] = 0.27
    xf = x0.copy(); xf[0] = 0.8
    N, n_ticks = 10, 100
    x_ref, pf_ref = planner.path_plan_init(x0, xf, 2000, N, 20, 1e-3, False, 0.2)
    xt, pt, C, sw = planner.mpc_tables(x_ref, pf_ref, n_ticks, N, 20, 1e-3, 0.02, 0.2)
    t = 0.2
    for j in range(n_ticks):
        np.testing.assert_array_equal(xt[j:j + N], planner.path_plan_grab(x_ref, 20 * j, N, 20))
        np.testing.assert_array_equal(pt[j:j + N], planner.path_plan_grab(pf_ref, 20 * j, N, 20))
        for i in range(20):
            t = t + 1e-3
            if i == 0:
                np.testing.assert_array_equal(C[j], ho.gait_map(N, 0.02, t, 0, ho.Params()))
            want = pf_ref[20 * j + i]
            got = pt[j] if i < sw[j] else pt[j + 1]
            np.testing.assert_array_equal(got, want)


def test_batch_tables_match_per_hopper_planner():
    """The vectorised planner used for synthetic batches follows the reference planner's formulas."""
    B, N, n_ticks = 5, 10, 30
    sc = scenarios.make_batch(B, N=N, n_ticks=n_ticks, seed=9, t_p=0.8, phase_ticks=40)
    N_run = (n_ticks + 40 + 20) * 20
    for b in range(B):
        x0 = np.zeros(12); xf = np.zeros(12)
        # recover the planner end points from the tables: row 0 of hopper's reference at offset
        off = int(sc["tick_offset"][b])
        curve = bool(sc["curve"][b])
        # rebuild end points exactly as make_batch does
        u = scenarios._uniforms(9, b, 1, 41)[0]
        # compare against the full-rate planner run with the same end points
        x0 = sc["_x0p"][b]; xf = sc["_xfp"][b]
        x_ref, pf_ref = planner.path_plan_init(x0, xf, N_run, N, 20, 1e-3, curve, 0.2)
        xt, pt, C, sw = planner.mpc_tables(x_ref, pf_ref, off + n_ticks, N, 20, 1e-3, 0.02, 0.2)
        np.testing.assert_allclose(sc["xref_tab"][:, :, b], xt[off:off + n_ticks + N], rtol=0, atol=1e-9)
        np.testing.assert_allclose(sc["pf_tab"][:, :, b], pt[off:off + n_ticks + N + 1], rtol=0, atol=1e-9)
        np.testing.assert_array_equal(sc["C"][:, b], C[off:off + n_ticks])
        np.testing.assert_array_equal(sc["C_tab"][:, b], cbits_from_C(C[off:off + n_ticks]))
        # switch steps agree wherever the footstep actually changes inside the tick
        for j in range(n_ticks):
            for i in range(20):
                got = sc["pf_tab"][j, :, b] if i < sc["pf_switch"][j, b] else sc["pf_tab"][j + 1, :, b]
                np.testing.assert_allclose(got, pf_ref[20 * (off + j) + i], rtol=0, atol=1e-9)


def test_scenarios_do_not_depend_on_sharding():
    full = scenarios.make_batch(10, N=10, n_ticks=6, seed=77)
    for world in (2, 3):
        parts = []
        for r in range(world):
            lo, hi = sharding.shard_range(10, r, world)
            parts.append(scenarios.make_batch(hi - lo, idx0=lo, N=10, n_ticks=6, seed=77))
        for key in ("X0", "Qdiag", "Rdiag", "xref_tab", "pf_tab", "C_tab", "pf_switch"):
            np.testing.assert_array_equal(np.concatenate([p[key] for p in parts], axis=-1), full[key])


def test_shard_ranges_partition_the_batch():
    for B in (1, 7, 4096, 1048576):
        for world in (1, 2, 3, 4, 8):
            rs = [sharding.shard_range(B, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == B
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("dyn,N,n_ticks", [("3f", 10, 40), ("2f", 10, 12), ("3f", 20, 25)])
def test_device_planner_source_is_bit_identical_to_the_numpy_planner(dyn, N, n_ticks):
    """csrc/hmpc_plan.cuh compiled for the host (tests/emul) against planner.batch_tables: every generated row,
    contact mask and switch step equal bit for bit; the GPU run of the same source is checked in
    tests/test_gpu.py."""
    from tests import emul
    sc = scenarios.make_batch(96, N=N, n_ticks=n_ticks, seed=5, dyn=dyn)
    p = sc["plan"]
    gt = planner.global_tables(**p["global_args"])
    out = emul.plan_tables(p["x0"], p["xf"], p["curve"], p["tick_offset"], gt, N, n_ticks)
    assert np.array_equal(out["xref_tab"], sc["xref_tab"])
    assert np.array_equal(out["pf_tab"], sc["pf_tab"])
    assert np.array_equal(out["C_tab"], sc["C_tab"])
    assert np.array_equal(out["pf_switch"], sc["pf_switch"])
    # a window starting later in the run (what hmpc_rollout_planned generates per tick)
    out5 = emul.plan_tables(p["x0"], p["xf"], p["curve"], p["tick_offset"], gt, N, 1, tick0=5)
    assert np.array_equal(out5["xref_tab"], sc["xref_tab"][5:5 + N + 1])
    assert np.array_equal(out5["pf_tab"], sc["pf_tab"][5:5 + N + 2])
    assert np.array_equal(out5["C_tab"][0], sc["C_tab"][5]) and np.array_equal(out5["pf_switch"][0], sc["pf_switch"][5])
    # tables=False builds the same initial states without the tables
    sc2 = scenarios.make_batch(96, N=N, n_ticks=n_ticks, seed=5, dyn=dyn, tables=False)
    assert np.array_equal(sc2["X0"], sc["X0"]) and "pf_tab" not in sc2


def test_gate_masks_follow_the_run_clock():
    """planner.gate_masks: bit i of tick j = gait_scheduler at the reference's run-loop time of simulator step
    20 j + i (robotrunner.py:97-99: t = t + dt, s = gait_scheduler(t, t0)); per-hopper tables index the common clock."""
    n_ticks, mf, dt, t_start = 90, 20, 1e-3, 0.2
    gm = planner.gate_masks(n_ticks, mf, dt, t_start)
    assert gm.dtype == np.uint32
    t = t_start
    prm = ho.Params()
    for j in range(n_ticks):
        for i in range(mf):
            t = t + dt
            assert ((int(gm[j]) >> i) & 1) == int(ho.gait_scheduler(t, 0, prm)), (j, i)
    assert 0 < np.count_nonzero(gm == 0) and np.count_nonzero(gm == (1 << mf) - 1) > 0      # whole swing / stance ticks
    assert np.any((gm != 0) & (gm != (1 << mf) - 1))                                        # ticks with a contact switch inside
    sc = scenarios.make_batch(6, N=10, n_ticks=12, seed=4)
    gt = planner.global_tables(**sc["plan"]["global_args"])
    for b in range(6):
        off = int(sc["tick_offset"][b])
        assert np.array_equal(sc["gate_tab"][:, b], gt["gate_glob"][off:off + 12])
