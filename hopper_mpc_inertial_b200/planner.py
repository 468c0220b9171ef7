"""Host-side planning for the MPC hot path: reference trajectories, footsteps and contact schedules.

This is the caller side of the hot path: a one-off precompute per run, done in
numpy/scipy on the host exactly as the reference does, then shipped to the GPU as MPC-rate tables.

Reference behaviour (file:line into the reference's src/):
  gait_scheduler / gait_map   robotrunner.py:166-180  (accumulated float sums)
  path_plan_init              robotrunner.py:182-226  (incl. the --curve column quirk, App. D4)
  path_plan_grab              robotrunner.py:228-230
  run-loop time stepping      robotrunner.py:83-113   (t += dt before use; MPC at k = 0, 20, 40 ...)
"""
from __future__ import annotations

import numpy as np

T_P = 0.8           # gait period (robotrunner.py:43)
PHI_SWITCH = 0.5    # stance fraction (robotrunner.py:44)
STEP_ADJUSTMENT = -115  # robotrunner.py:79


def gait_scheduler(t, t0=0.0, t_p=T_P, phi_switch=PHI_SWITCH):
    """1 = scheduled stance, 0 = swing (robotrunner.py:166-171).  Works on scalars and arrays."""
    phi = np.mod((t - t0) / t_p, 1)
    return np.where(phi > phi_switch, 0, 1)


def gait_map(N, dt, ts, t0=0.0, t_p=T_P, phi_switch=PHI_SWITCH):
    """Contact flags over a horizon with the reference's running sum ts += dt (robotrunner.py:172-180).

    ``ts`` may be an array (one start time per hopper); returns shape (N,) or (B, N) float64."""
    ts = np.array(ts, dtype=float)
    out = np.zeros(ts.shape + (N,))
    for k in range(N):
        out[..., k] = gait_scheduler(ts, t0, t_p, phi_switch)
        ts = ts + dt
    return out


def path_plan_init(x_in, xf, N_run, N, mpc_factor, dt, curve, t_start, t_p=T_P, phi_switch=PHI_SWITCH,
                   step_adjustment=STEP_ADJUSTMENT):
    """Full-rate reference for one hopper: x_ref (N_run + N*mpc_factor, 12), pf_ref (.., 3)."""
    from scipy.interpolate import CubicSpline
    from scipy.signal import find_peaks
    N_k = N * mpc_factor
    t_ref = N_run + N_k
    x_ref = np.linspace(start=x_in, stop=xf, num=N_run)
    if curve:
        knots = np.array([0, N_run * 0.5, N_run])
        k_all = np.arange(N_run)
        x_ref[:, 0] = CubicSpline(knots, np.array([x_in[1], xf[1] * 0.9, xf[1]]))(k_all)
        s45 = np.sin(45 * np.pi / 180)
        x_ref[:, 5] = CubicSpline(knots, np.array([0, -s45 * 0.4, -s45]))(k_all)
        x_ref[:-1, 11] = (x_ref[1:, 11] - x_ref[:-1, 11]) / dt
    x_ref = np.vstack((x_ref, np.tile(xf, (N_k, 1))))
    amp = t_p / 4
    phase = np.pi * 3 / 2
    x_ref[:, 2] = [x_in[2] + amp + amp * np.sin(2 * np.pi / t_p * (i * dt) + phase) for i in range(t_ref)]
    x_ref[:-1, 6:9] = (x_ref[1:, 0:3] - x_ref[:-1, 0:3]) / dt
    Cmap = gait_map(t_ref, dt, t_start, 0.0, t_p, phi_switch)
    idx_pf = find_peaks(-x_ref[:, 2])[0] + step_adjustment
    idx_pf = np.hstack((0, idx_pf, t_ref - 1))
    edges = np.zeros(t_ref, dtype=np.int64)
    edges[1:] = (Cmap[:-1] == 1) & (Cmap[1:] == 0)
    kf = np.minimum(np.cumsum(edges), idx_pf.shape[0] - 1)
    pf_ref = np.zeros((t_ref, 3))
    pf_ref[1:, 0:2] = x_ref[idx_pf[kf[1:]], 0:2]
    return x_ref, pf_ref


def path_plan_grab(x_ref, k, N, mpc_factor):
    return x_ref[k:(k + N * mpc_factor):mpc_factor, :]


def run_clock(n_steps, dt, t_start):
    """Times the run loop sees: t_k = t_start + dt + ... + dt (k+1 additions), robotrunner.py:83,97."""
    t = np.empty(n_steps)
    acc = t_start
    for k in range(n_steps):
        acc = acc + dt
        t[k] = acc
    return t


def gate_masks(n_ticks, mpc_factor, dt, t_start, t_p=T_P, phi_switch=PHI_SWITCH, t0=0.0):
    """Scheduled contact at every simulator step, packed per MPC tick: bit i of entry j = gait_scheduler(t_k, t0) at
    k = mpc_factor j + i on the run clock (robotrunner.py:97-99: `t = t + self.dt`, `s = self.gait_scheduler(t, t0)`).
    This is the factor of the reference's commented-out gate `f_hist[k, :] = U[0, :]  # * s` (robotrunner.py:111).
    Returns (n_ticks,) uint32 (needs mpc_factor <= 32)."""
    if mpc_factor > 32:
        raise ValueError("gate masks hold one bit per simulator step: mpc_factor <= 32")
    s = gait_scheduler(run_clock(n_ticks * mpc_factor, dt, t_start), t0, t_p, phi_switch).reshape(n_ticks, mpc_factor)
    w = (np.uint64(1) << np.arange(mpc_factor, dtype=np.uint64))
    return ((s != 0).astype(np.uint64) * w).sum(axis=-1).astype(np.uint32)


def mpc_tables(x_ref, pf_ref, n_ticks, N, mpc_factor, dt, mpc_dt, t_start, t_p=T_P, phi_switch=PHI_SWITCH):
    """MPC-rate tables of one hopper for ``hmpc_rollout``.

    Returns xref_tab (n_ticks+N, 12), pf_tab (n_ticks+N+1, 3), C (n_ticks, N) and pf_switch (n_ticks,)
    such that pf_ref[20 j + i] == (pf_tab[j] if i < pf_switch[j] else pf_tab[j+1])."""
    rows = np.arange(n_ticks + N + 1) * mpc_factor
    rows = np.minimum(rows, x_ref.shape[0] - 1)
    xref_tab = x_ref[rows[:-1]]
    pf_tab = pf_ref[rows]
    t = run_clock(n_ticks * mpc_factor, dt, t_start)
    C = gait_map(N, mpc_dt, t[::mpc_factor], 0.0, t_p, phi_switch)
    sw = np.full(n_ticks, mpc_factor, dtype=np.uint8)
    for j in range(n_ticks):
        seg = pf_ref[j * mpc_factor:(j + 1) * mpc_factor]
        diff = np.any(seg != pf_tab[j], axis=1)
        if diff.any():
            s = int(np.argmax(diff))
            if not np.all(seg[s:] == pf_tab[j + 1]):
                raise ValueError("footstep changes more than once inside an MPC tick")
            sw[j] = s
    return xref_tab, pf_tab, C, sw


# ------------------------------------------------------------------------------------------------
# vectorised planner for synthetic batches (same formulas, all hoppers at once, MPC-rate rows only)
# ------------------------------------------------------------------------------------------------
def _parabola(y0, y1, y2, T, k):
    """Value at k of the parabola through (0,y0), (T/2,y1), (T,y2): what a 3-point not-a-knot
    CubicSpline reduces to (robotrunner.py:192-196)."""
    s = k / T
    return y0 * (1 - s) * (1 - 2 * s) + y1 * 4 * s * (1 - s) + y2 * s * (2 * s - 1)


def global_tables(N_run, N, max_tick, mpc_factor=20, dt=1e-3, mpc_dt=0.02, t_start=0.5 * T_P * PHI_SWITCH, t_p=T_P,
                  phi_switch=PHI_SWITCH, step_adjustment=STEP_ADJUSTMENT, n_sim=None):
    """Everything in the planner that depends only on the common clock, not on the hopper (the host half of the
    device-side planner, include/hmpc.h: hmpc_plan_set):
      sin_tab (n_sim,)   sine of the height wave at simulator step k           (robotrunner.py:210)
      pf_idx (n_sim,)    simulator index whose reference xy is the footstep in force at step k  (:214-224)
      Cglob (max_tick,N) contact flags of the MPC window of global tick j, cmask the same as bit masks (:97-102,172-180)
      sw_glob (max_tick,) simulator step inside tick j at which the footstep index changes (mpc_factor = never)
      gate_glob (max_tick,) uint32 scheduled contact at the simulator steps of tick j, one bit per step (gate_masks)
    with the reference's running float sums."""
    t_ref = N_run + N * mpc_factor
    if n_sim is None:
        n_sim = t_ref
    k = np.arange(n_sim).astype(float)
    sin_tab = np.sin(2 * np.pi / t_p * (k * dt) + np.pi * 3 / 2)
    cmap = gait_map(n_sim, dt, t_start, 0.0, t_p, phi_switch)
    edges = np.zeros(n_sim, dtype=np.int64)
    edges[1:] = (cmap[:-1] == 1) & (cmap[1:] == 0)
    period = int(round(t_p / dt))
    idx_pf = np.hstack((0, np.arange(period, t_ref - 1, period) + step_adjustment, t_ref - 1))
    kf = np.minimum(np.cumsum(edges), idx_pf.shape[0] - 1)
    pf_idx = idx_pf[kf]                                   # (n_sim,) sim index whose xy is the footstep
    # switch step inside each tick: pf_ref[20 (off+j) + i] == pf_tab[j] if i < sw else pf_tab[j+1]
    sw_glob = np.full(max_tick, mpc_factor, dtype=np.uint8)
    for J in range(max_tick):
        base = J * mpc_factor
        if base + mpc_factor > n_sim:
            break
        d = pf_idx[base:base + mpc_factor] != pf_idx[base]
        if d.any():
            sw_glob[J] = np.argmax(d)
    # MPC contact windows from the run clock t_k (robotrunner.py:97)
    tk = run_clock(max_tick * mpc_factor, dt, t_start)
    Cglob = gait_map(N, mpc_dt, tk[::mpc_factor], 0.0, t_p, phi_switch)      # (max_tick, N)
    w = (np.uint64(1) << np.arange(N, dtype=np.uint64))
    cmask = ((Cglob != 0).astype(np.uint64) * w).sum(axis=-1).astype(np.uint64)
    s45 = np.sin(45 * np.pi / 180)
    gate_glob = gate_masks(max_tick, mpc_factor, dt, t_start, t_p, phi_switch) if mpc_factor <= 32 else None
    return dict(sin_tab=sin_tab, pf_idx=pf_idx.astype(np.int32), sw_glob=sw_glob, Cglob=Cglob, cmask=cmask, gate_glob=gate_glob,
                curve_psi1=-0.4 * s45, curve_psi2=-s45, n_sim=n_sim, max_tick=max_tick, N_run=N_run, t_p=t_p)


def batch_tables(x0, xf, curve, tick_offset, N_run, n_ticks, N, mpc_factor=20, dt=1e-3, mpc_dt=0.02,
                 t_start=0.5 * T_P * PHI_SWITCH, t_p=T_P, phi_switch=PHI_SWITCH,
                 step_adjustment=STEP_ADJUSTMENT):
    """Planner + gait for B hoppers at once, MPC-rate rows only.

    Every hopper follows the reference's planner formulas (path_plan_init) between its own start x0[b]
    and goal xf[b] (B,12), optionally with the --curve yaw profile, and enters the run at its own tick
    ``tick_offset[b]`` (so the batch covers all gait phases at any instant).  The gait clock (common
    t_start, running sums) is the reference's.
    Returns dict of numpy arrays in the SoA layout of include/hmpc.h:
      xref_tab (n_ticks+N, 12, B), pf_tab (n_ticks+N+1, 3, B), C_tab (n_ticks, B) uint64,
      pf_switch (n_ticks, B) uint8, C (n_ticks, B, N) float, gate_tab (n_ticks, B) uint32 (gate_masks per hopper)."""
    x0 = np.asarray(x0, float); xf = np.asarray(xf, float)
    B = x0.shape[0]
    curve = np.asarray(curve, bool)
    off = np.asarray(tick_offset, dtype=np.int64)
    N_k = N * mpc_factor
    t_ref = N_run + N_k
    amp = t_p / 4
    max_tick = int(off.max()) + n_ticks
    step = (xf - x0) / (N_run - 1)
    s45 = np.sin(45 * np.pi / 180)

    def ref_rows(i):
        """x_ref[i[r, b]] for hopper b -> (R, B, 12), velocity columns (6:9) not filled."""
        ii = np.minimum(i, t_ref - 1).astype(float)[..., None]           # (R,B,1)
        k = ii[..., 0]
        lin = ii * step[None] + x0[None]
        lin = np.where(ii == N_run - 1, xf[None], lin)                    # linspace endpoint is exact
        out = lin.copy()
        T = float(N_run)
        out[..., 0] = np.where(curve[None], _parabola(x0[None, :, 1], 0.9 * xf[None, :, 1], xf[None, :, 1], T, k),
                               lin[..., 0])
        out[..., 5] = np.where(curve[None], _parabola(0.0, -0.4 * s45, -s45, T, k), lin[..., 5])
        nxt = np.where(k + 1 == N_run - 1, xf[None, :, 11], (k + 1) * step[None, :, 11] + x0[None, :, 11])
        d11 = np.where(k < N_run - 1, (nxt - lin[..., 11]) / dt, lin[..., 11])
        out[..., 11] = np.where(curve[None], d11, lin[..., 11])
        out = np.where((k >= N_run)[..., None], xf[None], out)
        out[..., 2] = x0[None, :, 2] + amp + amp * np.sin(2 * np.pi / t_p * (k * dt) + np.pi * 3 / 2)
        return out

    rows = np.minimum((off[None, :] + np.arange(n_ticks + N + 1)[:, None]) * mpc_factor, t_ref - 1)  # (R,B)
    r0 = ref_rows(rows)
    r1 = ref_rows(rows + 1)
    vel = (r1[..., 0:3] - r0[..., 0:3]) / dt
    r0[..., 6:9] = np.where((rows == t_ref - 1)[..., None], xf[None, :, 6:9], vel)   # last row keeps xf's
    xref_tab = r0[:-1]

    # everything that depends only on the common clock
    n_need = min(int(rows.max()) + 1, t_ref)
    gt = global_tables(N_run, N, max_tick, mpc_factor, dt, mpc_dt, t_start, t_p, phi_switch, step_adjustment, n_sim=n_need)
    pf_idx, sw_glob = gt["pf_idx"], gt["sw_glob"]
    sel = pf_idx[np.minimum(rows, n_need - 1)]            # (R,B)
    pf_tab = np.zeros(rows.shape + (3,))
    pf_tab[..., 0:2] = ref_rows(sel)[..., 0:2]
    jj = off[None, :] + np.arange(n_ticks)[:, None]       # (n_ticks,B) global tick index
    sw = sw_glob[jj]
    sw[np.all(pf_tab[:n_ticks] == pf_tab[1:n_ticks + 1], axis=-1)] = mpc_factor
    C = gt["Cglob"][jj]                                                       # (n_ticks,B,N)
    C_tab = gt["cmask"][jj]
    gate_tab = gt["gate_glob"][jj] if gt["gate_glob"] is not None else None      # (n_ticks,B) uint32 step masks
    return dict(xref_tab=np.ascontiguousarray(xref_tab.transpose(0, 2, 1)),
                pf_tab=np.ascontiguousarray(pf_tab.transpose(0, 2, 1)),
                C_tab=C_tab, pf_switch=sw, C=C, gate_tab=gate_tab)
