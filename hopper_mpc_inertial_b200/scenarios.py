"""Synthetic batches of independent hoppers (SURVEY 8d "Configs -> synthetic inputs").

Every per-hopper quantity is drawn from a counter-based generator keyed by the GLOBAL hopper index
(numpy Philox, key = seed, counter = hopper index), so a hopper's scenario does not depend on how the
batch is sharded across GPUs.
"""
from __future__ import annotations

import numpy as np

from . import planner

Q_REF = np.array([50., 50., 2., 1., 1., 50., 1., 1., 1., 10., 10., 10.])
R_REF = np.full(6, 0.001)


def _uniforms(seed, idx0, count, k):
    """(count, k) uniforms in [0,1): row i depends only on (seed, idx0 + i)."""
    out = np.empty((count, k))
    # Philox is counter based: advance() jumps straight to a hopper's private stream segment
    chunk = 4 * ((k + 3) // 4)
    bg = np.random.Philox(key=seed)
    bg.advance(int(idx0) * (chunk // 4))   # one Philox counter step yields 4 x 64 bits
    gen = np.random.Generator(bg)
    out[:] = gen.random((count, chunk))[:, :k]
    return out


def make_batch(B, idx0=0, seed=1234, N=10, n_ticks=100, N_run=2000, mpc_factor=20, dt=1e-3, mpc_dt=0.02,
               randomize_gains=True):
    """Scenario for hoppers idx0 .. idx0+B-1.  Returns numpy arrays in the SoA layout of include/hmpc.h:
    X0 (13,B), Qdiag (12,B), Rdiag (6,B), xref_tab, pf_tab, C_tab (uint64), pf_switch (uint8), C."""
    u = _uniforms(seed, idx0, B, 40)
    c = 0

    def take(n):
        nonlocal c
        v = u[:, c:c + n]
        c += n
        return v

    def rng(lo, hi, n):
        return lo + (hi - lo) * take(n)

    p_xy = rng(-0.5, 0.5, 2)
    z0 = rng(0.22, 0.45, 1)
    rp = rng(-0.1, 0.1, 2)
    yaw = rng(-np.pi / 4, np.pi / 4, 1)
    v_b = rng(-0.5, 0.5, 3)
    w_b = rng(-0.5, 0.5, 3)
    gq = np.exp(rng(np.log(0.5), np.log(2.0), 12)) if randomize_gains else np.ones((B, 12))
    gr = np.exp(rng(np.log(0.5), np.log(2.0), 6)) if randomize_gains else np.ones((B, 6))
    if not randomize_gains:
        c += 18
    speed = rng(0.2, 1.0, 1)[:, 0]
    heading = rng(-np.pi, np.pi, 1)[:, 0]
    curve = take(1)[:, 0] < 0.5
    t_start = rng(0.0, planner.T_P, 1)[:, 0]

    # quaternion from ZYX Euler angles (roll, pitch, yaw)
    cr, sr = np.cos(rp[:, 0] / 2), np.sin(rp[:, 0] / 2)
    cp, sp = np.cos(rp[:, 1] / 2), np.sin(rp[:, 1] / 2)
    cy, sy = np.cos(yaw[:, 0] / 2), np.sin(yaw[:, 0] / 2)
    q = np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                  cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], axis=1)
    X0 = np.concatenate([p_xy, z0, q, v_b, w_b], axis=1)          # (B,13)

    # planner end points: reference-style start (upright, at rest, nominal height) and goal
    T = N_run * dt
    x0p = np.zeros((B, 12)); xfp = np.zeros((B, 12))
    x0p[:, 0:2] = p_xy; x0p[:, 2] = 0.27
    dist = speed * T * 0.4 / 0.4
    xfp[:, 0] = p_xy[:, 0] + dist * np.cos(heading)
    xfp[:, 1] = p_xy[:, 1] + dist * np.sin(heading)
    xfp[:, 2] = 0.27
    tabs = planner.batch_tables(x0p, xfp, curve, t_start, N_run, n_ticks, N, mpc_factor, dt, mpc_dt)
    out = dict(X0=np.ascontiguousarray(X0.T), Qdiag=np.ascontiguousarray((Q_REF[None] * gq).T),
               Rdiag=np.ascontiguousarray((R_REF[None] * gr).T), curve=curve, t_start=t_start)
    out.update(tabs)
    return out
