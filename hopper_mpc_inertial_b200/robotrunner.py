"""Drop-in for the reference's robotrunner (robotrunner.py:19-230): same ``Runner`` constructor,
``run()``, ``convert`` -- with the MPC and the RK4 simulator running on the GPU.

``Runner.run`` follows the reference loop tick by tick through ``Mpc.mpcontrol`` (so the drop-in Mpc
is exercised exactly like the reference's) and keeps the reference's histories (``X_traj``, ``f_hist``,
``s_hist``).  ``Runner.run_fused`` runs the same closed loop through the fused ``hmpc_rollout`` path
(state resident in HBM) -- that is the path the batch API uses.  Plotting is optional and headless-safe
(SURVEY App. D11): the reference's blocking matplotlib calls are replaced by ``self.plot()``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import mpc_cvx_euler_2f, mpc_cvx_euler_3f, planner
from .batch import cbits_from_C
from .utils import H, L, R, quat2euler

try:
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(x, **k):
        return x

np.set_printoptions(suppress=True, linewidth=np.nan)


def convert(X_in):
    """SE(3) 13-state -> Euler 12-state on the host (robotrunner.py:19-28); used for the planner's
    end points only -- the per-tick conversion runs on the GPU (hmpc_convert)."""
    X_in = np.asarray(X_in, float)
    q = X_in[3:7]
    Rm = H.T @ L(q) @ R(q).T @ H                  # rotation body -> world (robotrunner.py:22)
    x0 = np.zeros(12)
    x0[0:3] = X_in[0:3]
    x0[3:6] = quat2euler(q)
    x0[6:9] = Rm @ X_in[7:10]
    x0[9:] = Rm @ X_in[10:13]
    return x0


class Runner:
    def __init__(self, dt=1e-3, dyn='2f', curve=False, N_run=5000, N=60, device=0, progress=True,
                 contact_gate="off", leg_max=None, **mpc_kwargs):
        self.dt, self.N_run, self.curve, self.dyn = dt, N_run, curve, dyn
        # contact gate of the applied control (robotrunner.py:111 `f_hist[k, :] = U[0, :]  # * s`): "off" = the
        # reference as shipped, "schedule" = the commented-out factor switched on, "detect" = leg reach <= leg_max
        if contact_gate not in ("off", "schedule", "detect"):
            raise ValueError("contact_gate must be 'off', 'schedule' or 'detect'")
        if contact_gate == "detect" and not (leg_max and leg_max > 0):
            raise ValueError("contact_gate='detect' needs leg_max > 0")
        self.contact_gate, self.leg_max = contact_gate, leg_max
        self.m = 7.5
        self.J = np.array([[76148072.89, 70089.52, 2067970.36],
                           [70089.52, 45477183.53, -87045.58],
                           [2067970.36, -87045.58, 76287220.47]]) * (10 ** (-9))
        self.Jinv = np.linalg.inv(self.J)
        self.rh = -np.array([0.02663114, 0.04435752, 6.61082088]) / 1000
        self.g = 9.807
        self.t_p = 0.8
        self.phi_switch = 0.5
        self.N = N                                   # robotrunner.py:46 uses 60
        self.mpc_dt = 0.02
        self.mpc_factor = int(self.mpc_dt / self.dt)
        self.N_time = self.N * self.mpc_dt
        self.N_k = int(self.N * self.mpc_factor)
        self.n_X, self.n_U = 13, 6
        self.dist = 0.4 * (N_run * dt)
        self.X_0 = np.array([0, 0, 0.27, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0], dtype=float)
        self.X_f = np.hstack([self.dist, 0, 0.27, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0]).T
        mu = 1
        mpc_dyn = {'2f': mpc_cvx_euler_2f, '3f': mpc_cvx_euler_3f}[dyn]
        # the device integrates with h = sim_dt and runs mpc_factor steps per tick in the fused path: both follow
        # this Runner's dt (robotrunner.py:48,154-164)
        if not 1 <= self.mpc_factor <= 255:
            raise ValueError("dt must give 1..255 simulator steps per MPC tick (mpc_dt / dt)")
        mpc_kwargs.setdefault("sim_dt", float(self.dt))
        mpc_kwargs.setdefault("mpc_factor", int(self.mpc_factor))
        self.mpc = mpc_dyn.Mpc(t=self.mpc_dt, N=self.N, m=self.m, g=self.g, mu=mu, Jinv=self.Jinv,
                               rh=self.rh, device=device, **mpc_kwargs)
        self.t_start = 0.5 * self.t_p * self.phi_switch
        self.step_adjustment = -115
        self.progress = progress
        self.X_traj = self.f_hist = self.s_hist = self.x_ref = self.pf_ref = None

    # -- planner / gait: host precompute (robotrunner.py:166-230) -------------------------------
    def gait_scheduler(self, t, t0):
        return int(planner.gait_scheduler(t, t0, self.t_p, self.phi_switch))

    def gait_map(self, N, dt, ts, t0):
        return planner.gait_map(N, dt, ts, t0, self.t_p, self.phi_switch)

    def path_plan_init(self, x_in, xf):
        return planner.path_plan_init(x_in, xf, self.N_run, self.N, self.mpc_factor, self.dt, self.curve,
                                      self.t_start, self.t_p, self.phi_switch, self.step_adjustment)

    def path_plan_grab(self, x_ref, k):
        return planner.path_plan_grab(x_ref, k, self.N, self.mpc_factor)

    # -- simulator on the GPU -----------------------------------------------------------------
    def rk4_normalized(self, xk, uk, pfk):
        bm = self.mpc._backend()
        X = torch.as_tensor(np.asarray(xk, float).reshape(13, 1).copy(), device=bm.device)
        U = torch.as_tensor(np.asarray(uk, float).reshape(6, 1).copy(), device=bm.device)
        pf = torch.as_tensor(np.asarray(pfk, float).reshape(3, 1).copy(), device=bm.device)
        bm.rk4(X, U, pf, 1)
        return X[:, 0].cpu().numpy()

    def run(self, plot=False):
        """The reference's closed loop (robotrunner.py:81-124), MPC + RK4 on the GPU."""
        if self.contact_gate == "detect":
            raise ValueError("contact_gate='detect' reads the state at every simulator step: use run_fused()")
        N_run = self.N_run + 1
        t = self.t_start
        t0 = 0
        mpc_factor = self.mpc_factor
        X_traj = np.tile(self.X_0, (N_run, 1))
        f_hist = np.zeros((N_run, self.n_U))
        s_hist = np.zeros(N_run)
        x_ref, pf_ref = self.path_plan_init(x_in=convert(X_traj[0, :]), xf=convert(self.X_f))
        bm = self.mpc._backend()
        dev = bm.device
        Xd = torch.as_tensor(self.X_0.reshape(13, 1).copy(), device=dev)
        init = True
        it = range(0, self.N_run, mpc_factor)
        for k0 in (tqdm(it) if self.progress else it):
            # clock exactly as the reference: t += dt before every sim step
            ts = []
            for _ in range(min(mpc_factor, self.N_run - k0)):
                t = t + self.dt
                ts.append(t)
            C = self.gait_map(self.N, self.mpc_dt, ts[0], t0)
            x_refk = self.path_plan_grab(x_ref, k0)
            pf_refk = self.path_plan_grab(pf_ref, k0)
            x_in = bm.convert(Xd)[:, 0].cpu().numpy()
            U = self.mpc.mpcontrol(x_in=x_in, x_ref_in=x_refk, pf=pf_refk, C=C, init=init)
            init = False
            Ud = torch.as_tensor(U[0].reshape(6, 1).copy(), device=dev)
            nst = len(ts)
            s_k = np.array([self.gait_scheduler(tt, t0) for tt in ts], float)
            gate = s_k if self.contact_gate == "schedule" else np.ones(nst)
            # integrate the tick; split where the footstep reference (or the contact gate) changes
            i = 0
            while i < nst:
                j = i + 1
                while j < nst and np.all(pf_ref[k0 + j] == pf_ref[k0 + i]) and gate[j] == gate[i]:
                    j += 1
                pfd = torch.as_tensor(pf_ref[k0 + i].reshape(3, 1).copy(), device=dev)
                Xs = bm.rk4(Xd, Ud * float(gate[i]), pfd, j - i, log_steps=True)
                X_traj[k0 + i + 1:k0 + j + 1] = Xs[:, :, 0].cpu().numpy()
                i = j
            f_hist[k0:k0 + nst] = U[0][None, :] * gate[:, None]
            s_hist[k0:k0 + nst] = s_k
        self.X_traj, self.f_hist, self.s_hist, self.x_ref, self.pf_ref = X_traj, f_hist, s_hist, x_ref, pf_ref
        if plot:
            self.plot()
        return None

    def run_fused(self):
        """Same closed loop through hmpc_rollout (one launch pair per tick, state stays in HBM).
        Returns X at tick boundaries (n_ticks+1, 13) and the applied control per tick (n_ticks, 6)."""
        n_ticks = self.N_run // self.mpc_factor
        x_ref, pf_ref = self.path_plan_init(x_in=convert(self.X_0), xf=convert(self.X_f))
        xt, pt, C, sw = planner.mpc_tables(x_ref, pf_ref, n_ticks, self.N, self.mpc_factor, self.dt,
                                           self.mpc_dt, self.t_start)
        bm = self.mpc._backend()
        self.mpc._push_gains(bm)
        dev = bm.device
        X = torch.as_tensor(self.X_0.reshape(13, 1).copy(), device=dev)
        if self.contact_gate == "schedule":
            gm = planner.gate_masks(n_ticks, self.mpc_factor, self.dt, self.t_start, self.t_p, self.phi_switch)
            bm.set_contact_gate("schedule", gate_tab=torch.as_tensor(gm.view(np.int32).reshape(n_ticks, 1).copy(), device=dev))
        elif self.contact_gate == "detect":
            bm.set_contact_gate("detect", leg_max=float(self.leg_max))
        else:
            bm.set_contact_gate("off")
        out = bm.rollout(X, torch.as_tensor(xt[:, :, None].copy(), device=dev),
                         torch.as_tensor(pt[:, :, None].copy(), device=dev),
                         torch.as_tensor(cbits_from_C(C).view(np.int64).reshape(n_ticks, 1).copy(), device=dev),
                         torch.as_tensor(sw.reshape(n_ticks, 1).copy(), device=dev),
                         tick0=0, n_ticks=n_ticks, init=True, log=True)
        st = int(out["status"].item())
        if st in (1, 2, 3):
            raise Exception("\n *** QP FAILED *** \n")
        self.x_ref, self.pf_ref = x_ref, pf_ref
        return out["X_log"][:, :, 0].cpu().numpy(), out["U_log"][:, :, 0].cpu().numpy()

    def plot(self):
        """Optional visualisation (reference: plots.py); silently skipped when matplotlib is absent."""
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except Exception:
            print("matplotlib not available: skipping plots")
            return
        fig, ax = plt.subplots(3, 1, sharex=True)
        for i, name in enumerate("xyz"):
            ax[i].plot(self.X_traj[:, i], label=name)
            ax[i].plot(self.x_ref[:self.X_traj.shape[0], i], "--", label=name + " ref")
            ax[i].legend()
        fig.savefig("hopper_run.png")
