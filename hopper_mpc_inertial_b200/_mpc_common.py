"""Shared implementation of the drop-in ``Mpc`` classes (mpc_cvx_euler_3f.py:10-160, 2f:10-158)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .batch import BatchMpc, cbits_from_C


class _Value:
    """Stands in for a cvxpy Variable: only ``.value`` is used by callers (mpc_cvx_euler_3f.py:58,61)."""

    def __init__(self, shape):
        self.shape = shape
        self.value = None


class MpcBase:
    DYN = None

    def __init__(self, t, N, m, g, mu, Jinv, rh, **kwargs):
        self.t, self.N, self.m, self.g, self.mu = t, N, m, g, mu
        self.Jinv = np.asarray(Jinv, dtype=float)
        self.rh = np.asarray(rh, dtype=float)
        self.f_max = np.array([352, 0, 206])
        self.f_min = -self.f_max
        self.n_x, self.n_u = 12, 6
        # constant parts of the continuous-time model and the per-stage discrete matrices, as public attributes like
        # the reference's (mpc_cvx_euler_3f.py:24-33; 2f has no constant force block, mpc_cvx_euler_2f.py:24-32)
        self.A = np.zeros((12, 12))
        self.A[0:3, 6:9] = np.eye(3)
        self.B = np.zeros((12, 6))
        if self.DYN == "3f":
            self.B[6:9, 0:3] = np.eye(3) / m
        self.G = np.zeros(12)
        self.G[8] = -g
        self.Ad = np.zeros((N, 12, 12))
        self.Bd = np.zeros((N, 12, 6))
        self.Gd = self.G * t
        self.Q = np.diag([50., 50., 2., 1., 1., 50., 1., 1., 1., 10., 10., 10.])
        self.R = np.eye(6) * 0.001
        self.x = _Value((N + 1, 12))
        self.u = _Value((N, 6))
        self.status = None
        self.iters = None
        self._device = int(kwargs.pop("device", 0))
        self._overrides = dict(kwargs)
        self._bm = None
        self._key = None

    # Q / R are the reference's public gain attributes (mpc_cvx_euler_3f.py:34-37).  The device path condenses with
    # diagonal weights (hmpc_set_gains), so anything else is rejected at assignment time, not in the middle of a run.
    @staticmethod
    def _diag_only(M, n, name):
        M = np.array(M, dtype=float)
        if M.shape != (n, n):
            raise ValueError(f"{name} must be a {n}x{n} matrix")
        if np.any(M - np.diag(np.diag(M))):
            raise ValueError(f"{name} must be diagonal: the GPU path condenses with diagonal weights only")
        if np.any(np.diag(M) < 0):
            raise ValueError(f"{name} must be positive semidefinite")
        return M

    @property
    def Q(self):
        return self._Q

    @Q.setter
    def Q(self, M):
        self._Q = self._diag_only(M, 12, "Q")

    @property
    def R(self):
        return self._R

    @R.setter
    def R(self, M):
        self._R = self._diag_only(M, 6, "R")

    def gen_dt_dynamics(self, x, pf):
        """Fills ``self.Ad`` (N, 12, 12) and ``self.Bd`` (N, 12, 6) like mpc_cvx_euler_3f.py:71-94 (on the GPU)."""
        bm = self._backend()
        dev = bm.device
        xg = torch.as_tensor(np.ascontiguousarray(x, dtype=float)[:, :, None], device=dev)
        pft = torch.as_tensor(np.ascontiguousarray(pf, dtype=float)[:, :, None], device=dev)
        Ad, Bd = bm.linearize(xg.contiguous(), pft.contiguous())
        self.Ad = Ad[..., 0].cpu().numpy()
        self.Bd = Bd[..., 0].cpu().numpy()
        return None

    def _backend(self):
        key = (float(self.t), int(self.N), float(self.m), float(self.g), float(self.mu),
               self.Jinv.tobytes(), self.rh.tobytes(), float(self.f_max[2]), repr(sorted(self._overrides.items(), key=lambda kv: kv[0])))
        if self._bm is None or key != self._key:
            if self._bm is not None:
                self._bm.close()
            ov = dict(mpc_dt=float(self.t), m=float(self.m), g=float(self.g), mu=float(self.mu),
                      Jinv=self.Jinv.reshape(-1), J=np.linalg.inv(self.Jinv).reshape(-1),
                      rh=self.rh, fz_max=float(self.f_max[2]))
            ov.update(self._overrides)
            self._bm = BatchMpc(1, dyn=self.DYN, N=self.N, device=self._device, **ov)
            self._key = key
        return self._bm

    def _push_gains(self, bm):
        # in-place edits (mpc.Q[0, 1] = ...) bypass the setters: validate again
        Q, R = self._diag_only(self._Q, 12, "Q"), self._diag_only(self._R, 6, "R")
        qd = torch.as_tensor(np.diag(Q).copy()[:, None], device=bm.device).contiguous()
        rd = torch.as_tensor(np.diag(R).copy()[:, None], device=bm.device).contiguous()
        bm.set_gains(qd, rd)

    def mpcontrol(self, x_in, x_ref_in, pf, C, init):
        """Same contract as the reference (mpc_cvx_euler_3f.py:41-69): returns u (N, 6) float64."""
        bm = self._backend()
        self._push_gains(bm)
        dev = bm.device
        N = self.N
        xi = torch.as_tensor(np.asarray(x_in, float).reshape(12, 1).copy(), device=dev)
        xr = torch.as_tensor(np.asarray(x_ref_in, float).reshape(N, 12, 1).copy(), device=dev)
        pft = torch.as_tensor(np.asarray(pf, float).reshape(N, 3, 1).copy(), device=dev)
        cb = torch.as_tensor(cbits_from_C(np.asarray(C).reshape(1, N)).view(np.int64), device=dev)
        U, Xs, st, it = bm.solve(xi, xr, pft, cb, bool(init))
        st = int(st.item())
        self.status, self.iters = st, int(it.item())
        if st in (_lib.STATUS_INFEASIBLE, _lib.STATUS_NON_FINITE, _lib.STATUS_MAX_ITER):
            self.u.value = None
            self.x.value = None
            raise Exception("\n *** QP FAILED *** \n")   # mpc_cvx_euler_3f.py:158-159
        self.u.value = U[..., 0].cpu().numpy()
        self.x.value = Xs[..., 0].cpu().numpy()
        return self.u.value
