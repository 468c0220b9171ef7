"""CPU oracle for the hopper-MPC hot path.  *** TEST INFRASTRUCTURE ONLY ***

This module restates, in plain numpy FP64, the arithmetic of the reference's closed-loop MPC hot
path (bbokser/hopper-mpc-inertial).  It is the *checker* for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  The product package (``hopper_mpc_inertial_b200``) never does.

PARITY PINNING.  The reference ships no tests, golden vectors or fixtures, and its QP back end
(cvxpy -> OSQP) and ``transforms3d`` are not installed in this image (no network).  The pins are:
  * every function that can run (simulator, convert, gait, planner, ``gen_dt_dynamics``) is compared
    against the reference's own modules imported through ``oracle/refshim.py`` (in the build
    container only) and frozen as fixtures under ``tests/golden/`` by ``oracle/make_golden.py``;
  * the QP *data* (P, q, A, l, u) are pinned by executing the reference's own ``build_qp`` through the
    mini-cvxpy shim ``oracle/minicvx.py``;
  * the QP *solution* is pinned by uniqueness: the condensed Hessian is positive definite, so the
    optimum is unique and is certified solver-independently by ``kkt_certificate``.
What cannot be pinned is OSQP's own iterate at eps=1e-5 (binary absent) -> "parity unpinned" for
that one aspect; see DESIGN.md.

Reference citations are ``file:line`` into ``/root/reference/src``.
"""
from __future__ import annotations

import dataclasses
import numpy as np

# --------------------------------------------------------------------------------------------
# constants (robotrunner.py:37-59,68,78-79)
# --------------------------------------------------------------------------------------------
J_REF = np.array([[76148072.89, 70089.52, 2067970.36],
                  [70089.52, 45477183.53, -87045.58],
                  [2067970.36, -87045.58, 76287220.47]]) * (10 ** (-9))
RH_REF = -np.array([0.02663114, 0.04435752, 6.61082088]) / 1000
Q_DIAG_REF = np.array([50., 50., 2., 1., 1., 50., 1., 1., 1., 10., 10., 10.])  # mpc_cvx_euler_3f.py:35
R_DIAG_REF = np.full(6, 0.001)                                                # mpc_cvx_euler_3f.py:37
TAU_MAX_REF = np.array([7.78, 7.78, 4.0])                                     # mpc_cvx_euler_3f.py:123-128
FZ_MAX_REF = 206.0                                                            # mpc_cvx_euler_3f.py:20,146
Z_MIN_REF = 0.1                                                               # mpc_cvx_euler_3f.py:129
KF_TERMINAL = 100.0                                                           # mpc_cvx_euler_3f.py:113
INF = 1e30  # "no bound" marker shared with the CUDA side (include/hmpc.h HMPC_INF)


@dataclasses.dataclass
class Params:
    """Physical + MPC constants of one hopper (robotrunner.py:37-59, mpc_cvx_euler_3f.py:12-37)."""
    dyn: str = "3f"
    N: int = 60
    mpc_dt: float = 0.02
    sim_dt: float = 1e-3
    mpc_factor: int = 20
    m: float = 7.5
    g: float = 9.807
    mu: float = 1.0
    J: np.ndarray = dataclasses.field(default_factory=lambda: J_REF.copy())
    rh: np.ndarray = dataclasses.field(default_factory=lambda: RH_REF.copy())
    Qdiag: np.ndarray = dataclasses.field(default_factory=lambda: Q_DIAG_REF.copy())
    Rdiag: np.ndarray = dataclasses.field(default_factory=lambda: R_DIAG_REF.copy())
    tau_max: np.ndarray = dataclasses.field(default_factory=lambda: TAU_MAX_REF.copy())
    fz_max: float = FZ_MAX_REF
    z_min: float = Z_MIN_REF
    kf: float = KF_TERMINAL
    uref_mode: str = "aliased"  # "aliased" (reference-faithful) | "per_stage"
    t_p: float = 0.8
    phi_switch: float = 0.5

    @property
    def Jinv(self):
        return np.linalg.inv(self.J)


# --------------------------------------------------------------------------------------------
# utils.py restatements
# --------------------------------------------------------------------------------------------
def hat(w):
    """Skew-symmetric cross-product matrix (utils.py:21-25)."""
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def rz(phi):
    """Linearised yaw rotation, world->body, i.e. Rz(phi)^T (utils.py:46-51)."""
    c, s = np.cos(phi), np.sin(phi)
    return np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])


def quat_rotm(q):
    """Body->world rotation H^T L(q) R(q)^T H used at robotrunner.py:25-27,140-147.

    For a (possibly un-normalised) quaternion q=(w,v) the product L(q) R(q)^T restricted to the
    vector part equals  (w^2 - v.v) I + 2 v v^T + 2 w hat(v)  -- no division by |q|^2
    (utils.py:28-43), which matters inside RK4 stages where q is not unit.
    """
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


_EPS4 = np.finfo(float).eps * 4.0


def quat2euler(q):
    """ZYX Euler angles returned as [roll, pitch, yaw] (utils.py:54-62).

    Restates transforms3d.euler.quat2euler(q, axes='rzyx') [EXT, package absent]: normalising
    rotation matrix (s = 2/|q|^2), cy = hypot(M00, M10), atan2 extraction.
    """
    w, x, y, z = q
    nq = w * w + x * x + y * y + z * z
    if nq < np.finfo(float).eps:
        return np.zeros(3)
    s = 2.0 / nq
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ = w * X, w * Y, w * Z
    xX, xY, xZ = x * X, x * Y, x * Z
    yY, yZ, zZ = y * Y, y * Z, z * Z
    m00 = 1.0 - (yY + zZ)
    m10 = xY + wZ
    m20 = xZ - wY
    m21 = yZ + wX
    m22 = 1.0 - (xX + yY)
    m11 = 1.0 - (xX + zZ)
    m12 = yZ - wX
    cy = np.sqrt(m00 * m00 + m10 * m10)
    if cy > _EPS4:
        roll = np.arctan2(m21, m22)
        pitch = np.arctan2(-m20, cy)
        yaw = np.arctan2(m10, m00)
    else:
        roll = np.arctan2(-m12, m11)
        pitch = np.arctan2(-m20, cy)
        yaw = 0.0
    return np.array([roll, pitch, yaw])


# --------------------------------------------------------------------------------------------
# simulator (robotrunner.py:19-28,126-164)
# --------------------------------------------------------------------------------------------
def convert(X):
    """SE(3) 13-state -> Euler 12-state (robotrunner.py:19-28)."""
    Rm = quat_rotm(X[3:7])
    x = np.empty(12)
    x[0:3] = X[0:3]
    x[3:6] = quat2euler(X[3:7])
    x[6:9] = Rm @ X[7:10]
    x[9:12] = Rm @ X[10:13]
    return x


def dynamics_ct(X, U, pf, prm: Params):
    """Continuous-time single-rigid-body dynamics (robotrunner.py:126-152).

    U[0:3] is always applied as a world-frame force (robotrunner.py:137), also for 2f (App. D5)."""
    p, q, v, w = X[0:3], X[3:7], X[7:10], X[10:13]
    Fw, tau = U[0:3], U[3:6]
    Rm = quat_rotm(q)
    Fg = np.array([0.0, 0.0, -prm.g]) * prm.m
    Ftb = Rm.T @ (Fg + Fw)
    r = prm.rh + Rm.T @ (pf - p)
    Fb = Rm.T @ Fw
    tau_tot = tau + np.cross(r, Fb)
    dp = Rm @ v
    qw, qv = q[0], q[1:4]
    dq = 0.5 * np.concatenate(([-qv @ w], qw * w + np.cross(qv, w)))  # 0.5 L(q) H w
    dv = Ftb / prm.m - np.cross(w, v)
    dw = np.linalg.solve(prm.J, tau_tot - np.cross(w, prm.J @ w))
    return np.concatenate((dp, dq, dv, dw))


def rk4_normalized(X, U, pf, prm: Params):
    """Classic RK4 with h = sim_dt, then quaternion renormalisation (robotrunner.py:154-164)."""
    h = prm.sim_dt
    f1 = dynamics_ct(X, U, pf, prm)
    f2 = dynamics_ct(X + 0.5 * h * f1, U, pf, prm)
    f3 = dynamics_ct(X + 0.5 * h * f2, U, pf, prm)
    f4 = dynamics_ct(X + h * f3, U, pf, prm)
    Xn = X + (h / 6.0) * (f1 + 2 * f2 + 2 * f3 + f4)
    Xn[3:7] = Xn[3:7] / np.linalg.norm(Xn[3:7])
    return Xn


# --------------------------------------------------------------------------------------------
# gait (robotrunner.py:166-180) -- accumulated float sums must be reproduced exactly (App. D6)
# --------------------------------------------------------------------------------------------
def gait_scheduler(t, t0, prm: Params):
    phi = np.mod((t - t0) / prm.t_p, 1)
    return 0 if phi > prm.phi_switch else 1


def gait_map(N, dt, ts, t0, prm: Params):
    C = np.zeros(N)
    for k in range(N):
        C[k] = gait_scheduler(ts, t0, prm)
        ts += dt
    return C


# --------------------------------------------------------------------------------------------
# path planner (robotrunner.py:182-230); scipy CubicSpline / find_peaks are present in the image
# --------------------------------------------------------------------------------------------
def path_plan_init(x_in, xf, prm: Params, N_run, curve, t_start, step_adjustment=-115):
    from scipy.interpolate import CubicSpline
    from scipy.signal import find_peaks
    N_k = prm.N * prm.mpc_factor
    dt = prm.sim_dt
    t_traj = int(N_run)
    t_ref = N_run + N_k
    x_ref = np.linspace(start=x_in, stop=xf, num=t_traj)
    if curve:
        st = np.array([0, t_traj * 0.5, t_traj])
        csy = CubicSpline(st, np.array([x_in[1], xf[1] * 0.9, xf[1]]))
        s45 = np.sin(45 * np.pi / 180)
        cspsi = CubicSpline(st, np.array([0, -s45 * 0.4, -s45]))
        kk = np.arange(t_traj)
        x_ref[:, 0] = csy(kk)      # App. D4: y-spline is written into column 0
        x_ref[:, 5] = cspsi(kk)
        x_ref[:-1, 11] = (x_ref[1:N_run, 11] - x_ref[0:N_run - 1, 11]) / dt  # stays 0 (D4)
    x_ref = np.vstack((x_ref, np.tile(xf, (N_k, 1))))
    amp = prm.t_p / 4
    ii = np.arange(t_ref)
    x_ref[:, 2] = [x_in[2] + amp + amp * np.sin(2 * np.pi / prm.t_p * (i * dt) + np.pi * 3 / 2)
                   for i in range(t_ref)]
    x_ref[:-1, 6:9] = (x_ref[1:, 0:3] - x_ref[:-1, 0:3]) / dt
    C = gait_map(t_ref, dt, t_start, 0, prm)
    idx_pf = find_peaks(-x_ref[:, 2])[0] + step_adjustment
    idx_pf = np.hstack((0, idx_pf, t_ref - 1))
    pf_ref = np.zeros((t_ref, 3))
    kf = 0
    n_idx = idx_pf.shape[0]
    for k in range(1, t_ref):
        if C[k - 1] == 1 and C[k] == 0 and kf < n_idx:
            kf += 1
        pf_ref[k, 0:2] = x_ref[idx_pf[kf], 0:2]
    return x_ref, pf_ref


def path_plan_grab(x_ref, k, prm: Params):
    return x_ref[k:(k + prm.N * prm.mpc_factor):prm.mpc_factor, :]


# --------------------------------------------------------------------------------------------
# linearisation (mpc_cvx_euler_3f.py:71-94, mpc_cvx_euler_2f.py:70-94)
# --------------------------------------------------------------------------------------------
def gen_dt_dynamics(x_guess, pf, prm: Params):
    N, dt = prm.N, prm.mpc_dt
    Jinv = prm.Jinv
    Ad = np.zeros((N, 12, 12))
    Bd = np.zeros((N, 12, 6))
    for k in range(N):
        Rz = rz(x_guess[k, 5])
        rf = prm.rh + Rz @ (pf[k] - x_guess[k, 0:3])
        Jw_inv = Rz @ Jinv @ Rz.T
        A = np.zeros((12, 12))
        B = np.zeros((12, 6))
        A[0:3, 6:9] = np.eye(3)
        A[3:6, 9:12] = Rz
        if prm.dyn == "3f":
            B[6:9, 0:3] = np.eye(3) / prm.m
            B[9:12, 0:3] = Jw_inv @ hat(Rz.T @ rf)
        else:
            B[6:9, 0:3] = Rz.T / prm.m
            B[9:12, 0:3] = Jw_inv @ Rz.T @ hat(rf)
        B[9:12, 3:6] = Jw_inv @ Rz.T
        Ad[k] = np.eye(12) + A * dt
        Bd[k] = B * dt
    Gd = np.zeros(12)
    Gd[8] = -prm.g * dt
    return Ad, Bd, Gd


def uref_z(C, prm: Params):
    """Effective per-stage fz reference (App. D1).  aliased: last-written value for every stage."""
    N = prm.N
    if prm.uref_mode == "aliased":
        return np.full(N, 2 * prm.m * prm.g if C[N - 1] != 0 else 0.0)
    return np.where(np.asarray(C) != 0, 2 * prm.m * prm.g, 0.0)


# --------------------------------------------------------------------------------------------
# QP assembly, full (cvxpy-shaped) form -- SURVEY App. A, from mpc_cvx_euler_3f.py:96-153
#   variables v = [x(0..N) row-major (12 each) ; u(0..N-1) (6 each)];  min 1/2 v'Pv + q'v,  l<=Av<=u
# --------------------------------------------------------------------------------------------
def build_qp_full(x_in, x_ref, Ad, Bd, Gd, C, prm: Params):
    N = prm.N
    nx, nu = 12 * (N + 1), 6 * N
    nv = nx + nu
    Pd = np.zeros(nv)
    q = np.zeros(nv)
    ubar = uref_z(C, prm)
    const = 0.0
    for k in range(N):
        kf = prm.kf if k == N - 1 else 1.0
        kuf = 0.0 if k == N - 1 else 1.0
        ix = 12 * (k + 1)
        Pd[ix:ix + 12] = 2 * prm.Qdiag * kf
        q[ix:ix + 12] = -2 * prm.Qdiag * kf * x_ref[k]
        iu = nx + 6 * k
        Pd[iu:iu + 6] = 2 * prm.Rdiag * kuf
        q[iu + 2] = -2 * prm.Rdiag[2] * kuf * ubar[k]
        const += kf * x_ref[k] @ (prm.Qdiag * x_ref[k]) + kuf * prm.Rdiag[2] * ubar[k] ** 2
    rows, lo, hi = [], [], []

    def add(coefs, l, u):
        r = np.zeros(nv)
        for i, c in coefs:
            r[i] += c
        rows.append(r); lo.append(l); hi.append(u)

    mu = prm.mu
    for k in range(N):
        iu = nx + 6 * k
        fx, fy, fz = iu, iu + 1, iu + 2
        for a in range(3):
            add([(iu + 3 + a, 1.0)], -prm.tau_max[a], prm.tau_max[a])
        add([(12 * k + 2, 1.0)], prm.z_min, INF)
        # dynamics: x[k+1] - Ad x[k] - Bd u[k] = Gd
        for r_ in range(12):
            co = [(12 * (k + 1) + r_, 1.0)]
            co += [(12 * k + c_, -Ad[k, r_, c_]) for c_ in range(12) if Ad[k, r_, c_] != 0.0]
            co += [(iu + c_, -Bd[k, r_, c_]) for c_ in range(6) if Bd[k, r_, c_] != 0.0]
            add(co, Gd[r_], Gd[r_])
        if prm.dyn == "2f":
            add([(fy, 1.0)], 0.0, 0.0)
        if C[k] == 0:
            add([(fx, 1.0)], 0.0, 0.0)
            if prm.dyn == "3f":
                add([(fy, 1.0)], 0.0, 0.0)
            add([(fz, 1.0)], 0.0, 0.0)
        else:
            add([(fx, 1.0), (fz, -mu)], -INF, 0.0)
            add([(fx, -1.0), (fz, -mu)], -INF, 0.0)
            if prm.dyn == "3f":
                add([(fy, 1.0), (fz, -mu)], -INF, 0.0)
                add([(fy, -1.0), (fz, -mu)], -INF, 0.0)
            add([(fz, 1.0)], 0.0, prm.fz_max)
    for r_ in range(12):
        add([(r_, 1.0)], x_in[r_], x_in[r_])
    return dict(P=np.diag(Pd), q=q, A=np.array(rows), l=np.array(lo), u=np.array(hi), const=const,
                nx=nx, nu=nu)


# --------------------------------------------------------------------------------------------
# QP assembly, condensed (inputs-only) form -- SURVEY App. A "Condensed form"
#   X = c + S U ;  H = 2(S'QS + R) ; g = 2(S'Q(c - xref) - R ubar)
#   rows: [identity box on 6N inputs ; 4 friction slots per stage ; 1 height slot per stage]
#   m = 11 N in a FIXED slot layout (unused slots carry l=-INF,u=+INF and a zero/idle row) so that
#   the CUDA kernel is branch-free; this layout is the one include/hmpc.h documents.
# --------------------------------------------------------------------------------------------
def free_response_and_S(x_in, Ad, Bd, Gd, prm: Params):
    N = prm.N
    c = np.zeros((N + 1, 12))
    S = np.zeros((N + 1, 12, 6 * N))
    c[0] = x_in
    for k in range(N):
        c[k + 1] = Ad[k] @ c[k] + Gd
        S[k + 1] = Ad[k] @ S[k]
        S[k + 1][:, 6 * k:6 * k + 6] += Bd[k]
    return c, S


def build_qp_condensed(x_in, x_ref, Ad, Bd, Gd, C, prm: Params):
    N = prm.N
    n = 6 * N
    c, S = free_response_and_S(x_in, Ad, Bd, Gd, prm)
    ubar = uref_z(C, prm)
    Hm = np.zeros((n, n))
    g = np.zeros(n)
    for k in range(N):
        kf = prm.kf if k == N - 1 else 1.0
        kuf = 0.0 if k == N - 1 else 1.0
        Sk = S[k + 1]
        Qk = prm.Qdiag * kf
        Hm += 2 * Sk.T @ (Qk[:, None] * Sk)
        g += 2 * Sk.T @ (Qk * (c[k + 1] - x_ref[k]))
        Hm[6 * k:6 * k + 6, 6 * k:6 * k + 6] += 2 * np.diag(prm.Rdiag * kuf)
        g[6 * k + 2] += -2 * prm.Rdiag[2] * kuf * ubar[k]
    m = 11 * N
    A = np.zeros((m, n))
    lo = np.full(m, -INF)
    hi = np.full(m, INF)
    A[:n, :n] = np.eye(n)
    mu = prm.mu
    infeasible = False
    for k in range(N):
        b = 6 * k
        lo[b + 3:b + 6] = -prm.tau_max
        hi[b + 3:b + 6] = prm.tau_max
        if prm.dyn == "2f":
            lo[b + 1] = hi[b + 1] = 0.0
        if C[k] == 0:
            lo[b:b + 3] = 0.0
            hi[b:b + 3] = 0.0
        else:
            lo[b + 2], hi[b + 2] = 0.0, prm.fz_max
            fr = n + 4 * k
            A[fr + 0, [b, b + 2]] = [1.0, -mu]; hi[fr + 0] = 0.0
            A[fr + 1, [b, b + 2]] = [-1.0, -mu]; hi[fr + 1] = 0.0
            if prm.dyn == "3f":
                A[fr + 2, [b + 1, b + 2]] = [1.0, -mu]; hi[fr + 2] = 0.0
                A[fr + 3, [b + 1, b + 2]] = [-1.0, -mu]; hi[fr + 3] = 0.0
        zr = n + 4 * N + k
        if k >= 2:
            A[zr] = S[k][2]
            lo[zr] = prm.z_min - c[k][2]
        else:
            if c[k][2] < prm.z_min:
                infeasible = True  # App. D2: u-independent height rows k=0,1
    # exact feasibility: the height rows are monotone in the stance fz (all coefficients >= 0) and
    # fz = fz_max, fx = fy = 0 satisfies every other row, so the QP is feasible iff z_k(fz_max) >= z_min
    zmax = np.zeros(N)
    for k in range(N):
        zr = n + 4 * N + k
        ub = np.where(hi[:n] > 1e20, 0.0, hi[:n])
        coef = S[k][2]
        zmax[k] = c[k][2] + np.sum(np.where(coef > 0, coef * ub, coef * np.where(lo[:n] < -1e20, 0.0, lo[:n])))
    height_infeasible = bool(np.any(zmax < prm.z_min))
    return dict(H=Hm, g=g, A=A, l=lo, u=hi, c=c, S=S, infeasible=infeasible or height_infeasible,
                infeasible_const=infeasible, zmax=zmax)


def rollout_linear(x_in, U, Ad, Bd, Gd, prm: Params):
    X = np.zeros((prm.N + 1, 12))
    X[0] = x_in
    for k in range(prm.N):
        X[k + 1] = Ad[k] @ X[k] + Bd[k] @ U[k] + Gd
    return X
