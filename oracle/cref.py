"""ctypes binding of oracle/_ref/libhopper_ref.so (oracle/c/hopper_ref.cpp): the compiled C++ restatement of the
reference's per-tick recipe.  *** TEST / BASELINE INFRASTRUCTURE ONLY *** -- used by tests/ (cross-check of the numpy
oracle) and by bench.py's cpu_baseline / --impl reference legs; never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "c", "hopper_ref.cpp")
_SO = os.path.join(_HERE, "_ref", "libhopper_ref.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or (os.path.exists(_SRC) and os.path.getmtime(_SRC) > os.path.getmtime(_SO)):
        subprocess.run(["make", "-C", os.path.join(_HERE, "c")] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ref_closed_loop.restype = C.c_int
        _lib.ref_osqp_dense.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def osqp_dense(P, q, A, l, u, eps=1e-5, max_iter=10000, scaling=10, polish=True):
    """OSQP restatement on dense inputs.  Returns dict(x, y, status, iters, polished, n_fac, rho, pri, dua)."""
    P = np.ascontiguousarray(P, float); A = np.ascontiguousarray(A, float)
    n, m = P.shape[0], A.shape[0]
    q = np.ascontiguousarray(q, float)
    l = np.ascontiguousarray(np.maximum(l, -1e30), float); u = np.ascontiguousarray(np.minimum(u, 1e30), float)
    x, y = np.zeros(n), np.zeros(m)
    info = np.zeros(3, np.int32); dinfo = np.zeros(3)
    st = load().ref_osqp_dense(n, m, _p(P), _p(q), _p(A), _p(l), _p(u), C.c_double(eps), int(max_iter), int(scaling),
                               int(bool(polish)), _p(x), _p(y), _p(info), _p(dinfo))
    return dict(x=x, y=y, status={0: "solved", 1: "max_iter"}.get(st, "failed"), iters=int(info[0]), polished=bool(info[1]),
                n_fac=int(info[2]), rho=dinfo[0], pri=dinfo[1], dua=dinfo[2])


def build_qp(dyn, N, Qd, Rd, x_in, x_ref, x_guess, pf, Cvec):
    """The reference's full (cvxpy-shaped) QP assembled by the C++ restatement: dict(Pdiag, q, A, l, u)."""
    nv = 12 * (N + 1) + 6 * N
    mmax = 40 * N + 12
    Pd, q = np.zeros(nv), np.zeros(nv)
    A = np.zeros((mmax, nv)); l = np.zeros(mmax); u = np.zeros(mmax)
    m = C.c_int()
    cb = np.ascontiguousarray(np.asarray(Cvec) != 0, np.uint8)
    load().ref_build_qp(int(dyn), int(N), _p(np.ascontiguousarray(Qd, float)), _p(np.ascontiguousarray(Rd, float)),
                        _p(np.ascontiguousarray(x_in, float)), _p(np.ascontiguousarray(x_ref, float)),
                        _p(np.ascontiguousarray(x_guess, float)), _p(np.ascontiguousarray(pf, float)), _p(cb),
                        _p(Pd), _p(q), _p(A), _p(l), _p(u), C.byref(m))
    return dict(Pdiag=Pd, q=q, A=A[:m.value].copy(), l=l[:m.value].copy(), u=u[:m.value].copy())


def closed_loop(dyn, N, Qd, Rd, X0, xref_tab, pf_tab, Cmat, pf_switch, n_ticks, budget_s=0.0, eps=1e-5):
    """One hopper's closed loop with the reference's per-tick recipe (fresh full QP, OSQP cold start, polish, 20 RK4
    steps).  Returns dict(ticks, X_log, U_log, solve_us, iters, failed, inaccurate): inaccurate = solves that
    stopped at max_iter (the loop continues with OSQP's last iterate, as cvxpy does)."""
    X = np.ascontiguousarray(X0, float).copy()
    Xl = np.zeros((n_ticks + 1, 13)); Ul = np.zeros((n_ticks, 6))
    us = np.zeros(n_ticks); it = np.zeros(n_ticks, np.int32)
    ninacc = C.c_int(0)
    cb = np.ascontiguousarray(np.asarray(Cmat)[:n_ticks] != 0, np.uint8)
    sw = np.ascontiguousarray(pf_switch[:n_ticks], np.uint8)
    r = load().ref_closed_loop(int(dyn), int(N), _p(np.ascontiguousarray(Qd, float)), _p(np.ascontiguousarray(Rd, float)),
                               _p(X), _p(np.ascontiguousarray(xref_tab, float)), _p(np.ascontiguousarray(pf_tab, float)),
                               _p(cb), _p(sw), int(n_ticks), C.c_double(budget_s), C.c_double(eps), _p(Xl), _p(Ul), _p(us), _p(it),
                               C.byref(ninacc))
    ticks = r if r >= 0 else -1 - r
    return dict(ticks=ticks, X_log=Xl[:ticks + 1], U_log=Ul[:ticks], solve_us=us[:ticks], iters=it[:ticks], failed=r < 0,
                inaccurate=ninacc.value)


def rk4(X, U, pf, nsteps):
    X = np.ascontiguousarray(X, float).copy()
    x = np.zeros(12)
    load().ref_rk4(_p(X), _p(np.ascontiguousarray(U, float)), _p(np.ascontiguousarray(pf, float)), int(nsteps), _p(x))
    return X, x
