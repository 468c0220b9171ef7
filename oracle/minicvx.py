"""Mini cvxpy stand-in.  *** TEST INFRASTRUCTURE ONLY *** (see hopper_oracle.py header)

cvxpy (unpinned third-party dependency, README.md:45) is not installed in this image.  This module
implements exactly the slice of its API that the reference touches
(mpc_cvx_euler_3f.py:38-39,96-160): ``Variable((r,c))`` with 2-D indexing, affine arithmetic with
numpy operands, ``quad_form``, ``==``/``<=``/``>=`` constraints, ``Problem(Minimize(cost),
constr).solve(solver=OSQP)`` and ``Variable.value``.  With it installed as ``sys.modules['cvxpy']``
(oracle/refshim.py) the reference's *own* ``build_qp``/``solve_qp``/``mpcontrol`` run unchanged and
emit the QP data (P, q, A, l, u) in cvxpy's OSQP layout [EXT, SURVEY App. C1]: equality rows first
(l=u=b), then one-sided inequality rows (l=-inf).

Constant semantics [EXT]: cvxpy wraps a float64 ndarray operand without copying and reads it only at
``solve()``.  This shim keeps references (not copies) to ndarray operands and evaluates them at solve
time, which reproduces the ``u_ref`` aliasing of mpc_cvx_euler_3f.py:107,131,138.
``Problem.solve(copy_constants=True)`` style behaviour is available via ``minicvx.ALIAS = False``.
"""
from __future__ import annotations

import numpy as np

from . import qp_solvers

OSQP = "OSQP"
ALIAS = True            # keep references to ndarray constants (cvxpy behaviour)
SOLVER_OPTS = {}        # extra keyword arguments for qp_solvers.osqp_solve (tests tighten eps here)
LAST = {}               # last canonicalised problem + solver result, for the tests


def _const(v):
    if isinstance(v, np.ndarray):
        return v if ALIAS else v.copy()
    return np.asarray(v, dtype=float)


class Expr:
    """Affine expression  M(v) @ v + c  over the stacked decision vector; M, c evaluated lazily."""
    __array_ufunc__ = None   # make numpy defer to our reflected operators (as cvxpy does)

    def __init__(self, shape, fn, vars_=()):
        self.shape = shape       # () or (k,)
        self._fn = fn            # fn(nv) -> (M (k,nv), c (k,))   (k=1 for scalars)
        self.vars = list(vars_)  # Variables referenced (identity-deduplicated at solve time)

    def eval(self, nv):
        return self._fn(nv)

    # --- arithmetic -----------------------------------------------------------------------
    def _lift(self, other):
        if isinstance(other, Expr):
            return other
        cref = _const(other)

        def fn(nv, cref=cref, k=int(np.prod(self.shape)) if self.shape else 1):
            c = np.array(cref, dtype=float).reshape(-1)
            if c.size == 1 and k > 1:
                c = np.full(k, c[0])
            return np.zeros((c.size, nv)), c
        return Expr(self.shape, fn)

    def __add__(self, other):
        o = self._lift(other)

        def fn(nv):
            M1, c1 = self.eval(nv); M2, c2 = o.eval(nv)
            return M1 + M2, c1 + c2
        return Expr(self.shape if self.shape else o.shape, fn, self.vars + o.vars)

    __radd__ = __add__

    def __neg__(self):
        def fn(nv):
            M, c = self.eval(nv)
            return -M, -c
        return Expr(self.shape, fn, self.vars)

    def __sub__(self, other):
        return self + (-self._lift(other))

    def __rsub__(self, other):
        return self._lift(other) + (-self)

    def __mul__(self, other):
        s = _const(other)

        def fn(nv):
            M, c = self.eval(nv)
            f = float(np.asarray(s).reshape(-1)[0])
            return M * f, c * f
        return Expr(self.shape, fn, self.vars)

    __rmul__ = __mul__

    def __rmatmul__(self, other):
        mat = _const(other)

        def fn(nv):
            M, c = self.eval(nv)
            Mm = np.asarray(mat, dtype=float)
            return Mm @ M, Mm @ c
        return Expr((np.asarray(other).shape[0],), fn, self.vars)

    # --- constraints ----------------------------------------------------------------------
    def __eq__(self, other):
        return Constraint(self - other, "eq")

    def __le__(self, other):
        return Constraint(self - other, "le")

    def __ge__(self, other):
        return Constraint(self._lift(other) - self, "le")

    __hash__ = None


class Constraint:
    def __init__(self, expr, kind):
        self.expr, self.kind = expr, kind   # expr == 0  or  expr <= 0


class Variable:
    _count = 0

    def __init__(self, shape):
        self.shape = tuple(shape)
        self.value = None
        self.size = int(np.prod(self.shape))
        Variable._count += 1
        self._serial = Variable._count

    def __getitem__(self, key):
        r, c = key
        idx = np.arange(self.size).reshape(self.shape)[r, c]
        idx = np.atleast_1d(idx).reshape(-1)
        shape = () if (np.ndim(np.arange(self.size).reshape(self.shape)[r, c]) == 0) else (idx.size,)

        def fn(nv, idx=idx, var=self):
            M = np.zeros((idx.size, nv))
            M[np.arange(idx.size), var._offset + idx] = 1.0
            return M, np.zeros(idx.size)
        return Expr(shape, fn, [self])


class QuadForm:
    def __init__(self, expr, P):
        self.terms = [(expr, _const(P))]

    def __add__(self, other):
        if isinstance(other, QuadForm):
            q = QuadForm.__new__(QuadForm)
            q.terms = self.terms + other.terms
            return q
        if isinstance(other, (int, float)) and other == 0:
            return self
        return NotImplemented

    __radd__ = __add__


def quad_form(expr, P):
    return QuadForm(expr, P)


class Minimize:
    def __init__(self, cost):
        self.cost = cost


class Problem:
    def __init__(self, objective, constraints):
        self.objective, self.constraints = objective, constraints
        self.status = None

    def canonicalize(self, variables):
        off = 0
        for v in variables:
            v._offset = off
            off += v.size
        nv = off
        P = np.zeros((nv, nv)); q = np.zeros(nv); const = 0.0
        for expr, Pm in self.objective.cost.terms:
            M, c = expr.eval(nv)
            Pm = np.asarray(Pm, dtype=float)
            P += 2.0 * M.T @ Pm @ M            # OSQP: 1/2 v'Pv ; quad_form has no 1/2
            q += 2.0 * M.T @ (Pm @ c)
            const += c @ Pm @ c
        Ae, be, Ai, bi = [], [], [], []
        for con in self.constraints:
            M, c = con.expr.eval(nv)
            (Ae if con.kind == "eq" else Ai).append(M)
            (be if con.kind == "eq" else bi).append(-c)
        Ae = np.vstack(Ae) if Ae else np.zeros((0, nv)); be = np.concatenate(be) if be else np.zeros(0)
        Ai = np.vstack(Ai) if Ai else np.zeros((0, nv)); bi = np.concatenate(bi) if bi else np.zeros(0)
        A = np.vstack((Ae, Ai))
        l = np.concatenate((be, np.full(bi.shape, -qp_solvers.INF)))
        u = np.concatenate((be, bi))
        return dict(P=P, q=q, A=A, l=l, u=u, const=const, n_eq=len(be), n_ineq=len(bi))

    def solve(self, solver=None, **kw):
        variables = _collect_variables(self)
        qp = self.canonicalize(variables)
        res = qp_solvers.osqp_solve(qp["P"], qp["q"], qp["A"], qp["l"], qp["u"], **SOLVER_OPTS)
        LAST.clear(); LAST.update(qp=qp, res=res)
        self.status = "optimal" if res["status"] == "solved" else res["status"]
        if res["status"] != "solved":
            for v in variables:
                v.value = None
            return None
        for v in variables:
            v.value = res["x"][v._offset:v._offset + v.size].reshape(v.shape).copy()
        return 0.5 * res["x"] @ qp["P"] @ res["x"] + qp["q"] @ res["x"] + qp["const"]


def _collect_variables(problem):
    """Variables in order of creation (cvxpy orders by variable id, i.e. creation order [EXT])."""
    seen, out = set(), []
    exprs = [t[0] for t in problem.objective.cost.terms] + [c.expr for c in problem.constraints]
    for e in exprs:
        for v in e.vars:
            if id(v) not in seen:
                seen.add(id(v)); out.append(v)
    out.sort(key=lambda v: v._serial)
    return out
