"""Capture stand-in for `cvxpy` (test infrastructure only; used by oracle/make_golden.py).

Implements only what SCP_controller.py:135-146 touches, records the dense QP
(P, q, Aineq, bineq, lb, ub) the reference hands to its solver, and forwards it to a pluggable
solver `SOLVER(P, q, A, b, lb, ub) -> x` (the oracle's coneqp restatement), because neither Gurobi
(the shipped back-end) nor CVXOPT (the baseline back-end) can be installed here.
"""
import numpy as np

GUROBI = "GUROBI"
CVXOPT = "CVXOPT"
SOLVER = None          # set by the bridge
CAPTURE = []           # one dict per Problem.solve call


class Variable:
    __array_ufunc__ = None

    def __init__(self, shape):
        self.shape = tuple(shape)
        self.value = None

    def __rmatmul__(self, M):
        return _Affine(np.asarray(M, dtype=float), self)

    def __le__(self, rhs):
        return ("ub", np.asarray(rhs, dtype=float))

    def __ge__(self, rhs):
        return ("lb", np.asarray(rhs, dtype=float))


class _Affine:
    __array_ufunc__ = None

    def __init__(self, M, var):
        self.M, self.var = M, var

    def __le__(self, rhs):
        return ("A", self.M, np.asarray(rhs, dtype=float))


class _Quad:
    __array_ufunc__ = None

    def __init__(self, P, scale=1.0, lin=None):
        self.P, self.scale, self.lin = P, scale, lin

    def __rmul__(self, c):
        return _Quad(self.P, self.scale * float(c), self.lin)

    def __add__(self, other):
        assert isinstance(other, _Affine)
        return _Quad(self.P, self.scale, other.M)


def quad_form(x, P):
    return _Quad(np.asarray(P, dtype=float))


class Minimize:
    def __init__(self, expr):
        self.expr = expr


class Problem:
    def __init__(self, objective, constraints):
        self.objective, self.constraints = objective, constraints
        self.value = None

    def solve(self, solver=None, verbose=False):
        e = self.objective.expr
        P = 2.0 * e.scale * e.P                 # cost = scale * x'Px + q'x  ->  1/2 x'(2 scale P)x
        q = np.asarray(e.lin, dtype=float).reshape(-1)
        A = b = lb = ub = None
        var = None
        for c in self.constraints:
            if c[0] == "A":
                A, b = c[1], c[2].reshape(-1)
            elif c[0] == "ub":
                ub = c[1].reshape(-1)
            elif c[0] == "lb":
                lb = c[1].reshape(-1)
        x = SOLVER(P, q, A, b, lb, ub)
        CAPTURE.append(dict(P=P.copy(), q=q.copy(), A=A.copy(), b=b.copy(), lb=lb.copy(), ub=ub.copy(), x=x.copy()))
        for c in self.constraints:
            pass
        self._x = x
        # the Variable object is reachable through the affine objective term
        self.objective.expr_var = None
        Problem._last_var.value = x.reshape(-1, 1)
        self.value = float(0.5 * x @ P @ x + q @ x)
        return self.value


# The reference creates exactly one Variable per Problem, immediately before it (SCP_controller.py:135).
_orig_init = Variable.__init__


def _tracking_init(self, shape):
    _orig_init(self, shape)
    Problem._last_var = self


Variable.__init__ = _tracking_init
