"""Minimal stand-in for `optimobo.problem` (pymoo-derived; pymoo is not a dependency here).

User problems subclass `Problem` (vectorised: `_evaluate(X (m, n_var), out)`) or
`ElementwiseProblem` (`_evaluate(x (n_var,), out)`) exactly as in the reference README, and
optionally `_evaluate_constraints` writing `out["G"]` (problem.py:405, added by the reference's
author).  `evaluate(x)` returns F with the leading axis dropped for a single input
(problem.py:250-300).  Objective evaluation is the expensive black box: it stays on the host.
"""
from __future__ import annotations

import numpy as np


class Problem:
    elementwise = False

    def __init__(self, n_var=-1, n_obj=1, n_ieq_constr=0, n_eq_constr=0, xl=None, xu=None, vtype=None,
                 elementwise=None, **kwargs):
        self.n_var, self.n_obj = n_var, n_obj
        self.n_ieq_constr, self.n_eq_constr = n_ieq_constr, n_eq_constr
        self.vtype = vtype
        if elementwise is not None:
            self.elementwise = elementwise
        self.xl = None if xl is None else np.broadcast_to(np.asarray(xl, dtype=float), (n_var,)).copy()
        self.xu = None if xu is None else np.broadcast_to(np.asarray(xu, dtype=float), (n_var,)).copy()

    # ---- user hooks -------------------------------------------------------------------------
    def _evaluate(self, x, out, *args, **kwargs):          # pragma: no cover - abstract
        raise NotImplementedError

    def _evaluate_constraints(self, x, out, *args, **kwargs):   # pragma: no cover - abstract
        raise NotImplementedError

    # ---- driver -------------------------------------------------------------------------------
    def _run(self, hook, key, X, width):
        X = np.asarray(X, dtype=float)
        single = X.ndim == 1
        X2 = np.atleast_2d(X)
        assert X2.shape[1] == self.n_var, f"Input dimension {X2.shape[1]} are not equal to n_var {self.n_var}!"
        if self.elementwise:
            rows = []
            for x in X2:
                out = {}
                hook(x, out)
                rows.append(np.asarray(out[key], dtype=float).reshape(-1))
            V = np.vstack(rows)
        else:
            out = {}
            hook(X2, out)
            v = out[key]
            V = np.column_stack([np.asarray(c, dtype=float).reshape(-1) for c in v]) if isinstance(v, (list, tuple)) \
                else np.asarray(v, dtype=float).reshape(len(X2), -1)
        assert V.shape[1] == width, f"{key} has {V.shape[1]} columns, expected {width}"
        return V[0] if single else V

    def evaluate(self, X, *args, **kwargs):
        return self._run(self._evaluate, "F", X, self.n_obj)

    def evaluate_constraints(self, X, *args, **kwargs):
        return self._run(self._evaluate_constraints, "G", X, self.n_ieq_constr + self.n_eq_constr)

    def bounds(self):
        return self.xl, self.xu


class ElementwiseProblem(Problem):
    elementwise = True
