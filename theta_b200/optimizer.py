"""Optimizer (python/Optimizer.py:43-165) backed by the CUDA solve kernels.

Optimizer(r, rN, m, n, tau, lower_bound=0, upper_bound=1).solve(C) returns what the reference
returns: ((mu, 1-mu), likelihood, vals) for n=2, (mu ndarray, likelihood, vals) for n=3, or None.
"""
import numpy as np

from . import _lib

_KT = (2, 3, 4, 5, 6, 7, 8, 10, 12, 15)  # level-count templates of the solve kernels (theta_n2.cu::pick_KT)


class Optimizer(object):
    def __init__(self, r, rN, m, n, tau, lower_bound=0, upper_bound=1):
        self.r, self.rN, self.m, self.n, self.tau = r, rN, m, n, tau
        self.lB, self.uB = lower_bound, upper_bound  # normal-admixture bracket (Optimizer.py:43,109-112); n=2 only (Q15)
        self._plan = None    # n = 3
        self._plans = {}     # n = 2: one explicit-column plan per level-count template, made when first needed
        _lib.init()

    def _plan_for(self, kmax):
        kt = next((o for o in _KT if kmax <= o), None)
        if kt is None:
            raise ValueError("tumour copy numbers must be integers in [0, 15]")
        pl = self._plans.get(kt)
        if pl is None:
            # explicit-column plan: only sizes the kernels' tables, so it works for any m (the number of monotone
            # columns in [0, kt]^m is never counted -- THETA_PLAN_TILES_ONLY)
            pl = _lib.Plan(2, self.r, self.rN, [0] * self.m, [kt] * self.m, tau=self.tau, max_normal=float(self.uB),
                           min_normal=float(self.lB), tiles_only=True)
            self._plans[kt] = pl
        return pl

    def solve(self, C):
        C = np.asarray(C, dtype=np.float64)
        if self.n == 2:
            return self._solve_n2(C)
        from .search_n3 import optimizer_solve_n3
        return optimizer_solve_n3(self, C)

    def _solve_n2(self, C):
        if C.shape != (self.m, 2) or not np.all(C[:, 0] == self.tau):
            raise ValueError("C must be m x 2 with column 0 == tau")
        col = C[:, 1]
        if np.any(col != np.floor(col)) or col.min() < 0 or col.max() > 15:
            raise ValueError("tumour copy numbers must be integers in [0, 15]")
        mu, nll, st = self._plan_for(int(col.max())).solve_tiles(col.astype(np.int8)[None, :])
        if st[0] >= _lib.ST_NONE_SIGN:
            return None  # Optimizer.py:119-121
        mu0 = np.float64(mu[0])
        Cw = C * np.asarray(self.rN, dtype=np.float64)[:, None]  # weighted_C (Optimizer.py:176-182): input marshalling
        _, vals, _ = _lib.nll_single([mu0, 1 - mu0], Cw, self.r)
        return ((mu0, 1 - mu0), np.float64(nll[0]), [np.float64(v) for v in vals])
