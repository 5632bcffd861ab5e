"""Enumerator (python/Enumerator.py) backed by the on-chip enumeration kernels.

Same constructor and generate_next_C() contract: successive admissible interval count matrices
as m x n float64 arrays (column 0 = tau), then False.  Candidates are produced on the GPU in
blocks (theta_enumerate_n2 / theta_enumerate_n3) and handed out one at a time.
"""
import numpy as np

from . import _lib


class Enumerator(object):
    BLOCK = 1 << 14

    def __init__(self, n, m, k, tau, lower_bound=None, upper_bound=None, multi_event=False):
        self.m = m
        self.n = n - 1          # the reference stores n-1 (Enumerator.py:52)
        self.tau = tau
        if upper_bound is None:
            upper_bound = [k] * m
        if lower_bound is None:
            lower_bound = [0] * m
        self._n_pop = n
        dummy = [1] * m  # read counts are not needed to enumerate
        self._plan = _lib.Plan(n, dummy, dummy, lower_bound, upper_bound, tau=tau)
        lo, up = self._plan.bounds()
        # Enumerator._check_bound_order monotonises the caller's lists in place (Enumerator.py:90-113, Q16)
        for i in range(m):
            lower_bound[i] = int(lo[i])
            upper_bound[i] = int(up[i])
        self.lower_bound, self.upper_bound = lower_bound, upper_bound
        self.k = max(self.upper_bound)  # Enumerator.py:58 (Q4)
        self._total = self._plan.count() if n == 2 else None
        self._next = 0
        self._buf = None
        self._buf0 = 0
        self.iter = list(self.lower_bound) if n == 2 else [0] * m

    def _C_to_array(self):
        """Enumerator._C_to_array (Enumerator.py:154-160): the matrix of the current iterator state."""
        C = np.zeros((self.m, self.n + 1))
        C[:, 0] = self.tau
        if self._n_pop == 2:
            C[:, 1] = self.iter
        return C

    def generate_next_C(self):
        if self._n_pop != 2:
            from .search_n3 import enumerator_next_n3
            return enumerator_next_n3(self)
        if self._next >= self._total:
            return False
        if self._buf is None or self._next >= self._buf0 + len(self._buf):
            cnt = min(self.BLOCK, self._total - self._next)
            self._buf = self._plan.enumerate(self._next, cnt)
            self._buf0 = self._next
        col = self._buf[self._next - self._buf0]
        self._next += 1
        self.iter = [int(v) for v in col]
        return self._C_to_array()

    def count(self):
        """TimeEstimate.count_number_matrices_2 (TimeEstimate.py:91-111)."""
        return self._total
