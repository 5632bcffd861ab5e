"""CalcAllC.L2 / L3 (python/CalcAllC.py:44-76) on the GPU: the masked NLL of one weighted
interval-count matrix at a given mu, as used by calc_all_c_* and un_meta_cluster_results_*.

Same signatures, return values ((np.float64, list with 'X' for masked rows)), the
ValueError on shape mismatch, and L2's in-place scaling of the caller's C (Q10).
"""
import numpy as np

from . import _lib


def L2(mu, C, m, r):
    if m != C.shape[0]:
        raise ValueError('m not equal to first dimension of C')
    nll, vals, valid = _lib.nll_single([mu, 1 - mu], C[:, :2], r)
    # the reference rescales the caller's matrix (CalcAllC.py:54-55); callers rely on it (Q10)
    C[:, 0] = C[:, 0] * mu
    C[:, 1] = C[:, 1] * (1 - mu)
    return (np.float64(nll), [np.float64(v) if ok else 'X' for v, ok in zip(vals, valid)])


def L3(mu, C, m, r, n):
    if m != C.shape[0]:
        raise ValueError('m not equal to first dimension of C')
    if n != C.shape[1]:
        raise ValueError('n not equal to second dimension of C')
    nll, vals, valid = _lib.nll_single(np.asarray(mu, dtype=np.float64), C, r)
    return (np.float64(nll), [np.float64(v) if ok else 'X' for v, ok in zip(vals, valid)])
