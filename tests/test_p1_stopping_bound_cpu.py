"""CPU check of the mathematics behind pass 1's stopping test (theta_core.cuh: n2_halley_step<P1>, n2_root_p1).

Claim: for f(x) = -sum_v R_v e_v / den_v(x), den_v = x S1 + v (1-x) A, e_v = S1 - v A, with a sign change on (0, 1), the
error left after ONE Halley step from a point at distance dN (the Newton step) of the root is at most 2 W^2 |dN|^3,
W = 1/min(x, 1-x); hence  dN^2 <= 1e-9 x min(x,1-x)  and  |dN| <= 1e-2 min(x,1-x)  imply an error <= 2e-11 |x| -- the
accuracy contract of the generic loop (THETA_NEWTON_PRED).  Checked here in extended precision on random grouped
statistics: the root by bisection in longdouble, the Halley step in float64 exactly as the kernel takes it."""
import numpy as np

LD = np.longdouble


def _f_terms(R, S1, A, x, dtype):
    v = np.arange(len(R), dtype=dtype)
    den = x * S1 + v * (1 - x) * A
    e = S1 - v * A
    w = e / den
    return -(R * w).sum(), (R * w * w).sum(), -(R * w ** 3).sum()  # f, f', f''/2


def _root_ld(R, S1, A):
    lo, hi = LD(0), LD(1)
    for _ in range(200):
        mid = (lo + hi) / 2
        f, _, _ = _f_terms(R.astype(LD), LD(S1), LD(A), mid, LD)
        if f < 0:
            lo = mid
        else:
            hi = mid
    return (lo + hi) / 2


def test_halley_step_error_bound_and_p1_test():
    rng = np.random.default_rng(7)
    checked = passed = 0
    worst = 0.0
    for _ in range(400):
        K = int(rng.integers(2, 9))
        m = int(rng.integers(20, 300))
        rN = rng.integers(10_000, 5_000_000, m).astype(np.float64)
        c = np.sort(rng.integers(0, K + 1, m))
        if c[0] != 0:
            c[0] = 0  # reads at copy number 0: f(0) = -inf (the common case of the search)
        mu = rng.uniform(0.03, 0.97)
        p = rN * (2 * mu + c * (1 - mu))
        r = rng.multinomial(int(rN.sum()), p / p.sum()).astype(np.float64)
        R = np.array([r[c == v].sum() for v in range(K + 1)])
        A, S1 = rN.sum(), (rN * c).sum()
        if not (S1 > 0):
            continue
        f1 = -(R * (S1 - np.arange(K + 1) * A)).sum()  # sign of f(1)
        if not (R[0] > 0 and f1 > 0):
            continue
        xs = _root_ld(R, S1, A)
        if not (1e-6 < xs < 1 - 1e-6):
            continue
        for rel in (3e-7, 1e-6, 1e-5, 1e-4, -1e-6, -3e-5, 1e-3, -1e-2):
            x = float(xs) * (1.0 + rel)
            if not (0.0 < x < 1.0):
                continue
            f, fp, fpp2 = _f_terms(R, S1, A, x, np.float64)
            dN = -f / fp
            c2 = fpp2 / fp
            xn = x + dN / (1.0 + dN * c2)  # the Halley step of n2_root_p1
            mx = min(x, 1.0 - x)
            err = abs(float(LD(xn) - xs))
            W = 1.0 / mx
            checked += 1
            # the bound itself (leading order: allow 5 % for the higher-order terms in the asymptotic regime)
            if abs(dN) <= 1e-2 * mx:
                assert err <= 1.05 * 2.0 * W * W * abs(dN) ** 3 + 1e-14 * x, (rel, err, 2 * W * W * abs(dN) ** 3)  # + fp64 evaluation noise
            # the P1 test implies the accuracy contract
            if dN * dN <= 1e-9 * x * mx and abs(dN) <= 1e-2 * mx:
                passed += 1
                assert err <= 2e-11 * x, (rel, err / x)
                worst = max(worst, err / x)
    assert checked > 1500 and passed > 700
    # the fp32-refined start of the search (error ~1e-6 relative) passes the test with orders of magnitude to spare
    assert worst < 2e-11
