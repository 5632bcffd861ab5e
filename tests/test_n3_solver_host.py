"""CPU: the n=3 device solver code (hybrj restatement, BFGS-fallback classification, grouped KKT system)
compiled for the host and checked against SciPy and the golden vectors of the live reference.

These harnesses are test infrastructure: they compile the *product's* headers
(theta_b200/csrc/hybrj4.cuh, bfgs_ls.cuh, theta_n3.cuh) into small host libraries so the arithmetic
that runs inside the CUDA kernels is also exercised where no GPU exists.  The product never runs them.
"""
import collections
import ctypes as C
import math
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
HARN = os.path.join(ROOT, "tests", "harness")


def collinear(key):
    """Rank-deficient candidate: its distinct rows (a, b) lie on one line, so mu is not identifiable."""
    pts = sorted(set((int(key[i]), int(key[i + 1])) for i in range(0, len(key), 2)))
    if len(pts) <= 2:
        return True
    (x0, y0), (x1, y1) = pts[0], pts[1]
    return all((x1 - x0) * (y - y0) - (y1 - y0) * (x - x0) == 0 for x, y in pts[2:])


@pytest.fixture(scope="module")
def hybrj_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("h") / "hybrj_harness.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out,
                           os.path.join(HARN, "hybrj_harness.cpp")])
    return C.CDLL(out)


@pytest.fixture(scope="module")
def n3_lib(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("n") / "n3_host.so")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
                           "-o", out, os.path.join(HARN, "n3_host_harness.cu")], stderr=subprocess.DEVNULL)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_hybrj_port_matches_scipy_fsolve(hybrj_lib, golden):
    """Same function, same arithmetic order -> the MINPACK restatement must reproduce SciPy's iterates."""
    import theta_oracle as O
    from scipy import optimize
    import warnings
    warnings.simplefilter("ignore")
    case = golden("n3_small_full.json.gz")[0]
    r = np.array(case["r"], float)
    tot = bit = close = 0
    for s in case["stream"][1::5]:
        if collinear(s[0]):
            continue
        Cm = np.zeros((case["m"], 3))
        Cm[:, 0] = 2
        Cm[:, 1] = [int(ch) for ch in s[0][0::2]]
        Cm[:, 2] = [int(ch) for ch in s[0][1::2]]
        Ch = np.ascontiguousarray(O.normalize_C(O.weighted_C(Cm, case["rN"])))
        eq, jac = O._kkt_fun_jac(r, Ch)
        ref = optimize.fsolve(eq, [1 / 3] * 3 + [1], fprime=jac)
        x = np.array([1 / 3, 1 / 3, 1 / 3, 1.0])
        nfev, njev = C.c_int(), C.c_int()
        hybrj_lib.hybrj_ref_system(case["m"], _p(r), _p(Ch), _p(x), C.byref(nfev), C.byref(njev))
        tot += 1
        bit += int((x == ref).all())
        close += int(np.abs(x[:3] - ref[:3]).max() < 1e-12)
    assert tot > 300
    assert close == tot, (close, tot)
    assert bit >= 0.98 * tot, (bit, tot)


def test_n3_device_solver_on_host_matches_reference(n3_lib, golden):
    """n3_solve (grouped statistics + hybrj + fallback classification) vs every regular candidate of the golden
    streams: identical None set, mu within 1e-9, NLL within 1e-10 relative."""
    for case in golden("n3_small_full.json.gz"):
        r = np.array(case["r"], float)
        rN = np.array(case["rN"], float)
        m = case["m"]
        cat = collections.Counter()
        for s in case["stream"][1:]:
            key = s[0]
            if collinear(key):
                continue
            a = np.array([int(ch) for ch in key[0::2]], dtype=np.int32)
            b = np.array([int(ch) for ch in key[1::2]], dtype=np.int32)
            mu, mh = np.zeros(3), np.zeros(3)
            nll, nfev = C.c_double(), C.c_int()
            st = n3_lib.n3_solve_host(m, 2, _p(r), _p(rN), _p(a), _p(b), _p(mu), _p(mh), C.byref(nll), C.byref(nfev))
            if s[2] is None:
                assert st >= 8, (key, st)
                cat["none"] += 1
                continue
            assert st < 8, (key, st)
            assert abs(nll.value - s[2]) <= 1e-10 * abs(s[2]), (key, nll.value, s[2])
            assert np.abs(mu - np.array(s[1])).max() <= 1e-9, (key, mu, s[1])
            cat["st%d" % st] += 1
        assert cat["st0"] > 1000 and cat["st3"] > 500, dict(cat)
        if m == 6:
            assert cat["none"] == 590  # every BFGS-fallback candidate the reference drops (H2 outcome b)
