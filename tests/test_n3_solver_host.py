"""CPU: the n=3 device solver code compiled for the host and checked against SciPy / libm and EVERY golden candidate of
the live reference -- no exemptions: rank-deficient candidates (rows on one line, zero tumour columns)
are reproduced by the reference-order solver (theta_b200/csrc/n3_ref.cuh), the rest by the grouped-statistics solver.

These harnesses are test infrastructure: they compile the *product's* headers (hybrj4.cuh, bfgs_ls.cuh, theta_n3.cuh,
n3_ref.cuh, glibc_pow2.cuh) into small host libraries so the arithmetic that runs inside the CUDA kernels is also
exercised where no GPU exists.  The product never runs them.
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
GXX = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-fno-builtin", "-Wno-unknown-pragmas", "-shared", "-fPIC"]


def collinear(key):
    """Rank-deficient candidate: its distinct rows (a, b) lie on one line, so mu is not identifiable.
    Used only to label statistics -- no test skips these any more."""
    pts = sorted(set((int(key[i]), int(key[i + 1])) for i in range(0, len(key), 2)))
    if len(pts) <= 2:
        return True
    (x0, y0), (x1, y1) = pts[0], pts[1]
    return all((x1 - x0) * (y - y0) - (y1 - y0) * (x - x0) == 0 for x, y in pts[2:])


def _build(tmp_path_factory, src, name):
    out = str(tmp_path_factory.mktemp("h") / name)
    subprocess.check_call(GXX + ["-o", out, os.path.join(HARN, src)])
    return C.CDLL(out)


@pytest.fixture(scope="module")
def hybrj_lib(tmp_path_factory):
    # -fno-builtin: g++ would fold pow(d, 2.0) into d*d, which is NOT what libm returns (see glibc_pow2.cuh)
    return _build(tmp_path_factory, "hybrj_harness.cpp", "hybrj_harness.so")


@pytest.fixture(scope="module")
def ref_lib(tmp_path_factory):
    return _build(tmp_path_factory, "n3_ref_harness.cpp", "n3_ref_harness.so")


@pytest.fixture(scope="module")
def n3_lib(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("n") / "n3_host.so")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-mfma,-ffp-contract=off", "-shared",
                           "-o", out, os.path.join(HARN, "n3_host_harness.cu")], stderr=subprocess.DEVNULL)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ab(key):
    return (np.array([int(ch) for ch in key[0::2]], dtype=np.int32), np.array([int(ch) for ch in key[1::2]], dtype=np.int32))


def test_pow2_is_libm_pow(ref_lib):
    """glibc_pow2.cuh == libm pow(x, 2.0) bit for bit (x*x is not: ~0.08 % of inputs differ by an ulp)."""
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.random(300000), np.exp(rng.uniform(-200, 200, 200000)), -np.exp(rng.uniform(-30, 30, 100000)),
                         1 + rng.uniform(-1e-3, 1e-3, 100000), [0.0, -0.0, 1.0, np.inf, -np.inf, 1e-200, 1e120, 5e-324]])
    ys = np.zeros_like(xs)
    ref_lib.pow2_batch(_p(xs), _p(ys), C.c_long(len(xs)))
    ref = np.array([math.pow(v, 2.0) for v in xs])
    assert np.array_equal(ys, ref)
    assert (xs * xs != ref).sum() > 100  # the distinction is real
    nan = np.array([np.nan])
    out = np.zeros(1)
    ref_lib.pow2_batch(_p(nan), _p(out), C.c_long(1))
    assert out[0] != out[0]


def test_hybrj_port_matches_scipy_fsolve(hybrj_lib, golden):
    """Same function, same arithmetic order -> the MINPACK restatement reproduces SciPy's iterates bit for bit,
    including on the singular systems of rank-deficient candidates."""
    import theta_oracle as O
    from scipy import optimize
    import warnings
    warnings.simplefilter("ignore")
    for case, step in zip(golden("n3_small_full.json.gz"), (5, 11, 31)):
        r = np.array(case["r"], float)
        tot = bit = ncol = 0
        for n, s in enumerate(case["stream"][1:]):
            col = collinear(s[0])
            if not col and n % step:
                continue
            a, b = _ab(s[0])
            if not a.any() or not b.any():
                continue  # zero column: C_hat is NaN, fsolve never moves (covered by the end-to-end test below)
            Cm = np.zeros((case["m"], 3))
            Cm[:, 0] = 2
            Cm[:, 1], Cm[:, 2] = a, b
            Ch = np.ascontiguousarray(O.normalize_C(O.weighted_C(Cm, case["rN"])))
            eq, jac = O._kkt_fun_jac(r, Ch)
            ref = optimize.fsolve(eq, [1 / 3] * 3 + [1], fprime=jac)
            x = np.array([1 / 3, 1 / 3, 1 / 3, 1.0])
            nfev, njev = C.c_int(), C.c_int()
            hybrj_lib.hybrj_ref_system(case["m"], _p(r), _p(Ch), _p(x), C.byref(nfev), C.byref(njev))
            tot += 1
            ncol += col
            bit += int((x == ref).all())
        assert tot > 300 and ncol > 200
        assert bit == tot, (bit, tot)


def _run_all(lib_fn, case, mode=None):
    r = np.array(case["r"], float)
    rN = np.array(case["rN"], float)
    m = case["m"]
    cat = collections.Counter()
    for s in case["stream"][1:]:
        key = s[0]
        a, b = _ab(key)
        mu, mh = np.zeros(3), np.zeros(3)
        nll, nfev = C.c_double(), C.c_int()
        args = [m, 2, _p(r), _p(rN), _p(a), _p(b), _p(mu), _p(mh), C.byref(nll), C.byref(nfev)]
        if mode is not None:
            args.append(mode)
        st = lib_fn(*args)
        tag = "col" if collinear(key) else "reg"
        if s[2] is None:
            assert st >= 8, (key, st)
            cat[tag + "_none"] += 1
            continue
        assert st < 8, (key, st)
        if s[2] != s[2]:  # the reference kept a NaN likelihood (Q6): so must we
            assert nll.value != nll.value, (key, nll.value)
            cat[tag + "_nan"] += 1
            continue
        assert abs(nll.value - s[2]) <= 1e-10 * abs(s[2]), (key, nll.value, s[2])
        assert np.abs(mu - np.array(s[1])).max() <= 1e-9, (key, mu, s[1])
        cat[tag + "_st%d" % st] += 1
        cat["bit"] += int(nll.value == s[2] and (mu == np.array(s[1])).all())
        cat["solved"] += 1
    return cat


def test_reference_order_solver_reproduces_every_golden_candidate(ref_lib, golden):
    """ref3::solve (weighted_C/normalize_C, hybrj on per-interval sums with libm-exact pow, inRange, BFGS fallback,
    M3 through hybrd, L3) vs all 29 881 candidates: identical None set, NaN set, and (mu, NLL) -- bit for bit on all
    but the handful where numpy's SIMD log and libm's log differ in the last place."""
    tot = collections.Counter()
    for case in golden("n3_small_full.json.gz"):
        cat = _run_all(ref_lib.n3_ref_solve_host, case)
        tot.update(cat)
        assert cat["bit"] >= 0.999 * cat["solved"], dict(cat)
        if case["m"] == 6:  # the space with zero tumour columns: the reference's garbage-mu and NaN-NLL outcomes
            assert cat["col_nan"] == 15 and cat["col_none"] == 26 and cat["reg_none"] == 590, dict(cat)
    assert tot["solved"] + tot["col_none"] + tot["reg_none"] + tot["col_nan"] + tot["reg_nan"] == 29881, dict(tot)


def test_n3_device_solver_on_host_matches_reference(n3_lib, golden):
    """n3_solve_auto as the search kernels call it (mode 0: grouped statistics + hybrj + fallback classification,
    reference-order solver for rank-deficient candidates) vs every candidate of the golden streams: identical None set,
    mu within 1e-9, NLL within 1e-10 relative.  No candidate is exempt."""
    n_col = 0
    for case in golden("n3_small_full.json.gz"):
        cat = _run_all(n3_lib.n3_solve_host, case, mode=0)
        assert cat["reg_st0"] > 1000 and cat["reg_st3"] > 500, dict(cat)
        n_col += sum(v for k, v in cat.items() if k.startswith("col_"))
        if case["m"] == 6:
            assert cat["reg_none"] == 590  # every BFGS-fallback candidate the reference drops (H2 outcome b)
    assert n_col == 227 + 378 + 873


def _check_rankdef(solve, g):
    """solve(path) -> (status, mu[3], nll); g = the n3_rankdef_m16 fixture."""
    cat = collections.Counter()
    for path, mu_ref, nll_ref in g["records"]:
        st, mu, nll = solve(path)
        if nll_ref is None:
            assert st >= 8, (path, st)
            cat["none"] += 1
        elif nll_ref != nll_ref:
            assert st < 8 and nll != nll, (path, st, nll)
            cat["nan"] += 1
        else:
            assert st < 8, (path, st)
            assert abs(nll - nll_ref) <= 1e-10 * abs(nll_ref), (path, nll, nll_ref)
            assert np.abs(np.asarray(mu) - np.array(mu_ref)).max() <= 1e-9, (path, mu, mu_ref)
            cat["solved"] += 1
    return cat


def test_rank_deficient_candidates_of_a_real_shaped_space(n3_lib, golden):
    """14 380 rank-deficient candidates (collinear rows, identical columns, zero columns) of a real-data m=16 space,
    solved by the live reference (oracle/make_golden.py::gen_n3_rankdef): the kernel's solver dispatch (mode 0) must
    send every one of them to the reference-order solver and reproduce the None set, the NaN-likelihood set (Q6) and
    (mu, NLL) within the parity bar."""
    g = golden("n3_rankdef_m16.json.gz")
    r = np.array(g["r"], float)
    rN = np.array(g["rN"], float)
    rows = g["rows"]

    def solve(path):
        a = np.array([rows[q][0] for q in path], dtype=np.int32)
        b = np.array([rows[q][1] for q in path], dtype=np.int32)
        mu, mh = np.zeros(3), np.zeros(3)
        nll, nfev = C.c_double(), C.c_int()
        st = n3_lib.n3_solve_host(g["m"], 2, _p(r), _p(rN), _p(a), _p(b), _p(mu), _p(mh), C.byref(nll), C.byref(nfev), 0)
        return st, mu, nll.value

    cat = _check_rankdef(solve, g)
    assert cat["solved"] + cat["none"] + cat["nan"] == len(g["records"]) >= 10000
    assert cat["none"] == 1169 and cat["nan"] == 48, dict(cat)
