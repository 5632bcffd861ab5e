"""CPU: the oracle restatement (oracle/theta_oracle.py) against the golden vectors frozen from the
live reference (oracle/make_golden.py) and the reference's own L2/L3 pickles.  Bit-exact where the
reference's arithmetic order is reproduced (everything except CalcAllC's numpy dot products)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import theta_oracle as O  # noqa: E402


def test_calcallc_pickles(golden):
    k = golden("calcallc_kat.json")
    a = k["L2"]
    C = np.array(a["C"], dtype=float)
    nll, vals = O.calcallc_L2(a["mu"], C, a["m"], a["r"])
    assert float(nll) == a["nll_branch"] == 7565.513640605675
    assert abs(a["nll_master"] - a["nll_branch"]) <= 4e-16 * a["nll_branch"]
    assert [v if isinstance(v, str) else float(v) for v in vals] == a["vals_branch"]
    b = k["L3"]
    nll, vals = O.calcallc_L3(np.array(b["mu"]), np.array(b["C"], dtype=float), b["m"], b["r"], b["n"])
    assert float(nll) == b["nll_branch"] == 9961.895651323804
    assert [v if isinstance(v, str) else float(v) for v in vals] == b["vals_branch"]
    with pytest.raises(ValueError):
        O.calcallc_L2(0.3, np.zeros((3, 2)), 4, [1, 2, 3, 4])


def test_n2_small_spaces_bit_exact(golden):
    """Enumeration order, every candidate's (mu, NLL | None) and the tie fold, bit for bit."""
    for case in golden("n2_small_full.json"):
        cols = list(O.enumerate_n2(case["lower"], case["upper"]))
        assert ["".join(map(str, c)) for c in cols] == [s[0] for s in case["stream"]]
        assert len(cols) == O.count_n2(*O.check_bound_order(case["lower"], case["upper"]))
        for col, s in zip(cols, case["stream"]):
            got = O.solve_n2(O.column_to_matrix(col), case["r"], case["rN"], case["max_normal"])
            if s[1] is None:
                assert got is None
            else:
                assert float(got[0][0]) == s[1] and float(got[1]) == s[2]
        best, stats = O.search_single(2, case["lower"], case["upper"], case["r"], case["rN"], case["max_normal"],
                                      case["sorted_index"])
        assert stats["count"] == len(cols) + 1  # Q1: first matrix twice
        assert len(best) == len(case["best"])
        for b, g in zip(best, case["best"]):
            assert np.array_equal(b[0], np.array(g["C"]))
            assert float(b[1][0]) == g["mu"][0] and float(b[2]) == g["nll"]
            assert [float(v) for v in b[3]] == g["vals"]


def test_unrank_matches_enumeration(golden):
    for case in golden("n2_small_full.json"):
        for i, col in enumerate(O.enumerate_n2(case["lower"], case["upper"])):
            if i % 7 == 0:
                assert O.unrank_n2(case["lower"], case["upper"], i) == col


def test_synthetic_generator_and_samples(golden):
    data = golden("synth_n2_candidates.json")
    for name, case in data.items():
        P = O.synthetic_problem_n2(case["m"], case["k"], seed=case["seed"])
        assert P["r"] == case["r"] and P["rN"] == case["rN"]
        assert P["lower"] == case["lower"] and P["upper"] == case["upper"]
        assert O.count_n2(P["lower"], P["upper"]) == case["n_candidates"]
        for q in case["records"][::9]:
            col = O.unrank_n2(case["lower"], case["upper"], q["rank"])
            got = O.solve_n2(O.column_to_matrix(col), case["r"], case["rN"], case["max_normal"])
            if q["nll"] is None:
                assert got is None
            else:
                assert float(got[0][0]) == q["mu0"] and float(got[1]) == q["nll"]
    assert data["m150_k5"]["n_candidates"] == 25251156
    assert data["m150_k7"]["n_candidates"] == 607163886


def test_edge_cases(golden):
    e = golden("edge_n2.json")
    for name, c in e["cases"].items():
        got = O.solve_n2(O.column_to_matrix(c["col"]), c.get("r", e["r"]), e["rN"])
        if c["nll"] is None:
            assert got is None, name
        else:
            assert float(got[0][0]) == c["mu0"] and float(got[1]) == c["nll"], name
    assert e["cases"]["all_tau"]["mu0"] == 0.0          # flat likelihood -> brenth returns lB (Q17)
    assert e["cases"]["all_zero"]["nll"] is None        # zero tumour column -> NaN -> None
    assert e["cases"]["zero_r_c0"]["nll"] is None       # 0/0 at x = 0


def test_config1_summary(golden):
    g = golden("config1_search.json")
    inp = g["inputs"]
    assert inp["m"] == 61 and g["n_candidates"] == 142560 and len(g["none_ranks"]) == 290
    assert O.count_n2(inp["lower"], inp["upper"]) == 142560
    assert g["best"][0]["nll"] == 5497732728.26462
    assert g["best"][0]["mu"][0] == 0.535405337835399
    for q in g["top"][:5] + g["sample"][::200]:
        col = O.unrank_n2(inp["lower"], inp["upper"], q["rank"])
        got = O.solve_n2(O.column_to_matrix(col), inp["r"], inp["rN"])
        assert float(got[0][0]) == q["mu0"] and float(got[1]) == q["nll"]
    for rk in g["none_ranks"][::29]:
        col = O.unrank_n2(inp["lower"], inp["upper"], rk)
        assert O.solve_n2(O.column_to_matrix(col), inp["r"], inp["rN"]) is None


def test_fold_quirks():
    """Q5: appended lower values do not move the running minimum; NaN is close to everything (Q6)."""
    assert O.fold([("a", 10.0), ("b", 9.9995), ("c", 9.9991)]) == ["a", "b", "c"]
    assert O.fold([("a", 10.0), ("b", 9.9995), ("c", 9.9991), ("d", 9.9989)]) == ["d"]
    assert O.fold([("a", 10.0), ("b", 9.0), ("c", 9.0005), ("d", 8.9)]) == ["d"]
    assert O.fold([("a", 10.0), ("n", float("nan")), ("b", 10.0009)]) == ["a", "n", "b"]
    assert O.is_close(float("nan"), 1.0) and O.in_range([float("nan")])


def test_n3_small_space_bit_exact(golden):
    """n=3: enumeration order and a sample of candidates (incl. BFGS-fallback outcomes) bit for bit."""
    case = golden("n3_small_full.json.gz")[0]
    rows, _ = O.n3_graph(3, 2)
    assert [list(x) for x in rows] == case["rows"]
    paths = list(O.enumerate_n3(case["lower"], case["upper"]))
    keys = ["".join("%d%d" % rows[g] for g in p) for p in paths]
    assert keys == [s[0] for s in case["stream"][1:]]
    for p, s in list(zip(paths, case["stream"][1:]))[::41]:
        got = O.solve_n3(O.rows_to_matrix(p, rows), case["r"], case["rN"])
        if s[2] is None:
            assert got is None
        else:
            assert [float(v) for v in got[0]] == s[1] and float(got[1]) == s[2]


def test_n3_synthetic_configs(golden):
    data = golden("synth_n3_candidates.json")
    for name, case in data.items():
        P = O.synthetic_problem_n3(case["m"], case["k"], case["seed"], case["F"], case["tau_bounds"])
        assert P["r"] == case["r"] and P["lower"] == case["lower"] and P["upper"] == case["upper"]
        rows = [tuple(x) for x in case["rows"]]
        gen = O.enumerate_n3(case["lower"], case["upper"])
        for i, path in zip(range(3), gen):
            assert path == case["records"][i]["path"]
        q = case["records"][1]
        got = O.solve_n3(O.rows_to_matrix(q["path"], rows), case["r"], case["rN"])
        assert [float(v) for v in got[0]] == q["mu"] and float(got[1]) == q["nll"]
