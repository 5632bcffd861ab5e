"""GPU: the RunTHetA-compatible command line on BASELINE config #1 against the files the reference's own CLI wrote
(oracle/make_golden.py::gen_config1_cli): same .withBounds text, same .results line (C string, mu within 1e-9,
NLL within 1e-10 relative, p* within 1e-9)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _parse_results(path):
    with open(path) as f:
        lines = [ln.rstrip("\n").split("\t") for ln in f if not ln.startswith("#")]
    return [(float(a), [float(x) for x in b.split(",")], c, [v if v == "X" else float(v) for v in d.split(",")])
            for a, b, c, d in lines]


def test_cli_config1_matches_reference_outputs(tmp_path):
    from theta_b200 import run_theta
    out = str(tmp_path)
    run_theta.main([os.path.join(GOLD, "Example.intervals"), "-n", "2", "-k", "3", "--NUM_PROCESSES", "1", "-d", out, "-p", "ex"])
    with open(os.path.join(out, "ex.n2.withBounds")) as f, open(os.path.join(GOLD, "config1_cli.n2.withBounds")) as g:
        assert f.read() == g.read()
    got = _parse_results(os.path.join(out, "ex.n2.results"))
    ref = _parse_results(os.path.join(GOLD, "config1_cli.n2.results"))
    assert len(got) == len(ref) == 1
    (nll, mu, C, p), (rnll, rmu, rC, rp) = got[0], ref[0]
    assert C == rC
    assert abs(nll - rnll) <= 1e-10 * rnll and rnll == 5793657923.019962
    assert np.abs(np.array(mu) - np.array(rmu)).max() <= 1e-9
    for a, b in zip(p, rp):
        assert (a == "X") == (b == "X")
        if b != "X":
            assert abs(a - b) <= 1e-9 * abs(b)
    assert os.path.exists(os.path.join(out, "ex.RunN3.bash"))


def _same_results(path, ref_path, n):
    got, ref = _parse_results(path), _parse_results(ref_path)
    assert len(got) == len(ref)
    for (nll, mu, C, p), (rnll, rmu, rC, rp) in zip(got, ref):
        assert C == rC
        assert abs(nll - rnll) <= 1e-10 * rnll
        assert len(mu) == n and np.abs(np.array(mu) - np.array(rmu)).max() <= 1e-9
        for a, b in zip(p, rp):
            assert (a == "X") == (b == "X")
            if b != "X":
                assert abs(a - b) <= 1e-9 * abs(b)


def _stage_n2(tmp_path):
    """The reference CLI's n=2 outputs of config #1 under the names RunN3.bash would use."""
    import shutil
    out = str(tmp_path)
    shutil.copy(os.path.join(GOLD, "config1_cli.n2.results"), os.path.join(out, "ex.n2.results"))
    shutil.copy(os.path.join(GOLD, "config1_cli.n2.withBounds"), os.path.join(out, "ex.n2.withBounds"))
    return out


@pytest.mark.parametrize("flag,gold", [([], "cli_n3"), (["--NO_MULTI_EVENT"], "cli_n3_nme")])
def test_cli_n3_from_n2_files_matches_reference_outputs(tmp_path, flag, gold):
    """-n 3 --RESULTS <n=2 results> on the n=2 .withBounds file: interval selection from the n=2 copy numbers
    (select_intervals_n3), the n=3 search, and the extension to the unselected intervals by
    CalcAllC.calc_all_c_3_multi_event (default) / calc_all_c_3 (--NO_MULTI_EVENT) -- against the files the
    reference's own CLI wrote with the same flags (oracle/make_golden.py::gen_cli_default_flow; --NUM_INTERVALS 8
    keeps the reference's n=3 run to minutes)."""
    from theta_b200 import run_theta
    out = _stage_n2(tmp_path)
    run_theta.main([os.path.join(out, "ex.n2.withBounds"), "-n", "3", "-k", "3", "--NUM_INTERVALS", "8", "--NUM_PROCESSES", "1",
                    "--RESULTS", os.path.join(out, "ex.n2.results"), "-d", out, "-p", "t"] + flag)
    with open(os.path.join(out, "t.n3.withBounds")) as f, open(os.path.join(GOLD, gold + ".withBounds")) as g:
        assert f.read() == g.read()
    _same_results(os.path.join(out, "t.n3.results"), os.path.join(GOLD, gold + ".results"), 3)


def test_cli_default_flow_matches_reference_outputs(tmp_path):
    """No -n: n=2, then n=3 seeded by the n=2 files, then model selection, in one call -- against the reference's own
    run with the same flags (--NUM_INTERVALS 8 --FORCE)."""
    if not os.path.exists(os.path.join(GOLD, "cli_flow.n3.results")):
        pytest.skip("default-flow fixture not generated")
    from theta_b200 import run_theta
    out = str(tmp_path)
    run_theta.main([os.path.join(GOLD, "Example.intervals"), "-k", "3", "--NUM_INTERVALS", "8", "--FORCE", "--NUM_PROCESSES", "1",
                    "-d", out, "-p", "flow"])
    for n in (2, 3):
        with open(os.path.join(out, "flow.n%d.withBounds" % n)) as f, open(os.path.join(GOLD, "cli_flow.n%d.withBounds" % n)) as g:
            assert f.read() == g.read()
        _same_results(os.path.join(out, "flow.n%d.results" % n), os.path.join(GOLD, "cli_flow.n%d.results" % n), n)
    with open(os.path.join(GOLD, "cli_flow.BEST.results")) as g:
        ref_best = g.read()
    ref_n = 3 if ref_best == open(os.path.join(GOLD, "cli_flow.n3.results")).read() else 2
    with open(os.path.join(out, "flow.BEST.results")) as f, open(os.path.join(out, "flow.n%d.results" % ref_n)) as g:
        assert f.read() == g.read()


def _dump_lines(path):
    with open(path) as f:
        return [ln.rstrip("\n").split("\t") for ln in f]


def test_get_values_dump_matches_reference(tmp_path, golden):
    """--GET_VALUES (RunTHetA.py:210-215): the .likelihoods dump, line for line, against the reference's own dumps of
    the whole small n=2 spaces (first matrix twice: Q1) and against the live reference's per-candidate results for the
    n=3 spaces (all-[tau,0,0] matrix first: Q2): same column strings in the same order, mu within 1e-9, NLL within 1e-10."""
    from theta_b200 import _lib
    from theta_b200.search import dump_values_n2
    from theta_b200.search_n3 import dump_values_n3
    _lib.init(0)
    for case in golden("n2_small_full.json"):
        plan = _lib.Plan(2, case["r"], case["rN"], case["lower"], case["upper"], max_normal=case["max_normal"])
        out = str(tmp_path / "n2.likelihoods")
        dump_values_n2(plan, out)
        want = [s for s in case["stream"] if s[1] is not None]
        if case["stream"][0][1] is not None:
            want = [want[0]] + want
        got = _dump_lines(out)
        assert [g[0] for g in got] == [w[0] for w in want]
        for g, w in zip(got, want):
            assert abs(float(g[1]) - w[1]) <= 1e-9 and abs(float(g[2]) - w[2]) <= 1e-10 * abs(w[2]), (g, w)
        plan.close()
    for case in golden("n3_small_full.json.gz"):
        plan = _lib.Plan(3, case["r"], case["rN"], case["lower"], case["upper"])
        plan.prepare()
        out = str(tmp_path / "n3.likelihoods")
        dump_values_n3(plan, plan.rows(), out, leaves_per_window=(1 << 18) if case["m"] % 2 else 300)  # (one / many windows)
        want = []
        for s in case["stream"]:
            if s[2] is None:
                continue
            col1 = "0" * case["m"] if s[0] == "bogus" else s[0][0::2]
            want.append((col1, s[1][0], s[2]))
        got = _dump_lines(out)
        assert [g[0] for g in got] == [w[0] for w in want]
        for g, w in zip(got, want):
            assert abs(float(g[1]) - w[1]) <= 1e-9, (g, w)
            assert (float(g[2]) != float(g[2])) if w[2] != w[2] else abs(float(g[2]) - w[2]) <= 1e-10 * abs(w[2]), (g, w)
        plan.close()


def test_time_estimate_counts_messages_and_aborts(capsys, golden):
    """TimeEstimate.time_estimate (TimeEstimate.py:40-89): exact n=2 count, the reference's n=3 estimate, the printed
    ETA, the n=3 / m > 30 guard and the > 200 h abort without --FORCE."""
    import gzip, json
    from theta_b200.time_estimate import count_number_matrices_2, count_number_matrices_3, time_estimate
    with open(os.path.join(GOLD, "config1_search.json")) as f:
        inp = json.load(f)["inputs"]
    m = len(inp["r"])
    res = time_estimate(2, m, 3, 2, list(inp["lower"]), list(inp["upper"]), inp["r"], inp["rN"], 1.0, list(range(m)), 1, True, False)
    assert res["count"] == 142560 and res["rate"] > 1e8 and res["seconds"] < 1
    text = capsys.readouterr().out
    assert "Estimating time..." in text and "Estimated Total Time: 0 second(s)" in text
    # values of the live reference's count_number_matrices_3 (Python 2 integer division) on the golden spaces
    cases = golden("n3_small_full.json.gz")
    assert [count_number_matrices_3(c["m"], c["upper"], c["lower"]) for c in cases] == [2600, 6190, 33535]
    c = cases[0]
    res = time_estimate(3, c["m"], 3, 2, list(c["lower"]), list(c["upper"]), c["r"], c["rN"], 1.0, c["sorted_index"], 1, True, False)
    assert res["count"] == 2600
    capsys.readouterr()
    # n = 3 with more than 30 intervals needs --FORCE (TimeEstimate.py:48-50)
    big = 31
    with pytest.raises(SystemExit) as ex:
        time_estimate(3, big, 3, 2, [1] * big, [2] * big, [1000] * big, [1000] * big, 1.0, list(range(big)), 1, True, False)
    assert ex.value.code == 1 and "Run with --FORCE to continue" in capsys.readouterr().out
    # C(1007, 7) ~ 2e17 candidates: hundreds of hours even at 1e10 candidates/s -> abort unless forced (:74-85)
    mm = 1000
    rr = [1000 + 3 * i for i in range(mm)]
    rn = [1000] * mm
    assert count_number_matrices_2(mm, [7] * mm, [0] * mm) == 204032533091695451  # C(1007, 7)
    with pytest.raises(SystemExit) as ex:
        time_estimate(2, mm, 7, 2, [0] * mm, [7] * mm, rr, rn, 1.0, list(range(mm)), 1, True, False)
    out = capsys.readouterr().out
    assert ex.value.code == 1 and "hour(s)" in out and "the runtime is likely excessive" in out
    res = time_estimate(2, mm, 7, 2, [0] * mm, [7] * mm, rr, rn, 1.0, list(range(mm)), 1, True, True)
    assert res["seconds"] > 200 * 3600 and "hour(s)" in capsys.readouterr().out
