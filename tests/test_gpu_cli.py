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
