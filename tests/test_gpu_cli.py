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
