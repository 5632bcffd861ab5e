"""CPU: host-side pre/post steps (theta_b200/pipeline.py) against the golden inputs derived by the reference."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_config1_prestep_matches_reference(golden):
    from theta_b200 import pipeline
    g = golden("config1_search.json")["inputs"]
    lengths, tumor, normal, m, ub, lb = pipeline.read_interval_file(os.path.join(GOLD, "Example.intervals"))
    assert m == 84 and ub is None and lb is None
    order, lengths, tumor, normal = pipeline.select_intervals_n2(lengths, tumor, normal, m, 3, False, 100)
    assert order == g["order"] and len(order) == 61
    r, rN, si = pipeline.sort_r(normal, tumor)
    assert r == g["r"] and rN == g["rN"] and si == g["sorted_index"]
    up, lo = pipeline.calculate_bounds_heuristic(0.5, r, rN, len(r), 2, 3)
    assert up == g["upper"] and lo == g["lower"]
    assert pipeline.reverse_sort_list(pipeline.sort_by_sorted_index(list(range(61)), si), si) == list(range(61))


def test_find_mins_and_frac():
    from theta_b200 import pipeline
    a = [("c1", (0.5, 0.5), 10.0, [])]
    b = [("c2", (0.4, 0.6), 10.0005, [])]
    c = [("c3", (0.3, 0.7), 9.0, [])]
    assert pipeline.find_mins([a, b]) == a + b
    assert pipeline.find_mins([a, b, c]) == c
    assert pipeline.find_mins([[], a]) == a
    assert abs(pipeline.determine_frac_copy_num([10, 10, 10, 10], [10, 10, 20, 10], [1, 1, 2, 1], 0.1) - 1.0) < 1e-12


def test_model_selection_matches_reference(tmp_path):
    """ModelSelection (python/ModelSelection.py:43-149) on the reference's own n=2 / n=3 result files of the example:
    same choice, same BEST file."""
    import os
    import shutil
    from theta_b200 import pipeline
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    shutil.copy(os.path.join(gold, "config1_cli.n2.results"), os.path.join(str(tmp_path), "ex.n2.results"))
    shutil.copy(os.path.join(gold, "cli_n3.results"), os.path.join(str(tmp_path), "ex.n3.results"))
    sel, path = pipeline.model_selection(os.path.join(gold, "Example.intervals"), os.path.join(str(tmp_path), "ex.n2.results"),
                                         os.path.join(str(tmp_path), "ex.n3.results"))
    with open(path) as f, open(os.path.join(gold, "cli_BEST.results")) as g:
        assert f.read() == g.read()
    assert path.endswith("ex.BEST.results")
    # the criteria themselves
    C = [[2, "2", "2"], [2, "3", "3"], [2, "1", "2"], [2, "X", "X"]]
    assert pipeline.get_frac_breakdown(C, [10, 20, 30, 40]) == (0.1, 0.2, 0.3)
    # a forced n=3 win: penalised NLL favours n=3 and the criteria hold
    with open(os.path.join(str(tmp_path), "b.n2.results"), "w") as f:
        f.write("#NLL\tmu\tC\tp*\n9e9\t0.5,0.5\t2:3\t0.5,0.5\n")
    with open(os.path.join(str(tmp_path), "b.n3.results"), "w") as f:
        f.write("#NLL\tmu\tC\tp*\n1e9\t0.4,0.3,0.3\t3,3:1,2\t0.5,0.5\n")
    with open(os.path.join(str(tmp_path), "b.intervals"), "w") as f:
        f.write("#ID\tchrm\tstart\tend\ttumorCount\tnormalCount\n1\t1\t0\t100\t10\t10\n2\t1\t100\t200\t10\t10\n")
    sel, path = pipeline.model_selection(os.path.join(str(tmp_path), "b.intervals"), os.path.join(str(tmp_path), "b.n2.results"),
                                         os.path.join(str(tmp_path), "b.n3.results"))
    assert sel == 3 and open(path).read() == open(os.path.join(str(tmp_path), "b.n3.results")).read()


def test_cli_argument_surface(tmp_path):
    """The reference's flags (FileIO.parse_arguments, python/FileIO.py:45-227): names, defaults, and the quirks that
    are kept (n=3 default of 20 intervals only when -n 3 is given, FileIO.py:170) or fixed (K up to 7, Q4)."""
    import pytest
    from theta_b200 import run_theta
    f = tmp_path / "x.intervals"
    f.write_text("#ID\tchrm\tstart\tend\ttumorCount\tnormalCount\n1\t1\t0\t10\t5\t5\n2\t1\t10\t20\t6\t5\n")
    a = run_theta.parse_arguments([str(f)])
    assert (a.N, a.MAX_K, a.TAU, a.MAX_NORMAL, a.NUM_PROCESSES, a.NUM_INTERVALS) == (None, 3, 2, 1.0, 1, 100)
    assert a.OUTPUT_PREFIX == "x" and not a.BOUNDS_ONLY and not a.NO_MULTI_EVENT and not a.FORCE
    assert run_theta.parse_arguments([str(f), "-n", "3"]).NUM_INTERVALS == 20
    assert run_theta.parse_arguments([str(f), "-n", "3", "--NUM_INTERVALS", "8"]).NUM_INTERVALS == 8
    assert run_theta.parse_arguments([str(f), "-n", "2", "-k", "7", "-m", "0.6", "-p", "pre"]).MAX_K == 7
    for bad in (["-n", "4"], ["-k", "8"], ["-m", "1.5"], ["--BAF"]):
        with pytest.raises(SystemExit):
            run_theta.parse_arguments([str(f)] + bad)
    with pytest.raises(SystemExit):
        run_theta.parse_arguments([str(tmp_path / "missing.intervals")])
