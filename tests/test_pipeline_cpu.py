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
