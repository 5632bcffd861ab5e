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
