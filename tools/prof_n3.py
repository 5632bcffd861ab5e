"""Short n=3 driver for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import theta_oracle as O
from theta_b200 import _lib
_lib.init(0)
P = O.synthetic_problem_n3(150, 5, 1, 12, False)
plan = _lib.Plan(3, P["r"], P["rN"], P["lower"], P["upper"])
ni = plan.prepare(512)
n = int(sys.argv[1]) if len(sys.argv) > 1 else ni
for _ in range(2):
    g = plan.search_min_n3(0, min(n, ni))
st = plan.stats(); print("done items", ni, "min", g, "kernel ms %.1f" % st["ms_pass1"], "cand/s %.3e" % (st["n_candidates"] / (st["ms_pass1"] * 1e-3)), "exact", st["n_exact"], "of", st["n_candidates"], "violations", st["n_bound_violations"])
