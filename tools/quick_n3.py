"""n=3 search timing on the synthetic configs (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import theta_oracle as O
from theta_b200 import _lib
from theta_b200.search_n3 import search_n3
_lib.init(0)
CAPS = [int(x) for x in os.environ.get("LEAF_CAPS", "512").split(",")]
for (m, k, F, tb, cap, maxit) in [(150, 5, 12, False, c, 1 << 22) for c in CAPS] + ([(500, 7, 12, True, 4096, 200000)] if "SKIP500" not in os.environ else []):
    P = O.synthetic_problem_n3(m, k, 1, F, tb)
    plan = _lib.Plan(3, P["r"], P["rN"], P["lower"], P["upper"])
    t0 = time.perf_counter(); ni = plan.prepare(cap, maxit); t1 = time.perf_counter()
    w = plan.item_weights()
    print("m", m, "items", ni, "truncated", plan.truncated, "est leaves %.3g" % w.sum(), "prepare %.2fs" % (t1 - t0), flush=True)
    for rep in range(2):
        t0 = time.perf_counter()
        g = plan.search_min_n3(0, ni)
        t1 = time.perf_counter()
        recs, paths = plan.window_n3(g)
        t2 = time.perf_counter()
        st = plan.stats()
        print("  pass1 %.1f ms (kernel %.1f) pass2 %.1f ms (kernel %.1f) leaves %d solved %d cand/s %.3e nfev %.2f rescanned %d recs %d min %.6f"
              % (1e3 * (t1 - t0), st["ms_pass1"], 1e3 * (t2 - t1), st["ms_pass2"], st["n_candidates"], st["n_solved"],
                 st["n_candidates"] / (st["ms_pass1"] * 1e-3), st["newton_iters"], st["n_rescanned"], len(recs), g), flush=True)
    plan.close()
