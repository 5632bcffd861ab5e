"""Kernel-only timing of the fused search on the bench workloads (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from theta_b200 import _lib
_lib.init(0)
for wl in sys.argv[1:] or ["cfg2", "m150k7"]:
    P = bench.problem(wl)
    plan = _lib.Plan(2, P["r"], P["rN"], P["lower"], P["upper"])
    N = plan.count()
    if P["window"]:
        N = min(N, P["window"])
    for _ in range(3):
        g = plan.pass1(0, N); plan.pass2(g)
    p1, p2, wall = [], [], []
    for _ in range(5):
        t0 = time.perf_counter()
        g = plan.pass1(0, N); t1 = time.perf_counter(); recs = plan.pass2(g); t2 = time.perf_counter()
        st = plan.stats(); p1.append(st["ms_pass1"]); p2.append(st["ms_pass2"]); wall.append((t1 - t0, t2 - t1))
    print(wl, "N=%d" % N, "pass1 ms %.3f" % min(p1), "pass2 ms %.3f" % min(p2), "cand/s kernel %.3e" % (N / min(p1) * 1e3),
          "iters %.3f" % st["newton_iters"], "host pass1 %.3f ms pass2 %.3f ms" % (1e3 * min(w[0] for w in wall), 1e3 * min(w[1] for w in wall)),
          "rescanned", st["n_rescanned"], "recs", len(recs), "exact", st["n_exact"], "minb", os.environ.get("THETA_SEARCH_MINB"), "min", g)
    plan.close()
