"""Where the end-to-end time of do_optimization_single goes (development aid; wall clock per phase)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from theta_b200 import _lib, search
_lib.init(0)
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
P = bench.problem(wl)
T = {}
def tick(name, t0):
    T.setdefault(name, []).append(time.perf_counter() - t0)
for rep in range(25):
    t0 = time.perf_counter(); plan = _lib.Plan(2, P["r"], P["rN"], P["lower"], P["upper"]); tick("plan_create", t0)
    t0 = time.perf_counter(); plan.bounds(); N = plan.count(); tick("bounds+count", t0)
    t0 = time.perf_counter(); g = plan.pass1(0, N); tick("pass1", t0)
    t0 = time.perf_counter(); recs = plan.pass2(g); tick("pass2", t0)
    t0 = time.perf_counter(); idx = search.fold_replay(recs["rank"], recs["nll"]); tick("fold", t0)
    uniq = sorted(set(int(recs["rank"][i]) for i in idx))
    t0 = time.perf_counter(); c, v = plan.eval_ranks(uniq, want_vals=True); tick("eval_ranks", t0)
    t0 = time.perf_counter(); cols = [plan.enumerate(rk, 1) for rk in uniq]; tick("enumerate", t0)
    t0 = time.perf_counter(); plan.close(); tick("close", t0)
    t0 = time.perf_counter()
    best = search.do_optimization_single(2, P["m"], P["k"], 2, list(P["lower"]), list(P["upper"]), P["r"], P["rN"], 1.0, P["sorted_index"], True, False)
    tick("do_optimization_single", t0)
for k, v in T.items():
    v = sorted(v[5:])
    print("%-24s median %.3f ms  min %.3f ms" % (k, 1e3 * v[len(v) // 2], 1e3 * v[0]))
