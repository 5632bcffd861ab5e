"""Single-process multi-GPU check (NO torchrun): theta_mgpu_optimize on G GPUs must return exactly the single-GPU
answer of do_optimization_single, for n=2 and n=3.  Prints 'MGPU_CHECK OK <G> ...' per device count."""
import gzip
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402

from theta_b200 import _lib, search  # noqa: E402
import theta_oracle as O  # noqa: E402

_lib.init(0)
have = _lib.device_count()
counts = [int(x) for x in sys.argv[1:]] or [g for g in (1, 2, 4, 8) if g <= have]


def same(a, b):
    assert len(a) == len(b), (len(a), len(b))
    for x, y in zip(a, b):
        assert np.array_equal(x[0], y[0])
        assert np.array_equal(np.asarray(x[1], dtype=np.float64), np.asarray(y[1], dtype=np.float64)), (x[1], y[1])
        assert x[2] == y[2], (x[2], y[2])
        assert np.array_equal(np.asarray(x[3]), np.asarray(y[3]))


with open(os.path.join(ROOT, "tests/golden/config1_search.json")) as f:
    g1 = json.load(f)
inp = g1["inputs"]
cases2 = [("config1", dict(r=inp["r"], rN=inp["rN"], lower=inp["lower"], upper=inp["upper"], m=len(inp["r"]),
                           sorted_index=list(range(len(inp["r"]))))),
          ("m150k5", O.synthetic_problem_n2(150, 5, 1)), ("m60k4", O.synthetic_problem_n2(60, 4, 3))]
with gzip.open(os.path.join(ROOT, "tests/golden/n3_small_full.json.gz"), "rt") as f:
    cases3 = json.load(f)
P4 = O.synthetic_problem_n3(150, 5, 1, 12, False)

for G in counts:
    t0 = time.time()
    for name, P in cases2:
        m = len(P["r"])
        si = P.get("sorted_index", list(range(m)))
        one = search.do_optimization_single(2, m, 5, 2, list(P["lower"]), list(P["upper"]), P["r"], P["rN"], 1.0, si, True, False)
        many = search.do_optimization(2, m, 5, 2, list(P["lower"]), list(P["upper"]), P["r"], P["rN"], 1.0, si, G, True, False) \
            if G > 1 else search.best_from_records(2, m, 2, *_lib.MultiGpu(2, P["r"], P["rN"], P["lower"], P["upper"], n_devices=1).optimize()[2:5], si)
        same(one, many)
    for c in cases3:
        m = c["m"]
        one = search.do_optimization_single(3, m, c["k"], 2, list(c["lower"]), list(c["upper"]), c["r"], c["rN"], 1.0,
                                            c["sorted_index"], True, False)
        if G > 1:
            many = search.do_optimization(3, m, c["k"], 2, list(c["lower"]), list(c["upper"]), c["r"], c["rN"], 1.0,
                                          c["sorted_index"], G, True, False)
        else:
            mg = _lib.MultiGpu(3, c["r"], c["rN"], c["lower"], c["upper"], n_devices=1)
            many = search.best_from_records(3, m, 2, *mg.optimize()[2:5], c["sorted_index"], mg)
        same(one, many)
    # config #4 shape (34.5 M candidates): G GPUs == 1 GPU
    mg = _lib.MultiGpu(3, P4["r"], P4["rN"], P4["lower"], P4["upper"], n_devices=G)
    lo, up, recs, rows, vals, st, nc = mg.optimize()
    one = search.do_optimization_single(3, 150, 5, 2, list(P4["lower"]), list(P4["upper"]), P4["r"], P4["rN"], 1.0,
                                        P4["sorted_index"], True, False)
    same(one, search.best_from_records(3, 150, 2, recs, rows, vals, P4["sorted_index"], mg))
    t1 = time.perf_counter()
    mg.optimize()
    dt = time.perf_counter() - t1
    print("MGPU_CHECK OK %d  (n=2 x%d, n=3 x%d spaces identical to 1 GPU; cfg4 %d candidates in %.1f ms wall, device ms %s; %.1fs)"
          % (G, len(cases2), len(cases3) + 1, nc, dt * 1e3, np.round(mg.last_ms(), 2).tolist(), time.time() - t0), flush=True)
    mg.close()
