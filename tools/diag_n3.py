"""Diagnostic: n=3 GPU path vs the golden streams of the live reference."""
import gzip, json, math, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from theta_b200 import _lib
_lib.init(0)
d = json.load(gzip.open(os.path.join(ROOT, "tests/golden/n3_small_full.json.gz"), "rt"))
for c in d:
    m = c["m"]
    plan = _lib.Plan(3, c["r"], c["rN"], c["lower"], c["upper"])
    rows = plan.rows()
    assert rows == [tuple(x) for x in c["rows"]], (rows, c["rows"])
    ni = plan.prepare(64)
    paths = plan.enumerate_n3()
    keys = ["".join("%d%d" % rows[g] for g in p) for p in paths]
    gold = c["stream"][1:]
    print("m", m, "items", ni, "enumerated", len(keys), "golden", len(gold), "order equal", keys == [s[0] for s in gold])
    mu, mh, nll, st, _ = plan.solve_rows(paths)
    cat = collections.Counter()
    worst_mu = worst_nll = 0
    for i, s in enumerate(gold):
        a = s[0][0::2]; b = s[0][1::2]
        zero = all(ch == "0" for ch in a) or all(ch == "0" for ch in b)
        eq = a == b
        gnone = s[2] is None
        dnone = st[i] >= 8
        tag = ("zero" if zero else "eq" if eq else "reg")
        if gnone != dnone:
            cat[tag + ":none-mismatch(g=%s,st=%d)" % (gnone, st[i])] += 1
            continue
        if gnone:
            cat[tag + ":both-none"] += 1
            continue
        if isinstance(s[2], float) and math.isnan(s[2]):
            cat[tag + ":golden-nan"] += 1
            continue
        en = abs(nll[i] - s[2]) / abs(s[2])
        em = max(abs(mu[i][j] - s[1][j]) for j in range(3))
        ok = en <= 1e-10 and em <= 1e-9
        cat[tag + (":ok" if ok else ":bad(st=%d)" % st[i])] += 1
        if tag == "reg":
            worst_mu = max(worst_mu, em); worst_nll = max(worst_nll, en)
            if not ok and cat["printed"] < 5:
                cat["printed"] += 1
                print("  BAD", s[0], "st", st[i], "mu", mu[i], s[1], "nll", nll[i], s[2])
    print("  ", dict(cat), "worst reg mu %.2e nll %.2e" % (worst_mu, worst_nll))
    plan.close()
