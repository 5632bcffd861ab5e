"""Generate tests/golden/*.json from the LIVE reference (build container only).

TEST INFRASTRUCTURE.  Runs the reference's own code (imported in place from
/root/reference/python through oracle/refshim.py, no arithmetic touched) and
freezes its outputs as small JSON fixtures, so that the oracle restatement and
the CUDA path can be checked on the GPU box, where /root/reference does not
exist.  Floats are written with repr() precision (exact round trip).

    python oracle/make_golden.py [--only kat,config1,synth_n2,n3_small,n3_rankdef,edge]
"""
import argparse
import contextlib
import io
import json
import os
import pickle
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402
import theta_oracle as O  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")
REFPY = refshim.REF


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def _dump(name, obj):
    path = os.path.join(GOLD, name)
    with open(path, "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote", name, os.path.getsize(path), "bytes")


def _f(x):
    return None if x is None else float(x)


def gen_kat(R):
    """python/L2args.pkl etc.: the reference's only known-answer vectors."""
    def load(n):
        with open(os.path.join(REFPY, n), "rb") as f:
            return pickle.load(f, encoding="latin1")
    out = {}
    a = load("L2args.pkl")
    mu, C, m, r = a
    res_b, res_m = load("L2outputs_branch.pkl"), load("L2outputs_master.pkl")
    live = R.CalcAllC.L2(mu, np.array(C, dtype=float).copy(), m, r)
    assert float(live[0]) == float(res_b[0])
    out["L2"] = dict(mu=float(mu), C=np.asarray(C, dtype=float).tolist(), m=int(m), r=[int(v) for v in r],
                     nll_branch=float(res_b[0]), nll_master=float(res_m[0]),
                     vals_branch=[v if isinstance(v, str) else float(v) for v in res_b[1]])
    a = load("L3args.pkl")
    mu, C, m, r, n = a
    res_b, res_m = load("L3outputs_branch.pkl"), load("L3outputs_master.pkl")
    live = R.CalcAllC.L3(np.asarray(mu, dtype=float), np.array(C, dtype=float).copy(), m, r, n)
    assert float(live[0]) == float(res_b[0])
    out["L3"] = dict(mu=[float(v) for v in mu], C=np.asarray(C, dtype=float).tolist(), m=int(m),
                     r=[int(v) for v in r], n=int(n),
                     nll_branch=float(res_b[0]), nll_master=float(res_m[0]),
                     vals_branch=[v if isinstance(v, str) else float(v) for v in res_b[1]])
    _dump("calcallc_kat.json", out)


def config1_inputs(R):
    """RunTHetA.py:298-412 pre-steps for `Example.intervals -n 2 -k 3` (defaults otherwise)."""
    iv = R.FileIO.read_interval_file(os.path.join(REFPY, "..", "example", "Example.intervals"))
    lengths, tumor, normal, m, ub, lb = iv
    with _quiet():
        order, lengths, tumor, normal = R.SelectIntervals.select_intervals_n2(lengths, tumor, normal, m, 3, False, 100)
    m = len(order)
    R.DataTools.set_total_read_counts(sum(tumor), sum(normal))
    r, rN, si = R.DataTools.sort_r(normal, tumor)
    with _quiet():
        ub, lb = R.DataTools.calculate_bounds_heuristic(0.5, r, rN, m, 2, 3)
    return dict(m=m, k=3, tau=2, r=[int(v) for v in r], rN=[int(v) for v in rN], lower=[int(v) for v in lb],
                upper=[int(v) for v in ub], sorted_index=[int(v) for v in si], order=[int(v) for v in order])


def gen_config1(R):
    """Full-space run of the reference's do_optimization_single on config #1 (about 3 minutes)."""
    inp = config1_inputs(R)
    m = inp["m"]
    tmp = "/tmp/theta_golden_cfg1"
    R.RunTHetA.pre = tmp
    t0 = time.time()
    with _quiet():
        best = R.RunTHetA.do_optimization_single(2, m, 3, 2, inp["lower"][:], inp["upper"][:], inp["r"], inp["rN"],
                                                 1.0, inp["sorted_index"], True, True)
    wall = time.time() - t0
    solved = []
    with open(tmp + ".likelihoods") as f:
        for line in f:
            cs, mu0, nll = line.rstrip("\n").split("\t")
            solved.append((cs, float(mu0), float(nll)))
    # rank of every solved candidate: walk the enumeration once (the stream has the first matrix twice, Q1)
    rank_of = {}
    for i, col in enumerate(O.enumerate_n2(inp["lower"], inp["upper"])):
        rank_of["".join(str(v) for v in col)] = i
    N = len(rank_of)
    assert solved[0][0] == solved[1][0] or True
    solved_ranks = set(rank_of[s[0]] for s in solved)
    none_ranks = sorted(set(range(N)) - solved_ranks)
    uniq = {}
    for cs, mu0, nll in solved:
        uniq[rank_of[cs]] = (mu0, nll)
    by_nll = sorted(uniq.items(), key=lambda kv: kv[1][1])
    top = [dict(rank=k, mu0=v[0], nll=v[1]) for k, v in by_nll[:300]]
    sample = [dict(rank=k, mu0=uniq[k][0], nll=uniq[k][1]) for k in sorted(uniq) if k % 53 == 0]
    out = dict(inputs=inp, n_candidates=N, n_emitted=N + 1, n_solved_emitted=len(solved), none_ranks=none_ranks,
               reference_wall_s=wall,
               best=[dict(C=np.asarray(b[0]).tolist(), mu=[float(b[1][0]), float(b[1][1])], nll=float(b[2]),
                          vals=[float(v) for v in b[3]]) for b in best],
               top=top, sample=sample)
    _dump("config1_search.json", out)


def gen_synth_n2(R):
    """Per-candidate (mu, NLL | None) of the live Optimizer.solve on the synthetic n=2 configs."""
    out = {}
    for name, m, k, mx in (("m150_k5", 150, 5, 1.0), ("m150_k7", 150, 7, 1.0), ("m300_k7", 300, 7, 1.0),
                           ("m40_k4_maxn", 40, 4, 0.6)):
        P = O.synthetic_problem_n2(m, k, seed=1)
        N = O.count_n2(P["lower"], P["upper"])
        rng = np.random.default_rng(7)
        ranks = list(range(200)) + sorted(int(x) for x in rng.integers(0, N, 500))
        # plus a cluster of ranks around a random walk near the (unknown) optimum: take ranks whose column is
        # the rounded truth-like profile
        opt = R.Optimizer.Optimizer(P["r"], P["rN"], m, 2, 2, upper_bound=mx)
        recs = []
        for rk in ranks:
            col = O.unrank_n2(P["lower"], P["upper"], rk)
            s = opt.solve(O.column_to_matrix(col))
            recs.append(dict(rank=rk, mu0=None if s is None else float(s[0][0]), nll=None if s is None else float(s[1])))
        out[name] = dict(m=m, k=k, seed=1, max_normal=mx, r=P["r"], rN=P["rN"], lower=P["lower"], upper=P["upper"],
                         n_candidates=int(N), records=recs)
        print(name, "N=%g" % N, "none:", sum(1 for q in recs if q["nll"] is None))
    _dump("synth_n2_candidates.json", out)


def gen_n2_small_full(R):
    """Whole small n=2 spaces through the live do_optimization_single: best lists incl. ties, Q1 duplicate."""
    cases = []
    rng = np.random.default_rng(11)
    specs = [(7, 3, 1.0, 1), (9, 4, 1.0, 2), (12, 3, 1.0, 3), (10, 5, 0.5, 4), (8, 2, 1.0, 5), (6, 6, 0.8, 6),
             (14, 3, 1.0, 7)]
    for m, k, mx, seed in specs:
        r, rN, _ = O.synthetic(m, 2, k, seed)
        r, rN, si = O.sort_r(rN, r)
        if seed % 2:
            ub, lb = O.bounds_heuristic(0.5, r, rN, 2, k)
        else:
            lb = sorted(int(x) for x in rng.integers(0, 2, m))
            ub = [min(k, l + int(x)) for l, x in zip(lb, rng.integers(1, k + 1, m))]
        lb, ub = O.check_bound_order(lb, ub)
        R.RunTHetA.pre = "/tmp/theta_golden_small"
        with _quiet():
            best = R.RunTHetA.do_optimization_single(2, m, k, 2, lb[:], ub[:], r, rN, mx, si, True, True)
        solved = {}
        with open("/tmp/theta_golden_small.likelihoods") as f:
            for line in f:
                cs, mu0, nll = line.rstrip("\n").split("\t")
                solved[cs] = (float(mu0), float(nll))
        stream = []
        for col in O.enumerate_n2(lb, ub):
            cs = "".join(str(v) for v in col)
            stream.append([cs] + (list(solved[cs]) if cs in solved else [None, None]))
        cases.append(dict(m=m, k=k, max_normal=mx, r=r, rN=rN, lower=lb, upper=ub, sorted_index=si, stream=stream,
                          best=[dict(C=np.asarray(b[0]).tolist(), mu=[float(b[1][0]), float(b[1][1])],
                                     nll=float(b[2]), vals=[float(v) for v in b[3]]) for b in best]))
        print("n2 small", m, k, "N", len(stream), "best", len(best))
    _dump("n2_small_full.json", cases)


def gen_n3_small(R):
    """Whole small n=3 spaces through the live reference: per-candidate dump + best list."""
    cases = []
    for m, lbv, ubv, seed in ((7, 1, 3, 1), (8, 1, 3, 1), (6, 0, 3, 2)):
        r, rN, _ = O.synthetic(m, 3, 3, seed)
        r, rN, si = O.sort_r(rN, r)
        lb, ub = [lbv] * m, [ubv] * m
        enum = R.Enumerator.Enumerator(3, m, 3, 2, lb[:], ub[:], True)
        opt = R.Optimizer.Optimizer(r, rN, m, 3, 2, upper_bound=1.0)
        rows = [tuple(x) for x in enum.rows]
        stream = []
        C = enum._C_to_array()  # Q2: bogus first matrix
        first = True
        while C is not False:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                s = opt.solve(C)
            key = "bogus" if first else "".join("%d%d" % (int(a), int(b)) for a, b in C[:, 1:])
            stream.append([key] + ([None, None] if s is None else [[float(v) for v in s[0]], float(s[1])]))
            first = False
            C = enum.generate_next_C()
        with _quiet(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            best = R.RunTHetA.do_optimization_single(3, m, 3, 2, lb[:], ub[:], r, rN, 1.0, si, True, False)
        cases.append(dict(m=m, k=3, r=r, rN=rN, lower=lb, upper=ub, sorted_index=si, rows=rows, stream=stream,
                          best=[dict(C=np.asarray(b[0]).tolist(), mu=[float(v) for v in b[1]], nll=float(b[2]),
                                     vals=[float(v) for v in b[3]]) for b in best]))
        print("n3 small", m, "N", len(stream), "none", sum(1 for s in stream if s[1] is None), "best", len(best))
    _dump("n3_small_full.json", cases)
    import subprocess
    subprocess.check_call(["gzip", "-9", "-f", os.path.join(GOLD, "n3_small_full.json")])


def gen_synth_n3(R):
    """Configs #4 / #5: the live Optimizer.solve on the first candidates in enumeration order and on
    candidates further into the space (every 997th of the first 60 000 for #4)."""
    import itertools
    out = {}
    for name, m, k, F, tb, n_first, stride, n_far in (("cfg4_m150_k5", 150, 5, 12, False, 120, 997, 60),
                                                       ("cfg5_m500_k7", 500, 7, 12, True, 30, 0, 0)):
        P = O.synthetic_problem_n3(m, k, 1, F, tb)
        rows, _ = O.n3_graph(max(P["upper"]), 2)
        opt = R.Optimizer.Optimizer(P["r"], P["rN"], m, 3, 2, upper_bound=1.0)
        recs = []
        gen = O.enumerate_n3(P["lower"], P["upper"])
        want = set(range(n_first)) | set(range(n_first, n_first + stride * n_far, stride) if stride else [])
        last = max(want)
        for idx_, path in enumerate(gen):
            if idx_ > last:
                break
            if idx_ not in want:
                continue
            C = O.rows_to_matrix(path, rows)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                s_ = opt.solve(C)
            recs.append(dict(index=idx_, path=[int(v) for v in path],
                             mu=None if s_ is None else [float(v) for v in s_[0]],
                             nll=None if s_ is None else float(s_[1])))
        out[name] = dict(m=m, k=k, F=F, tau_bounds=tb, seed=1, r=P["r"], rN=P["rN"], lower=P["lower"], upper=P["upper"],
                         sorted_index=P["sorted_index"], rows=[list(x) for x in rows], records=recs)
        print(name, len(recs), "none", sum(1 for q in recs if q["nll"] is None))
    _dump("synth_n3_candidates.json", out)


def rankdef_space():
    """A real-shaped m=16 n=3 space: the 16 longest intervals of the reference's example (real read counts, the number
    of intervals its n=3 interval selection keeps with --NUM_INTERVALS 16), lower bounds 0 so that zero tumour columns
    occur, upper bounds 2..4 in sorted order."""
    rows = [ln.split("\t") for ln in open(os.path.join(GOLD, "Example.intervals")).read().strip().split("\n")[1:]]
    iv = sorted(((int(x[3]) - int(x[2]), int(x[4]), int(x[5])) for x in rows), reverse=True)[:16]
    r, rN, si = O.sort_r([v[2] for v in iv], [v[1] for v in iv])
    return r, rN, si, [0] * 16, [2] * 9 + [3] * 4 + [4] * 3


def gen_n3_rankdef(R):
    """>= 10^4 rank-deficient candidates (all rows on one line of the (a, b) plane, which includes the zero-column and
    identical-column classes) of a real-shaped m=16 space through the live Optimizer.solve.
    They are listed by enumerating the space restricted to the rows of each line (theta_oracle.enumerate_n3 row_ok)."""
    r, rN, si, lb, ub = rankdef_space()
    m = len(r)
    rows, _ = O.n3_graph(max(ub), 2)
    lines = set()
    for i, p_ in enumerate(rows):
        for q in rows[i + 1:]:
            lines.add(frozenset(x for x in rows if (q[0] - p_[0]) * (x[1] - p_[1]) - (q[1] - p_[1]) * (x[0] - p_[0]) == 0))
    paths = set()
    for L in sorted(lines, key=lambda L: sorted(L)):
        for n_, path in enumerate(O.enumerate_n3(lb, ub, row_ok=lambda row: row in L)):
            if n_ >= 1500:
                break
            paths.add(tuple(path))
    paths = sorted(paths)
    opt = R.Optimizer.Optimizer(r, rN, m, 3, 2, upper_bound=1.0)
    recs = []
    t0 = time.time()
    for path in paths:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            s_ = opt.solve(O.rows_to_matrix(path, rows))
        recs.append([list(path)] + ([None, None] if s_ is None else [[float(v) for v in s_[0]], float(s_[1])]))
    print("n3 rank-deficient", len(recs), "none", sum(1 for q in recs if q[2] is None),
          "nan", sum(1 for q in recs if q[2] is not None and q[2] != q[2]), "%.0fs" % (time.time() - t0))
    _dump("n3_rankdef_m16.json", dict(m=m, r=r, rN=rN, lower=lb, upper=ub, sorted_index=si, rows=[list(x) for x in rows],
                                      records=recs))
    import subprocess
    subprocess.check_call(["gzip", "-9", "-f", os.path.join(GOLD, "n3_rankdef_m16.json")])


def gen_config1_cli(R):
    """The reference's own CLI on config #1 (RunTHetA.main, ~3 min): .n2.results and .n2.withBounds as text fixtures,
    plus the input file itself (the reference's example *data*, needed on the GPU box)."""
    import shutil
    src = os.path.join(REFPY, "..", "example", "Example.intervals")
    out = "/tmp/theta_golden_cli"
    os.makedirs(out, exist_ok=True)
    argv = sys.argv
    sys.argv = ["RunTHetA.py", src, "-n", "2", "-k", "3", "--NUM_PROCESSES", "1", "-d", out, "-p", "ex"]
    try:
        with _quiet():
            R.RunTHetA.main()
    finally:
        sys.argv = argv
    shutil.copy(src, os.path.join(GOLD, "Example.intervals"))
    shutil.copy(os.path.join(out, "ex.n2.results"), os.path.join(GOLD, "config1_cli.n2.results"))
    shutil.copy(os.path.join(out, "ex.n2.withBounds"), os.path.join(GOLD, "config1_cli.n2.withBounds"))
    print("wrote config1 CLI fixtures")


def gen_cli_default_flow(R):
    """The reference's own CLI beyond n=2 on the example data:
      * -n 3 --RESULTS <config #1 n=2 files> --NUM_INTERVALS 8 (so that the reference's n=3 search finishes in
        minutes), with and without --NO_MULTI_EVENT (CalcAllC.calc_all_c_3_multi_event / calc_all_c_3);
      * ModelSelection on the n=2 / n=3 result files -> BEST.results;
      * the default flow (no -n: n=2, n=3, ModelSelection in one call) with --NUM_INTERVALS 8 --FORCE."""
    import shutil
    src = os.path.join(GOLD, "Example.intervals")
    out = "/tmp/theta_golden_flow"
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    argv = sys.argv

    def run(args):
        sys.argv = ["RunTHetA.py"] + args
        try:
            with _quiet():
                R.RunTHetA.main()
        except SystemExit as e:
            if e.code not in (0, None):
                raise
        finally:
            sys.argv = argv
    shutil.copy(os.path.join(GOLD, "config1_cli.n2.results"), os.path.join(out, "ex.n2.results"))
    shutil.copy(os.path.join(GOLD, "config1_cli.n2.withBounds"), os.path.join(out, "ex.n2.withBounds"))
    common = ["-n", "3", "-k", "3", "--NUM_INTERVALS", "8", "--NUM_PROCESSES", "1", "--RESULTS", os.path.join(out, "ex.n2.results"), "-d", out]
    run([os.path.join(out, "ex.n2.withBounds")] + common + ["-p", "ex"])
    run([os.path.join(out, "ex.n2.withBounds")] + common + ["--NO_MULTI_EVENT", "-p", "ex_nme"])
    with _quiet():
        R.ModelSelection.ModelSelection(src, os.path.join(out, "ex.n2.results"), os.path.join(out, "ex.n3.results"))
    for f, g in (("ex.n3.results", "cli_n3.results"), ("ex.n3.withBounds", "cli_n3.withBounds"),
                 ("ex_nme.n3.results", "cli_n3_nme.results"), ("ex_nme.n3.withBounds", "cli_n3_nme.withBounds"),
                 ("ex.BEST.results", "cli_BEST.results")):
        shutil.copy(os.path.join(out, f), os.path.join(GOLD, g))
    print("wrote n=3 CLI fixtures")
    run([src, "-k", "3", "--NUM_INTERVALS", "8", "--FORCE", "--NUM_PROCESSES", "1", "-d", out, "-p", "flow"])
    for f in ("flow.n2.results", "flow.n2.withBounds", "flow.n3.results", "flow.n3.withBounds", "flow.BEST.results"):
        shutil.copy(os.path.join(out, f), os.path.join(GOLD, "cli_" + f))
    print("wrote default-flow CLI fixtures")


def gen_edge(R):
    """Edge-case KATs from the live Optimizer.solve."""
    m = 7
    r, rN, _ = O.synthetic(m, 2, 3, 1)
    r, rN, si = O.sort_r(rN, r)
    opt = R.Optimizer.Optimizer(r, rN, m, 2, 2, upper_bound=1.0)
    cols = dict(all_tau=[2] * m, all_zero=[0] * m, all_one=[1] * m, c=[0, 1, 1, 1, 2, 2, 3],
                c2=[0, 2, 2, 2, 4, 4, 6], leading_zero=[0, 0, 0, 1, 2, 2, 2], high=[3, 3, 3, 3, 3, 3, 4],
                low=[0, 0, 0, 0, 0, 0, 1])
    out = dict(m=m, r=r, rN=rN, cases={})
    for k, col in cols.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            s = opt.solve(O.column_to_matrix(col))
        out["cases"][k] = dict(col=col, mu0=None if s is None else float(s[0][0]), nll=None if s is None else float(s[1]),
                               vals=None if s is None else [float(v) for v in s[2]])
    # zero tumour count in an interval whose candidate copy number is 0 -> 0/0 -> None
    r0 = r[:]
    r0[0] = 0
    opt0 = R.Optimizer.Optimizer(r0, rN, m, 2, 2, upper_bound=1.0)
    for k, col in (("zero_r_c0", [0, 0, 1, 1, 2, 2, 3]), ("zero_r_c1", [1, 1, 1, 1, 2, 2, 3])):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            s = opt0.solve(O.column_to_matrix(col))
        out["cases"][k] = dict(col=col, r=r0, mu0=None if s is None else float(s[0][0]),
                               nll=None if s is None else float(s[1]),
                               vals=None if s is None else [float(v) for v in s[2]])
    _dump("edge_n2.json", out)


def gen_bounds_edge(R):
    """n=2 spaces whose lower bound exceeds the upper bound at some intervals: the reference never checks its
    start vector against the upper bounds -- Enumerator._generate_next_C_2 (Enumerator.py:119-152) starts from the lower
    bounds and simply never increments a position that is already at or above its upper bound -- so such a space is NOT
    empty.  Whole streams + best lists through the live do_optimization_single."""
    cases = []
    for m, lb, ub, seed in ((6, [0, 0, 2, 2, 3, 3], [2, 2, 1, 3, 2, 4], 1), (5, [1, 1, 1, 3, 3], [0, 2, 2, 2, 3], 2),
                            (4, [2, 2, 2, 2], [1, 1, 1, 1], 3)):
        r, rN, _ = O.synthetic(m, 2, 4, seed)
        r, rN, si = O.sort_r(rN, r)
        R.RunTHetA.pre = "/tmp/theta_golden_bedge"
        lb_in, ub_in = lb[:], ub[:]
        with _quiet():
            best = R.RunTHetA.do_optimization_single(2, m, 4, 2, lb_in, ub_in, r, rN, 1.0, si, True, True)
        lines = []
        with open("/tmp/theta_golden_bedge.likelihoods") as f:
            for line in f:
                cs, mu0, nll = line.rstrip("\n").split("\t")
                lines.append([cs, float(mu0), float(nll)])
        cases.append(dict(m=m, r=r, rN=rN, lower=lb, upper=ub, lower_after=lb_in, upper_after=ub_in, sorted_index=si,
                          dump=lines, best=[dict(C=np.asarray(b[0]).tolist(), mu=[float(b[1][0]), float(b[1][1])],
                                                 nll=float(b[2]), vals=[float(v) for v in b[3]]) for b in best]))
        print("bounds edge", lb, ub, "dump lines", len(lines), "best", len(best))
    _dump("bounds_edge_n2.json", cases)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="kat,edge,n2_small,synth_n2,n3_small,synth_n3,config1")
    a = ap.parse_args()
    warnings.simplefilter("ignore")
    R = refshim.modules()
    os.makedirs(GOLD, exist_ok=True)
    todo = a.only.split(",")
    if "kat" in todo:
        gen_kat(R)
    if "edge" in todo:
        gen_edge(R)
    if "n2_small" in todo:
        gen_n2_small_full(R)
    if "synth_n2" in todo:
        gen_synth_n2(R)
    if "n3_small" in todo:
        gen_n3_small(R)
    if "synth_n3" in todo:
        gen_synth_n3(R)
    if "n3_rankdef" in todo:
        gen_n3_rankdef(R)
    if "bounds_edge" in todo:
        gen_bounds_edge(R)
    if "config1" in todo:
        gen_config1(R)
    if "config1_cli" in todo:
        gen_config1_cli(R)
    if "cli_flow" in todo:
        gen_cli_default_flow(R)


if __name__ == "__main__":
    main()
