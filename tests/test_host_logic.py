"""CPU: host-side logic of the product (no CUDA calls): the C-ABI library loads and exports every
symbol include/theta_b200.h declares, the fold replay matches the sequential fold, sharding covers
the space, and the product fails loudly without a GPU."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from theta_b200 import _lib
    lib = _lib.load()
    with open(os.path.join(ROOT, "include", "theta_b200.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"\b(theta_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"theta_tile_pitch"} - set(_lib.SYMBOLS)
    assert declared, "no declarations parsed"
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, "symbols declared in the header but not exported: %s" % missing
    assert set(_lib.SYMBOLS) == declared, (sorted(set(_lib.SYMBOLS) ^ declared))
    assert lib.theta_version().startswith(b"theta_b200")
    assert lib.theta_tile_pitch(150) == 152 and lib.theta_tile_pitch(61) == 72 and lib.theta_tile_pitch(8) == 8
    assert ctypes.sizeof(_lib.Cand) == 48


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail loudly (never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from theta_b200 import _lib
    from theta_b200.search import do_optimization_single
    with pytest.raises(_lib.ThetaError):
        _lib.init(0)
    with pytest.raises(_lib.ThetaError):
        do_optimization_single(2, 3, 2, 2, [0, 0, 0], [2, 2, 2], [5, 6, 7], [5, 5, 5], 1.0, [0, 1, 2], True, False)
    with pytest.raises(_lib.ThetaError):
        _lib.optimize_n2([5, 6, 7], [5, 5, 5], [0, 0, 0], [2, 2, 2])
    with pytest.raises(_lib.ThetaError):
        do_optimization_single(3, 3, 2, 2, [0, 0, 0], [2, 2, 2], [5, 6, 7], [5, 5, 5], 1.0, [0, 1, 2], True, False)
    with pytest.raises(_lib.ThetaError):
        _lib.nll_single([0.3, 0.7], [[2.0, 1.0], [2.0, 3.0]], [4, 5])
    # nothing of the product imports the oracle
    import glob
    for path in glob.glob(os.path.join(ROOT, "theta_b200", "*.py")):
        src = open(path).read()
        assert "theta_oracle" not in src and "import oracle" not in src and "refshim" not in src, path


def test_fold_replay_matches_sequential_fold():
    import theta_oracle as O
    from theta_b200.search import fold_replay
    rng = np.random.default_rng(0)
    for trial in range(300):
        n = int(rng.integers(1, 60))
        base = rng.choice([0.0, 1e-3, 5e-4, 2e-3, 1.0, 7.0], size=n)
        nll = 1e9 + base * rng.integers(0, 4, size=n) + rng.choice([0, 0, 1e-4, -1e-4], size=n)
        if trial % 5 == 0:
            nll[int(rng.integers(0, n))] = np.nan
        ranks = np.sort(rng.choice(10 * n, size=n, replace=False))
        if trial % 2 == 0:
            ranks[0] = 0
        stream = []
        for rk, v in zip(ranks, nll):
            if rk == 0:
                stream.append((int(rk), float(v)))  # Q1: the first matrix is evaluated twice
            stream.append((int(rk), float(v)))
        want = O.fold(stream)
        got = [int(ranks[i]) for i in fold_replay(ranks, nll, dup_first=True)]
        assert got == want


def _records_of_shard(ranks, nlls, g, prior, margin=1e-3):
    """Python model of pass 2 (theta_search_n2_pass2): strict prefix minima w.r.t. `prior` plus near-minimum set."""
    keep, run = [], prior
    for rk, v in zip(ranks, nlls):
        if v < run or v <= g + 2.5 * margin or v != v:
            keep.append((rk, v))
        if v < run:
            run = v
    return keep


def test_sharded_replay_equals_sequential():
    """The two-pass decomposition: per-shard prefix minima + near-min set, replayed in rank
    order, gives exactly the sequential fold for any number of shards."""
    import theta_oracle as O
    from theta_b200.search import fold_replay, shard_range
    rng = np.random.default_rng(1)
    for trial in range(200):
        n = int(rng.integers(5, 400))
        steps = rng.choice([0.0, 3e-4, 6e-4, 1.1e-3, 0.5, 3.0], size=n, p=[.3, .2, .2, .1, .1, .1])
        nll = 5e9 + np.cumsum(rng.choice([-1, 1], size=n) * steps)
        seq = O.fold([(0, float(nll[0]))] + [(i, float(v)) for i, v in enumerate(nll)])
        g = float(np.nanmin(nll))
        for world in (1, 2, 3, 8):
            recs, mins = [], []
            for rk in range(world):
                b, e = shard_range(n, rk, world)
                prior = min(mins) if mins else float("inf")
                recs += _records_of_shard(range(b, e), nll[b:e], g, prior)
                mins.append(float(np.min(nll[b:e])) if e > b else float("inf"))
            rr = np.array([r[0] for r in recs])
            vv = np.array([r[1] for r in recs])
            got = [int(rr[i]) for i in fold_replay(rr, vv, dup_first=True)]
            assert got == seq, (trial, world)


def test_shard_range_partitions():
    from theta_b200.search import shard_range
    for total in (0, 1, 7, 142560, 25251156, 169822123456):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def test_filter_records_host_function():
    """theta_filter_records (pure host code in the library): sort by rank, keep NaN / strict prefix minima / near ones."""
    import numpy as np
    from theta_b200 import _lib
    rng = np.random.default_rng(7)
    for trial in range(50):
        n = int(rng.integers(0, 400))
        recs = np.zeros(n, dtype=_lib.CAND_DTYPE)
        recs["rank"] = rng.permutation(n).astype(np.uint64) * 3
        recs["nll"] = 1e9 + rng.choice([0.0, 1e-4, 5e-4, 2e-3, 1.0, 50.0], size=n) * rng.integers(0, 40, size=n)
        if n > 5:
            recs["nll"][rng.integers(0, n, 2)] = np.nan
        g = float(np.nanmin(recs["nll"])) if n and not np.all(np.isnan(recs["nll"])) else float("inf")
        near = g + 2.5e-3
        prior = float(rng.choice([np.inf, g + 10.0]))
        want = []
        run = prior
        for i in np.argsort(recs["rank"], kind="stable"):
            v = float(recs["nll"][i])
            if v != v or v < run or v <= near:
                want.append(int(recs["rank"][i]))
            if v < run:
                run = v
        got = _lib.filter_records(recs.copy(), near, prior)
        assert [int(x) for x in got["rank"]] == want


def test_rank_free_window_fold_equals_sequential_fold():
    """The n=3 search keeps no prefix-minimum history: it replays the reference fold (RunTHetA.py:198-206) over the
    candidates within a window of the global minimum only, after checking that the 1e-3-connected cluster of the
    minimum ends inside the window (theta_b200/search_n3.py).  Randomised streams with exact ties, chains of near
    ties, 1e-3-edge cases and NaN likelihoods (Q6): the windowed fold must equal the sequential one whenever the
    check passes, and widening must terminate."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import theta_oracle as O
    from theta_b200.search import fold_replay
    from theta_b200.search_n3 import cluster_top
    rng = np.random.default_rng(11)
    widened = 0
    for trial in range(4000):
        n = int(rng.integers(1, 120))
        steps = rng.choice([0.0, 1e-3, 9.99e-4, 1.001e-3, 4e-4, 2.0, 50.0], size=n)
        nll = 1e6 + np.cumsum(rng.choice([-1, 1], size=n) * steps)
        if trial % 3 == 0:
            nll[rng.integers(0, n, max(1, n // 10))] = np.nan
        want = O.fold([(i, float(v)) for i, v in enumerate(nll)])
        fin = nll[~np.isnan(nll)]
        if len(fin) == 0:
            assert want == list(range(n))
            continue
        g = float(fin.min())
        window = 4e-3
        while True:
            keep = [i for i in range(n) if np.isnan(nll[i]) or nll[i] <= g + window]
            vals = [float(nll[i]) for i in keep if not np.isnan(nll[i])]
            if cluster_top(vals, g) + 1e-3 < g + window - 1e-6:
                break
            window *= 8
            widened += 1
            assert window < 1e9
        idx = fold_replay(np.array(keep, dtype=np.uint64) + 1, nll[keep], dup_first=False)
        assert [keep[i] for i in idx] == want, (trial, window)
    assert widened > 50  # the widening path is exercised


def test_bench_config_is_one_dict_for_both_arms():
    """bench.py: the `config` object of the JSON line comes from ONE function for our arm and for --impl reference, names the
    workload, the sharding and the L2 policy, and matches the committed bench line of the final tree."""
    import json
    import bench
    P = bench.headline(1)
    cfg = bench.config_of(P, bench.PER_GPU, 1)
    assert cfg["candidates"] == cfg["candidates_per_gpu"] == 25251156 and cfg["n"] == 2 and cfg["m"] == 150 and cfg["k"] == 5
    assert "workload" in cfg and "sharding" in cfg and "L2" in cfg["l2_policy"]
    with open(os.path.join(ROOT, "profiles", "r2_bench_n1_final_noextra.json")) as f:
        line = json.load(f)
    assert line["config"] == cfg
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert key in line, key
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])


def test_two_pass_pipeline_model_equals_sequential_fold():
    """Python model of the whole n=2 device pipeline against the sequential fold of the reference:
    pass 1 reports per-stripe minima that are exact only where it matters (min <= T + 3e-3, T = the best exact NLL a warp has
    seen in some earlier stripes -- modelled adversarially as the minimum over a random subset of the earlier stripes) and
    over-estimates (up to +inf) elsewhere; the stripe scan lists record / near stripes; pass 2 emits, per unit of U
    consecutive ranks, NaN | near-minimum | strict prefix minima w.r.t. the stripe's prior and the exact prefix of the unit;
    every shard filters its records (prior = +inf); the union goes through theta_filter_records (host code of the library)
    and the fold replay."""
    import theta_oracle as O
    from theta_b200 import _lib
    from theta_b200.search import fold_replay, shard_range
    rng = np.random.default_rng(11)
    margin = 1e-3
    for trial in range(150):
        n = int(rng.integers(40, 3000))
        steps = rng.choice([0.0, 3e-4, 6e-4, 1.1e-3, 0.5, 3.0, 40.0], size=n, p=[.25, .2, .15, .1, .1, .1, .1])
        nll = 5e9 + np.cumsum(rng.choice([-1, 1], size=n) * steps)
        if trial % 6 == 0:
            nll[rng.integers(0, n, 2)] = np.nan
        seq = O.fold([(0, float(nll[0]))] + [(i, float(v)) for i, v in enumerate(nll)])
        S = int(rng.choice([16, 64, 100, 256]))  # stripe length
        U = int(rng.choice([4, 32, 64]))         # pass-2 unit
        for world in (1, 3):
            blocks, mins = [], []
            shards = [shard_range(n, rk, world) for rk in range(world)]
            reported = []
            for (b, e) in shards:  # ---- pass 1
                rep = []
                for s0 in range(b, e, S):
                    v = nll[s0:min(e, s0 + S)]
                    true_min = np.nanmin(v) if np.isfinite(v).any() else np.inf
                    earlier = [np.nanmin(nll[t0:min(e, t0 + S)]) for t0 in range(b, s0, S) if np.isfinite(nll[t0:min(e, t0 + S)]).any()]
                    seen = [x for x in earlier if rng.random() < 0.6]
                    T = min(seen) if seen else np.inf
                    if true_min <= T + 3 * margin:
                        rep.append(true_min)
                    else:
                        rep.append(np.inf if rng.random() < 0.5 else true_min + rng.uniform(0, 50))
                reported.append(rep)
                mins.append(min(rep) if rep else np.inf)
            g = min(mins)  # the all-reduce(min)
            assert g == np.nanmin(nll)
            any_nan = bool(np.isnan(nll).any())
            for (b, e), rep in zip(shards, reported):  # ---- stripe scan + pass 2 + device filter, per shard
                raw, P = [], np.inf
                for si, s0 in enumerate(range(b, e, S)):
                    x = rep[si]
                    rec, near = x < P, x <= g + 2.5 * margin
                    if rec or near or any_nan:
                        prior = P if rec else -np.inf
                        s1 = min(e, s0 + S)
                        for u0 in range(s0, s1, U):
                            run = prior
                            for i in range(u0, min(s1, u0 + U)):
                                v = nll[i]
                                if v != v or v <= g + 2.5 * margin or v < run:
                                    raw.append((i, v))
                                if v < run:
                                    run = v
                    P = min(P, x)
                rng.shuffle(raw)  # (emission order is arbitrary)
                recs = np.zeros(len(raw), dtype=_lib.CAND_DTYPE)
                recs["rank"] = [r_[0] for r_ in raw]
                recs["nll"] = [r_[1] for r_ in raw]
                blocks.append(_lib.filter_records(recs, g + 2.5 * margin))  # == records_filter_kernel (tested on the GPU)
            allrecs = _lib.filter_records(np.concatenate(blocks), g + 2.5 * margin) if blocks else np.zeros(0, dtype=_lib.CAND_DTYPE)
            got = [int(allrecs["rank"][i]) for i in fold_replay(allrecs["rank"], allrecs["nll"], dup_first=True)]
            assert got == seq, (trial, world, S, U)
