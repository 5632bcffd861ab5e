"""n = 3 host side: do_optimization* for two tumour populations (python/RunTHetA.py:124-220 with
Enumerator._generate_next_C_3 and Optimizer._solve_n3plus), on the GPU through the C ABI.

Items (DFS sub-trees in enumeration order) are sharded contiguously by estimated weight over the ranks of a
torch.distributed group; every shard finds its exact minimum NLL, one all-reduce(min) gives the global one, every
shard returns its candidates inside the tie window, and the reference's fold is replayed over their union.
"""
import sys

import numpy as np

from . import _lib
from .search import (INF, _dist_state, _reverse_sort_list, _reverse_sort_rows, exchange_minima, fold_replay,
                     gather_records, is_close)

_NO_PROFILES = ("Error: No valid Copy Number Profiles exist for these intervals within the bounds specified. "
                "Exiting...")


def shard_items(weights, rank, world):
    """Contiguous item ranges of (about) equal estimated work."""
    n = len(weights)
    if world == 1:
        return 0, n
    cum = np.concatenate([[0.0], np.cumsum(weights)])
    total = cum[-1]
    lo = int(np.searchsorted(cum, total * rank / world, side="left"))
    hi = int(np.searchsorted(cum, total * (rank + 1) / world, side="left")) if rank + 1 < world else n
    return min(lo, n), min(max(hi, lo), n)


def gather_paths(paths, dist, world):
    if dist is None:
        return paths
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    m = paths.shape[1]
    n = torch.tensor([paths.shape[0]], dtype=torch.int64, device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    ns = [int(x.item()) for x in ns]
    cap = max(max(ns), 1)
    buf = np.zeros((cap, m), dtype=np.uint8)
    buf[:paths.shape[0]] = paths
    t = torch.from_numpy(buf).to(dev)
    outs = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[:k] for o, k in zip(outs, ns)])


class SearchResultN3(object):
    pass


def cluster_top(values, g):
    """Top of the 1e-3-connected cluster of NLL values that contains the global minimum g: walk the sorted values
    upwards from g and stop at the first gap that Misc.isClose would call "not close"."""
    h = g
    for v in sorted(float(x) for x in values if x == x):
        if v < g:
            continue
        if abs(v - h) > _lib.TIE_MARGIN:
            break
        h = v
    return h


def search_n3(plan, leaf_cap=0, stream=None, window=0.0, max_items=None):
    """Sharded n=3 search.  Returns a SearchResultN3 with `best` = list of
    (path or None for the bogus first matrix, mu[3], nll, status) in fold order.

    The fold of RunTHetA.py:198-206 is decided by the candidates inside the 1e-3-connected cluster of NLL values
    around the global minimum g, in enumeration order (any reset outside the cluster is undone by the first cluster
    member that follows).  So: every shard finds its exact minimum (theta_search_n3_min), one
    all-reduce(min) gives g, every shard returns its candidates within `window` of g, and the fold is replayed over
    their union -- after checking that the cluster really ends inside the window; if not, search again with a
    wider one.

    max_items (not in the reference): an explicit work-item budget -- search only the lexicographic prefix of the space
    that fits (the rank-budget runs of SURVEY 8d).  Without it a space that exceeds the item limit is an error, never a
    silent prefix."""
    rank, world, dist = _dist_state()
    n_items = plan.prepare(leaf_cap, max_items) if max_items else plan.prepare(leaf_cap)
    if getattr(plan, "truncated", False) and not max_items:
        raise _lib.ThetaError("the n=3 candidate space exceeds %d work items: theta_n3_prepare would cover only a "
                              "lexicographic prefix of it (pass an explicit item budget to search a prefix)" % n_items)
    w = plan.item_weights()
    b, e = shard_items(w, rank, world)
    # the reference's first "candidate" is the all-[tau,0,0] matrix built from iter = [0]*m, bounds ignored
    # (RunTHetA.py:148,188, Enumerator.py:54,154-160 -- Q2); Optimizer.solve returns garbage mu and (usually) the
    # all-normal likelihood for it: row index 0 is (0, 0), and the reference-order solver reproduces both
    bmu, _, bnll, bst, _ = plan.solve_rows(np.zeros((1, plan.m), dtype=np.uint8), mode=1)
    has_bogus = int(bst[0]) < _lib.ST_NONE_SIGN
    null_nll = float(bnll[0]) if has_bogus else INF
    window = float(window) if window > 0 else 4.0 * _lib.TIE_MARGIN
    use_comm = dist is not None and hasattr(plan, "handle") and dist.get_backend() == "nccl"
    while True:
        if use_comm:  # one C call per rank: minimum, ncclAllReduce(min), window records, ncclAllGather
            from .search import comm_for
            allrecs, allpaths, g_all, tot = comm_for(dist).search_n3(plan, b, e, window, null_nll, stream)
            st = plan.stats()
            extras = [list(tot)]
        else:
            local_min = plan.search_min_n3(b, e, INF, window, stream)
            st = plan.stats()
            g, _ = exchange_minima(local_min, dist, rank, world)
            g_all = min(g, null_nll) if null_nll == null_nll else g
            recs, paths = plan.window_n3(g_all, window, stream)
            allrecs, extras = gather_records(recs, dist, world, extra=[st["n_solved"], st["n_candidates"]])
            allpaths = gather_paths(paths, dist, world)
        vals = list(allrecs["nll"]) + ([null_nll] if has_bogus else [])
        fin = [v for v in vals if v == v]
        if not fin:
            break
        g_ref = min(fin)  # the reference-order values of the records (the search minimum is the grouped value)
        if cluster_top(fin, g_ref) + _lib.TIE_MARGIN < g_all + window - 1e-6:
            break
        window *= 8.0  # the tie cluster reaches the edge of the window: look further
    n_solved = sum(x[0] for x in extras)
    n_cand = sum(x[1] for x in extras)
    # the stream the reference folds: the all-[tau,0,0] matrix first, then the candidates in enumeration order
    nb = 1 if has_bogus else 0
    keys = np.concatenate([[0] * nb, allrecs["rank"].astype(np.uint64) + np.uint64(1)]).astype(np.uint64)
    nlls = np.concatenate([[null_nll] * nb, allrecs["nll"]])
    idx = fold_replay(keys, nlls, dup_first=False)
    res = SearchResultN3()
    res.best = []
    for i in idx:
        if i < nb:
            res.best.append((None, tuple(float(x) for x in bmu[0]), null_nll, int(bst[0])))
        else:
            r = allrecs[i - nb]
            res.best.append((allpaths[i - nb].copy(), tuple(float(x) for x in r["mu"]), float(r["nll"]), int(r["status"])))
    res.records, res.paths = allrecs, allpaths
    res.count = n_cand + 1
    res.n_solved = n_solved + 1
    res.min_nll = g_all
    res.window = window
    res.stats = st
    res.n_items = n_items
    return res


def _matrix(path, rows, tau, m):
    C = np.zeros((m, 3))
    C[:, 0] = tau
    if path is not None:
        for i, g in enumerate(path):
            C[i, 1], C[i, 2] = rows[int(g)]
    return C


def do_optimization_n3(m, k, tau, lower_bounds, upper_bounds, r, rN, sorted_index, get_values=False, prefix=None):
    """The n=3 branch of do_optimization_single (python/RunTHetA.py:173-220)."""
    try:
        plan = _lib.Plan(3, r, rN, lower_bounds, upper_bounds, tau=tau)
    except _lib.EmptySpaceError:
        print(_NO_PROFILES)
        sys.exit(1)
    try:
        lo, up = plan.bounds()
        for i in range(len(lower_bounds)):  # Q16
            lower_bounds[i] = int(lo[i])
            upper_bounds[i] = int(up[i])
        rows = plan.rows()
        res = search_n3(plan)
        best = []
        # Optimizer.solve's third return value for the winners: L3's numer_i / denom in reference order (mode 1); the
        # all-[tau,0,0] pseudo candidate is row 0 everywhere
        wp = [np.zeros(m, dtype=np.uint8) if b[0] is None else np.asarray(b[0], dtype=np.uint8) for b in res.best]
        vals = plan.solve_rows(np.stack(wp), want_vals=True, mode=1)[4] if wp else None
        for j, (path, mu, nll, st) in enumerate(res.best):
            C = _matrix(path, rows, tau, m)
            v = [np.float64(x) for x in vals[j]]
            best.append((_reverse_sort_rows(C, sorted_index), np.array(mu, dtype=np.float64), np.float64(nll),
                         _reverse_sort_list(v, sorted_index)))
        if get_values:
            dump_values_n3(plan, rows, (prefix or "theta") + ".likelihoods")
        return best
    finally:
        plan.close()


def _item_windows(plan, n_items, leaves_per_window):
    """Consecutive item ranges [b, e) holding about `leaves_per_window` candidates each (by the DP weights of the items),
    in enumeration order: callers walk the space in bounded pieces instead of materialising it."""
    w = plan.item_weights()
    b = 0
    while b < n_items:
        e, acc = b, 0.0
        while e < n_items and (e == b or acc + float(w[e]) <= leaves_per_window):
            acc += float(w[e])
            e += 1
        yield b, e
        b = e


def dump_values_n3(plan, rows, path, leaves_per_window=1 << 18):
    """--GET_VALUES for n=3 (RunTHetA.py:188-215): one line per solved candidate of the stream the driver folds -- the
    all-[tau,0,0] first matrix (Q2), then the candidates in enumeration order -- with column 1 of C, mu[0], likelihood.
    The space is walked in windows of items, so the dump of a large space needs memory for one window only."""
    n_items = plan.n_items if getattr(plan, "n_items", 0) else plan.prepare(1 << 12)

    def lines(paths):
        mu, _, nll, st, _ = plan.solve_rows(paths, mode=1)
        for i in range(len(paths)):
            if st[i] >= _lib.ST_NONE_SIGN:
                continue
            yield ("".join(str(rows[int(g)][0]) for g in paths[i]) + "\t" + str(np.float64(mu[i, 0])) + "\t" +
                   str(np.float64(nll[i])) + "\n")

    with open(path, "w") as f:
        f.writelines(lines(np.zeros((1, plan.m), dtype=np.uint8)))
        for b, e in _item_windows(plan, n_items, leaves_per_window):
            paths = plan.enumerate_n3(b, e)
            if len(paths):
                f.writelines(lines(paths))


def enumerator_next_n3(enum):
    """Enumerator.generate_next_C for n=3 (python/Enumerator.py:74-87,172-223): candidates are produced on the GPU one
    window of work items at a time (Enumerator.BLOCK candidates or so), never the whole space."""
    if getattr(enum, "_rows3", None) is None:
        enum._n_items3 = enum._plan.prepare(1 << 12)
        enum._windows3 = _item_windows(enum._plan, enum._n_items3, enum.BLOCK)
        enum._rows3 = enum._plan.rows()
        enum._paths3 = np.zeros((0, enum.m), dtype=np.uint8)
        enum._next = 0
    while enum._next >= len(enum._paths3):
        try:
            b, e = next(enum._windows3)
        except StopIteration:
            return False
        enum._paths3 = enum._plan.enumerate_n3(b, e)
        enum._next = 0
    C = _matrix(enum._paths3[enum._next], enum._rows3, enum.tau, enum.m)
    enum._next += 1
    return C


def optimizer_solve_n3(opt, C):
    """Optimizer._solve_n3plus (python/Optimizer.py:128-165) for one explicit matrix."""
    if opt._plan is None:
        opt._plan = _lib.Plan(3, opt.r, opt.rN, [0] * opt.m, [7] * opt.m, tau=opt.tau)
        opt._rowidx = {rc: i for i, rc in enumerate(opt._plan.rows())}
    try:
        path = np.array([[opt._rowidx[(int(a), int(b))] for a, b in C[:, 1:]]], dtype=np.uint8)
    except KeyError:
        raise ValueError("row outside the admissible set (Enumerator._is_valid_row)")
    mu, _, nll, st, vals = opt._plan.solve_rows(path, want_vals=True)
    if st[0] >= _lib.ST_NONE_SIGN:
        return None
    return (np.array(mu[0]), np.float64(nll[0]), [np.float64(v) for v in vals[0]])
