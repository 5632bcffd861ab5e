"""Host side of the likelihood search: the reference's do_optimization* interface.

Mirrors python/RunTHetA.py:124-220 (same names, argument meaning, return
structure and error behaviour).  The enumerate -> solve -> keep-best loop runs on
the GPU through the C ABI (theta_b200/_lib.py); this module only
  * shards the rank range over the ranks of an initialised torch.distributed
    group (one process per GPU) and exchanges the per-shard minima with one
    all-reduce(min) plus a tiny all-gather, and
  * replays the reference's order-dependent tie fold (RunTHetA.py:198-206,
    Misc.isClose python/Misc.py:36-47) over the few records pass 2 returns.
"""
import os
import sys

import numpy as np

from . import _lib
from ._lib import Plan, TIE_MARGIN

INF = float("inf")


def is_close(a, b, margin=TIE_MARGIN):
    """Misc.isClose for one pair (python/Misc.py:36-47): `not (|a-b| > margin)`, so NaN is close (Q6)."""
    return not (abs(a - b) > margin)


def fold_replay(ranks, nlls, dup_first=True):
    """The running-min fold of RunTHetA.py:198-206 over records sorted by enumeration rank.

    Returns the indices of the surviving records in fold order (an index can appear
    twice: the drivers evaluate the first matrix twice, RunTHetA.py:148,188 -- Q1).
    """
    best, m_ = [], INF
    ranks = ranks.tolist() if hasattr(ranks, "tolist") else list(ranks)
    nlls = nlls.tolist() if hasattr(nlls, "tolist") else list(nlls)
    for i, nll in enumerate(nlls):
        reps = 2 if (dup_first and ranks[i] == 0) else 1
        for _ in range(reps):
            if not (abs(nll - m_) > TIE_MARGIN):  # Misc.isClose: NaN counts as close (Q6)
                best.append(i)
            elif nll < m_:
                best = [i]
                m_ = nll
    return best


def _dist_state():
    """(rank, world, torch.distributed or None)."""
    try:
        import torch.distributed as dist
    except Exception:  # torch absent: single process
        return 0, 1, None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size(), dist
    return 0, 1, None


def shard_range(total, rank, world):
    """Contiguous rank range of one shard: [total*rank//world, total*(rank+1)//world)."""
    return (total * rank) // world, (total * (rank + 1)) // world


def _dev(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def exchange_minima(local_min, dist, rank, world):
    """One all-reduce(min) for the global best NLL (host value in, host value out; the gloo / test path).

    Returns (global_min, prior_min).  prior_min is always +inf: every shard emits the strict prefix minima of
    its own range, a superset of the global ones, and the fold replay over the union is unchanged by the extra
    records (each is >= an earlier record - margin, so it can neither reset the fold nor survive in the final
    tie list unless it is within two margins of the global minimum, in which case it is emitted anyway).
    """
    if dist is None:
        return local_min, INF
    import torch
    t = torch.tensor([local_min], dtype=torch.float64, device=_dev(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item()), INF


_BUFS = {}


def _bufs(dist, world, nbytes):
    """Per-process exchange buffers (device + pinned host), allocated once."""
    import torch
    dev = _dev(dist)
    key = (str(dev), world, nbytes)
    b = _BUFS.get(key)
    if b is None:
        pin = dev.type == "cuda"
        b = {"min": torch.empty(1, dtype=torch.float64, device=dev),
             "send_h": torch.empty(nbytes, dtype=torch.uint8, pin_memory=pin),
             "send_d": torch.empty(nbytes, dtype=torch.uint8, device=dev),
             "recv_d": torch.empty(world * nbytes, dtype=torch.uint8, device=dev),
             "recv_h": torch.empty(world * nbytes, dtype=torch.uint8, pin_memory=pin)}
        _BUFS[key] = b
    return b


def pass1_allreduce(plan, begin, end, dist, world, stream=None):
    """Pass 1 on this shard + the single all-reduce(min) over all shards (host values; the gloo / test path and
    callers that want the two passes separately).  Returns (global_min, local_min)."""
    lm = plan.pass1(begin, end, stream)
    if dist is None:
        return lm, lm
    return exchange_minima(lm, dist, 0, world)[0], lm


def search_shard(plan, begin, end, dist, world, stream=None):
    """Both passes on this shard with the one all-reduce(min) in between (host values: single GPU, gloo, host models
    of the tests).  Returns (records, global_min)."""
    if not hasattr(plan, "pass1_begin"):  # host model of the plan (tests/test_dist_gloo.py)
        g, _ = pass1_allreduce(plan, begin, end, dist, world, stream)
        return plan.pass2(g, INF, stream), g
    if dist is None:
        plan.pass1_begin(begin, end, None, stream)
        return plan.pass2_dev(None, INF, stream)
    g, _ = pass1_allreduce(plan, begin, end, dist, world, stream)
    return plan.pass2(g, INF, stream), g


_COMM = {}


def comm_for(dist):
    """The process's theta_comm over the ranks of `dist` (NCCL backend), made once: the NCCL id of rank 0 travels
    through torch's own broadcast, everything after that is C (theta_b200/csrc/theta_mgpu.cu)."""
    key = (dist.get_rank(), dist.get_world_size())
    c = _COMM.get(key)
    if c is None:
        c = _lib.Comm.from_torch(dist)
        _COMM[key] = c
    return c


def search_sharded(plan, begin, end, dist, world, stream=None):
    """The whole sharded search of one step.  Returns (records, global_min, n_solved): the records of all shards
    sorted by rank (a superset of what the fold replay needs), identical on every rank.

    On GPUs the step is ONE C call per rank (theta_search_n2_dist): pass 1 -> all-reduce(min) on the device double ->
    stripe scan + pass 2 + device record filter -> all-gather of the [status | records] blocks -> one host wait (both
    collectives as remote stores over NVLink peer memory, NCCL as the fallback; a CUDA graph from the second step).  The host
    models of the tests and the gloo backend take the two-pass Python path below."""
    if dist is not None and hasattr(plan, "handle") and dist.get_backend() == "nccl":
        return comm_for(dist).search_n2(plan, begin, end, stream)
    recs, g = search_shard(plan, begin, end, dist, world, stream)
    allrecs, extras = gather_records(recs, dist, world, extra=[plan.stats()["n_solved"]])
    return allrecs, g, sum(e[0] for e in extras)


_GATHER_CAP = 256  # records per shard carried by the fixed-size exchange
_GATHER_HEAD = 8 * 4  # int64 header: record count + up to three extras


def gather_records(recs, dist, world, extra=None):
    """All shards' pass-2 records, concatenated in rank order (they are already sorted per shard).

    One fixed-size all-gather carries up to _GATHER_CAP records per shard plus the shard's record count
    (and `extra`, up to three int64 such as the solved-candidate count); only when a shard has more
    records is a second, variable-size exchange needed.  Returns records, or (records, extras) if extra is given.
    """
    if dist is None:
        return recs if extra is None else (recs, [list(extra)])
    import torch
    item = _lib.CAND_DTYPE.itemsize
    nx = 0 if extra is None else len(extra)
    assert nx <= 3
    b = _bufs(dist, world, _GATHER_BYTES)
    sh = b["send_h"].numpy()
    head = sh[:_GATHER_HEAD].view(np.int64)
    head[:] = 0
    head[0] = len(recs)
    if nx:
        head[1:1 + nx] = extra
    k = min(len(recs), _GATHER_CAP)
    sh[_GATHER_HEAD:_GATHER_HEAD + k * item] = recs[:k].view(np.uint8).reshape(-1)
    b["send_d"].copy_(b["send_h"], non_blocking=True)
    dist.all_gather_into_tensor(b["recv_d"], b["send_d"])
    b["recv_h"].copy_(b["recv_d"], non_blocking=True)
    if b["recv_d"].is_cuda:
        torch.cuda.current_stream().synchronize()
    raw = b["recv_h"].numpy().reshape(world, _GATHER_BYTES)
    heads = raw[:, :_GATHER_HEAD].copy().view(np.int64)
    ns = [int(h[0]) for h in heads]
    if max(ns) <= _GATHER_CAP:
        parts = [raw[r, _GATHER_HEAD:_GATHER_HEAD + ns[r] * item].copy().view(_lib.CAND_DTYPE) for r in range(world)]
    else:
        dev = _dev(dist)
        cap = max(ns)
        big = np.zeros(cap * item, dtype=np.uint8)
        big[:len(recs) * item] = recs.view(np.uint8).reshape(-1)
        tb = torch.from_numpy(big).to(dev)
        outb = torch.empty(world * big.size, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(outb, tb)
        rawb = outb.cpu().numpy().reshape(world, big.size)
        parts = [rawb[r, :ns[r] * item].copy().view(_lib.CAND_DTYPE) for r in range(world)]
    allrecs = np.concatenate(parts) if parts else recs
    if extra is None:
        return allrecs
    return allrecs, [[int(v) for v in h[1:1 + nx]] for h in heads]


class SearchResult(object):
    def __init__(self):
        self.best = []          # [(rank, mu0, nll, status)] in fold order (may repeat rank 0)
        self.records = None     # all pass-2 records (global)
        self.count = 0          # candidates in the space (without the duplicate)
        self.n_solved = 0
        self.min_nll = INF
        self.stats = {}


def search_n2(plan, stream=None, rank_limit=None):
    """Sharded two-pass search; every rank returns the same SearchResult.

    rank_limit (optional, not in the reference): search only the first `rank_limit` candidates of the
    enumeration order -- the rank-budget runs of SURVEY 8d for spaces too large to finish."""
    rank, world, dist = _dist_state()
    total = plan.count()
    if rank_limit is not None:
        total = min(total, int(rank_limit))
    begin, end = shard_range(total, rank, world)
    allrecs, g, n_solved = search_sharded(plan, begin, end, dist, world, stream)
    stats = plan.stats()
    res = SearchResult()
    res.records = allrecs
    res.count = total
    res.min_nll = g
    res.stats = stats
    res.n_solved = n_solved
    idx = fold_replay(allrecs["rank"], allrecs["nll"], dup_first=True)
    res.best = [(int(allrecs["rank"][i]), float(allrecs["mu"][i][0]), float(allrecs["nll"][i]), int(allrecs["status"][i]))
                for i in idx]
    return res


def _reverse_sort_rows(C, sorted_index):
    """DataTools.reverse_sort_C (python/DataTools.py:132-145)."""
    out = np.zeros(C.shape)
    out[np.asarray(sorted_index, dtype=np.int64)] = C
    return out


def _reverse_sort_list(vec, sorted_index):
    """DataTools.reverse_sort_list (python/DataTools.py:147-159)."""
    out = [0] * len(sorted_index)
    for i, s in enumerate(sorted_index):
        out[s] = vec[i]
    return out


def _finish_n2(plan, res, tau, sorted_index):
    """Winners -> the reference's `best` tuples (C_new, mu, likelihood, vals) in original interval order."""
    best = []
    if not res.best:
        return best
    uniq = sorted(set(b[0] for b in res.best))
    cands, vals = plan.eval_ranks(uniq, want_vals=True)
    where = {rk: i for i, rk in enumerate(uniq)}
    cols = {rk: plan.enumerate(rk, 1)[0] for rk in uniq}
    for rk, mu0, nll, _st in res.best:
        C = np.zeros((plan.m, 2))
        C[:, 0] = tau
        C[:, 1] = cols[rk]
        mu0 = np.float64(mu0)
        v = [np.float64(x) for x in vals[where[rk]]]
        best.append((_reverse_sort_rows(C, sorted_index), (mu0, 1 - mu0), np.float64(nll),
                     _reverse_sort_list(v, sorted_index)))
    return best


_GATHER_BYTES = _GATHER_HEAD + _GATHER_CAP * _lib.CAND_DTYPE.itemsize


_NO_PROFILES = ("Error: No valid Copy Number Profiles exist for these intervals within the bounds specified. "
                "Exiting...")


def do_optimization_single(n, m, k, tau, lower_bounds, upper_bounds, r, rN, max_normal, sorted_index, multi_event,
                           get_values, prefix=None, rank_limit=None):
    """RunTHetA.do_optimization_single (python/RunTHetA.py:173-220) on the GPU.

    Same arguments and return value: a list of (C_new, mu, likelihood, vals) with
    C_new an m x n float64 array in the original interval order.  `k` and
    `multi_event` are accepted and, as in the reference, not used by the search
    (Enumerator.py:55,58 -- Q3/Q4).  With get_values, every candidate's
    (column, mu0, NLL) is written to <prefix>.likelihoods like RunTHetA.py:210-215.
    """
    if n == 3:
        from .search_n3 import do_optimization_n3
        return do_optimization_n3(m, k, tau, lower_bounds, upper_bounds, r, rN, sorted_index, get_values, prefix)
    if n != 2:
        raise ValueError("n must be 2 or 3 (FileIO.py:42)")
    if _dist_state()[2] is None and not get_values:
        # single GPU: the whole search is one C call (theta_optimize_n2: plan, both passes, fold, winners)
        try:
            lo, up, recs, cols, vals, _st, _nc = _lib.optimize_n2(r, rN, lower_bounds, upper_bounds, tau, max_normal, rank_limit)
        except _lib.EmptySpaceError:
            # Enumerator._check_bound_order edits the caller's lists before the reference notices (Q16)
            print(_NO_PROFILES)
            sys.exit(1)
        lower_bounds[:] = lo.tolist()
        upper_bounds[:] = up.tolist()
        best = []
        si = np.asarray(sorted_index, dtype=np.int64)
        for k in range(len(recs)):
            C = np.zeros((m, 2))
            C[si, 0] = tau
            C[si, 1] = cols[k]
            mu0 = np.float64(recs["mu"][k][0])
            v = [0] * m
            for i, x in zip(sorted_index, vals[k].tolist()):
                v[i] = np.float64(x)
            best.append((C, (mu0, 1 - mu0), np.float64(recs["nll"][k]), v))
        return best
    plan = Plan(2, r, rN, lower_bounds, upper_bounds, tau=tau, max_normal=max_normal)
    try:
        # Enumerator._check_bound_order edits the caller's lists in place (Q16)
        lo, up = plan.bounds()
        lower_bounds[:] = lo.tolist()
        upper_bounds[:] = up.tolist()
        try:
            res = search_n2(plan, rank_limit=rank_limit)
        except _lib.EmptySpaceError:
            print(_NO_PROFILES)
            sys.exit(1)
        if get_values:
            dump_values_n2(plan, (prefix or "theta") + ".likelihoods")
        return _finish_n2(plan, res, tau, sorted_index)
    finally:
        plan.close()


def do_optimization(n, m, k, tau, lower_bounds, upper_bounds, r, rN, max_normal, sorted_index, max_processes,
                    multi_event, get_values, prefix=None):
    """RunTHetA.do_optimization (python/RunTHetA.py:124-171).

    The reference farms candidates out to `max_processes - 1` worker processes and merges their tie lists in
    scheduling order (Q13); parity is defined against the single-process fold, which is what the GPU path reproduces
    on any number of GPUs.  `max_processes` maps to the number of GPUs: from a plain process (no torch.distributed
    group) the search runs on min(max_processes, visible GPUs) devices through ONE blocking C call (theta_mgpu_*: one
    host thread + stream per GPU, a single ncclAllReduce(min) between the passes).  Under torchrun every rank drives
    its own GPU and `max_processes` is ignored.
    """
    n_gpus = 1
    if _dist_state()[2] is None and not get_values and max_processes and int(max_processes) > 1:
        n_gpus = max(1, min(int(max_processes), _lib.device_count()))
    if n_gpus <= 1:
        return do_optimization_single(n, m, k, tau, lower_bounds, upper_bounds, r, rN, max_normal, sorted_index,
                                      multi_event, get_values, prefix)
    if n not in (2, 3):
        raise ValueError("n must be 2 or 3 (FileIO.py:42)")
    try:
        mg = _lib.MultiGpu(n, r, rN, lower_bounds, upper_bounds, tau=tau, max_normal=max_normal, n_devices=n_gpus)
    except _lib.EmptySpaceError:
        print(_NO_PROFILES)
        sys.exit(1)
    try:
        try:
            lo, up, recs, rows, vals, _st, _nc = mg.optimize()
        except _lib.EmptySpaceError:
            print(_NO_PROFILES)
            sys.exit(1)
        lower_bounds[:] = lo.tolist()  # Q16
        upper_bounds[:] = up.tolist()
        return best_from_records(n, m, tau, recs, rows, vals, sorted_index, mg)
    finally:
        mg.close()


def best_from_records(n, m, tau, recs, rows, vals, sorted_index, mg=None):
    """theta_mgpu_optimize's output -> the reference's `best` tuples (C_new, mu, likelihood, vals), original interval order."""
    best = []
    si = np.asarray(sorted_index, dtype=np.int64)
    table = None
    for j in range(len(recs)):
        C = np.zeros((m, n))
        C[si, 0] = tau
        v = [0] * m
        for i, x in zip(sorted_index, vals[j].tolist()):
            v[i] = np.float64(x)
        if n == 2:
            C[si, 1] = rows[j].view(np.int8)
            mu0 = np.float64(recs["mu"][j][0])
            best.append((C, (mu0, 1 - mu0), np.float64(recs["nll"][j]), v))
        else:
            if table is None:
                table = n3_row_table(tau, int(max(int(x) for x in mg._up))) if mg is not None else None
            if int(recs["rank"][j]) != 0xFFFFFFFFFFFFFFFF:
                ab = np.array([table[int(g)] for g in rows[j]], dtype=np.float64)
                C[si, 1] = ab[:, 0]
                C[si, 2] = ab[:, 1]
            best.append((C, np.array(recs["mu"][j], dtype=np.float64), np.float64(recs["nll"][j]), v))
    return best


def n3_row_table(tau, K):
    """Enumerator._create_graph's row list (python/Enumerator.py:272-298): first component fastest, rows with
    (tau - a)(tau - b) >= 0 -- the order theta_n3_rows reports."""
    return [(a, b) for b in range(K + 1) for a in range(K + 1) if (tau - a) * (tau - b) >= 0]


def dump_values_n2(plan, path, chunk=1 << 18):
    """--GET_VALUES (python/RunTHetA.py:210-215): one line per *solved* candidate in enumeration
    order, first matrix twice: tumour column string, mu[0], likelihood."""
    total = plan.count()
    with open(path, "w") as f:
        for b in range(0, total, chunk):
            cnt = min(chunk, total - b)
            cols = plan.enumerate(b, cnt)
            mu, nll, st = plan.solve_tiles(cols)
            for i in range(cnt):
                if st[i] >= _lib.ST_NONE_SIGN:
                    continue
                line = "".join(str(int(v)) for v in cols[i]) + "\t" + str(np.float64(mu[i])) + "\t" + str(np.float64(nll[i])) + "\n"
                f.write(line)
                if b + i == 0:
                    f.write(line)
