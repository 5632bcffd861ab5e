"""CPU, world_size 2 (gloo): the multi-GPU host path of theta_b200.search -- shard ranges, the
all-reduce(min) / all-gather exchange and the record gather -- driven by a Python model of the two
device passes, must give every rank the sequential fold's answer."""
import multiprocessing as mp
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ModelPlan(object):
    """Stands in for _lib.Plan on a GPU-less box: same pass1/pass2 contract over a given NLL stream."""

    def __init__(self, nll):
        self.nll = np.asarray(nll, dtype=np.float64)
        self.range = (0, 0)

    def count(self):
        return len(self.nll)

    def pass1(self, b, e, stream=None):
        self.range = (b, e)
        v = self.nll[b:e]
        v = v[~np.isnan(v)]
        return float(v.min()) if len(v) else float("inf")

    def pass2(self, g, prior=float("inf"), stream=None):
        from theta_b200 import _lib
        b, e = self.range
        keep, run = [], prior
        for rk in range(b, e):
            v = self.nll[rk]
            if np.isnan(v):
                continue  # None candidate
            if v < run or v <= g + 2.5e-3:
                keep.append((rk, v))
            if v < run:
                run = v
        out = np.zeros(len(keep), dtype=_lib.CAND_DTYPE)
        for i, (rk, v) in enumerate(keep):
            out[i]["rank"] = rk
            out[i]["nll"] = v
            out[i]["mu"][0] = 0.25
        return out

    def stats(self):
        b, e = self.range
        return {"n_solved": int((~np.isnan(self.nll[b:e])).sum()), "n_rescanned": 0, "ms_pass2": 0.0}


def _worker(rank, world, port, nll, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from theta_b200 import search
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = search.search_n2(ModelPlan(nll))
        q.put((rank, [b[0] for b in res.best], res.min_nll, res.n_solved, len(res.records)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_search_matches_sequential_fold():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import theta_oracle as O
    rng = np.random.default_rng(5)
    n = 5000
    nll = 3e9 + np.cumsum(rng.choice([-1, 1], size=n) * rng.choice([0.0, 4e-4, 9e-4, 2.0, 50.0], size=n))
    nll[rng.integers(0, n, 40)] = np.nan          # None candidates
    nll[3000] = np.nanmin(nll) - 5e-4              # a late candidate inside the tie window of the minimum
    seq = O.fold([(0, float(nll[0]))] * (0 if np.isnan(nll[0]) else 1) +
                 [(i, float(v)) for i, v in enumerate(nll) if not np.isnan(v)])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nll, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, best, g, solved, nrec in outs:
        assert best == seq, "rank %d" % rank
        assert g == float(np.nanmin(nll))
        assert solved == int((~np.isnan(nll)).sum())


class ModelPlanN3(object):
    """Stands in for an n=3 _lib.Plan: items of a few leaves each over a given NLL stream; same prepare /
    item_weights / search_min_n3 / window_n3 / solve_rows / stats contract as the C ABI."""

    def __init__(self, nll, per_item, null_nll, m=5):
        self.nll = np.asarray(nll, dtype=np.float64)
        self.per_item = per_item
        self._null = null_nll
        self.m = m
        self.range = (0, 0)
        self.windows = []

    def prepare(self, leaf_cap=0, max_items=1 << 22):
        self.n_items = (len(self.nll) + self.per_item - 1) // self.per_item
        self.truncated = False
        return self.n_items

    def item_weights(self):
        return np.full(self.n_items, float(self.per_item))

    def solve_rows(self, paths, want_vals=False, mode=0):  # only the all-[tau,0,0] pseudo candidate is asked for
        assert len(paths) == 1 and not np.any(paths) and mode == 1
        return np.full((1, 3), 1.0 / 3), None, np.array([self._null]), np.array([0], dtype=np.int32), None

    def _span(self):
        b, e = self.range
        return b * self.per_item, min(len(self.nll), e * self.per_item)

    def search_min_n3(self, b, e, t_seed=float("inf"), window=0.0, stream=None):
        self.range = (b, e)
        self.windows.append(window)
        lo, hi = self._span()
        v = self.nll[lo:hi]
        v = v[~np.isnan(v)]
        return float(v.min()) if len(v) else float("inf")

    def window_n3(self, g, window=0.0, stream=None):
        from theta_b200 import _lib
        lo, hi = self._span()
        keep = [(k, self.nll[k]) for k in range(lo, hi) if self.nll[k] <= g + window]
        out = np.zeros(len(keep), dtype=_lib.CAND_DTYPE)
        paths = np.zeros((len(keep), self.m), dtype=np.uint8)
        for i, (k, v) in enumerate(keep):
            out[i]["rank"] = (k // self.per_item << 32) | (k % self.per_item)
            out[i]["nll"] = v
            paths[i] = [(k >> (8 * j)) & 0xff for j in range(self.m)]  # the path encodes the leaf index
        return out, paths

    def stats(self):
        lo, hi = self._span()
        return {"n_solved": int((~np.isnan(self.nll[lo:hi])).sum()), "n_candidates": hi - lo, "n_rescanned": 0, "ms_pass2": 0.0}


def _worker_n3(rank, world, port, nll, null_nll, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from theta_b200 import search_n3
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ModelPlanN3(nll, 7, null_nll)
        res = search_n3.search_n3(plan)
        best = [None if b[0] is None else int(sum(int(x) << (8 * j) for j, x in enumerate(b[0]))) for b in res.best]
        q.put((rank, best, res.min_nll, res.count, res.n_solved, len(plan.windows)))
    finally:
        dist.destroy_process_group()


def test_two_rank_n3_search_matches_sequential_fold():
    """n=3 host path on 2 gloo ranks: items sharded by weight, min all-reduce, record + path gather, the check that
    the tie cluster ends inside the window (the stream below chains ties well past the first window, so the search
    has to widen it), and the fold over the stream that starts with the all-[tau,0,0] pseudo candidate (Q2)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import theta_oracle as O
    rng = np.random.default_rng(9)
    n = 3001
    nll = 5e8 + np.cumsum(rng.choice([-1, 1], size=n) * rng.choice([0.0, 3e-4, 8e-4, 1.5, 40.0], size=n))
    nll[rng.integers(0, n, 30)] = np.nan
    null_nll = float(np.nanmin(nll) + 7e-4)  # the pseudo candidate ties with the minimum
    seq = O.fold([(-1, null_nll)] + [(i, float(v)) for i, v in enumerate(nll) if not np.isnan(v)])
    want = [None if i == -1 else i for i in seq]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_n3, args=(r, 2, port, nll, null_nll, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, best, g, count, solved, n_windows in outs:
        assert best == want, "rank %d" % rank
        assert n_windows >= 1
        assert g == min(float(np.nanmin(nll)), null_nll)
        assert count == n + 1 and solved == int((~np.isnan(nll)).sum()) + 1
