"""Multi-GPU check (run under torchrun, one rank per GPU): the sharded search must give every rank exactly
the single-GPU answer.  Prints 'DIST_CHECK OK <world>' on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from theta_b200 import _lib, search, search_n3  # noqa: E402
import theta_oracle as O  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
_lib.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()


def single(fn):
    """Run fn with the process group hidden, i.e. as a 1-GPU search on this rank."""
    saved = search._dist_state
    search._dist_state = lambda: (0, 1, None)
    search_n3._dist_state = search._dist_state
    try:
        return fn()
    finally:
        search._dist_state = saved
        search_n3._dist_state = saved


# n = 2: config #1 (golden) and the m=150/k=5 space
with open(os.path.join(ROOT, "tests/golden/config1_search.json")) as f:
    g = json.load(f)
inp = g["inputs"]
for name, P in (("config1", dict(r=inp["r"], rN=inp["rN"], lower=inp["lower"], upper=inp["upper"])),
                ("m150k5", O.synthetic_problem_n2(150, 5, 1)), ("m60k4", O.synthetic_problem_n2(60, 4, 3))):
    plan = _lib.Plan(2, P["r"], P["rN"], P["lower"], P["upper"])
    multi = search.search_n2(plan)
    one = single(lambda: search.search_n2(plan))
    assert multi.best == one.best, (name, multi.best, one.best)
    assert multi.n_solved == one.n_solved and multi.min_nll == one.min_nll
    if name == "config1":
        assert abs(multi.best[0][2] - g["best"][0]["nll"]) <= 1e-10 * g["best"][0]["nll"]
    plan.close()

# n = 3: a whole small space
import gzip  # noqa: E402
with gzip.open(os.path.join(ROOT, "tests/golden/n3_small_full.json.gz"), "rt") as f:
    c = json.load(f)[2]
plan = _lib.Plan(3, c["r"], c["rN"], c["lower"], c["upper"])
multi = search_n3.search_n3(plan, leaf_cap=32)
one = single(lambda: search_n3.search_n3(plan, leaf_cap=32))
key = lambda res: [(None if b[0] is None else tuple(int(x) for x in b[0]), b[2]) for b in res.best]
assert key(multi) == key(one)
assert multi.count == one.count  # (n_solved counts exactly-solved leaves only: it depends on the screening)
plan.close()
dist.barrier()
if rank == 0:
    print("DIST_CHECK OK", world)
dist.destroy_process_group()
