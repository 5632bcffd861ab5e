"""Kernel time of each contiguous shard of the weak-scaling workload, run one after the other on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from theta_b200 import _lib, search
_lib.init(0)
P = bench.problem("cfg2w")
plan = _lib.Plan(2, P["r"], P["rN"], P["lower"], P["upper"])
for world in (2, 4, 8):
    total = min(plan.count(), bench.PER_GPU * world)
    ms = []
    for r in range(world):
        b, e = search.shard_range(total, r, world)
        for _ in range(2):
            plan.pass1(b, e)
        ms.append(plan.stats()["ms_pass1"])
    print(world, ["%.3f" % x for x in ms], "max/mean %.3f" % (max(ms) / (sum(ms) / len(ms))))
