"""Per-phase wall time of one sharded search step under torchrun (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from theta_b200 import _lib, search
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
_lib.init(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
P = bench.problem("cfg2w")
plan = _lib.Plan(2, P["r"], P["rN"], P["lower"], P["upper"])
total = min(plan.count(), bench.PER_GPU * world)
b, e = search.shard_range(total, rank, world)
_, _, D = search._dist_state()
T = {}
def tick(k, t0):
    torch.cuda.synchronize()
    T.setdefault(k, []).append(time.perf_counter() - t0)
t = torch.zeros(1, dtype=torch.float64, device="cuda")
for rep in range(40):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); g, lm = search.pass1_allreduce(plan, b, e, D, world, None); tick("pass1+allreduce", t0)
    t0 = time.perf_counter(); recs = plan.pass2(g, float("inf"), None); tick("pass2", t0)
    t0 = time.perf_counter(); allrecs = search.gather_records(recs, D, world); tick("gather_records", t0)
    t0 = time.perf_counter(); dist.all_reduce(t, op=dist.ReduceOp.MIN); tick("bare all_reduce", t0)
    t0 = time.perf_counter(); x = float(t.item()); tick("item()", t0)
    t0 = time.perf_counter(); lm2 = plan.pass1(b, e, None); tick("pass1 alone", t0)
if rank == 0:
    for k, v in T.items():
        v = sorted(v[5:])
        print("%-20s median %.3f ms  min %.3f ms" % (k, 1e3 * v[len(v) // 2], 1e3 * v[0]))
    print("kernel ms", plan.stats()["ms_pass1"])
dist.destroy_process_group()
