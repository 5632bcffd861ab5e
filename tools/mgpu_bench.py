"""ONE process, G GPUs (theta_mgpu_*: host thread + stream per GPU, one ncclAllReduce(min) per search): the headline
workload of bench.py (first G x 25,251,156 ranks of the widened configs[1] space) and BASELINE configs[2] (m=300, k=7,
whole space).  Prints one JSON line.  Run by bench.py's rank 0 in a child process, or by hand:
    python tools/mgpu_bench.py --gpus 8 --steps 20"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from theta_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=0)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--no-cfg3", action="store_true")
a = ap.parse_args()
G = a.gpus or _lib.device_count()
P = bench.headline(G)
mg = _lib.MultiGpu(2, P["r"], P["rN"], P["lower"], P["upper"], n_devices=G)
budget = bench.PER_GPU * G
for _ in range(3):
    out = mg.optimize(budget)
t0 = time.perf_counter()
dev = []
for _ in range(a.steps):
    out = mg.optimize(budget)
    dev.append(float(mg.last_ms().max()))
wall = (time.perf_counter() - t0) / a.steps
res = {"n_gpus": G, "api": "theta_mgpu_optimize (one blocking C call per search)", "candidates": out[6],
       "ms_per_step_wall": wall * 1e3, "ms_per_step_device_max": float(np.mean(dev)),
       "cand_per_s_wall": out[6] / wall, "min_nll": float(out[5]["min_nll"]), "n_best": int(len(out[2]))}
mg.close()
if not a.no_cfg3:
    P3 = bench.problem("cfg3")
    mg3 = _lib.MultiGpu(2, P3["r"], P3["rN"], P3["lower"], P3["upper"], n_devices=G)
    t0 = time.perf_counter()
    o3 = mg3.optimize()
    dt = time.perf_counter() - t0
    res["cfg3_strong"] = {"candidates": o3[6], "seconds": dt, "cand_per_s": o3[6] / dt, "min_nll": float(o3[5]["min_nll"]),
                          "device_ms": [round(float(x), 1) for x in mg3.last_ms()]}
    mg3.close()
print(json.dumps(res))
