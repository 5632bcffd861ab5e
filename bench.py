#!/usr/bin/env python
"""bench.py -- candidates/second of THetA's combinatorial likelihood search on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|m150k7|cfg3win]

One "step" = one complete search (enumerate every admissible C, solve mu, NLL, keep the
best with the reference's tie fold) of the workload's candidate space.  At N=1 the workload is
BASELINE.json configs[1]: synthetic m=150, n=2, k=5 (seed 1) -> 25,251,156 candidates.
For N>1 (torchrun, one rank per GPU) the scaling is WEAK: every GPU keeps the same 25,251,156
candidates per step; the job is the first N x 25,251,156 ranks of the same data's search with the
upper bounds widened by two copies (same r, rN, m, K=5 -> same cost per candidate; 4.196e8 candidates),
cut into contiguous rank ranges.  The only communication is the min exchange + the record gather.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      the streaming NLL kernel (theta_nll_tiles_n2) on tiles larger than L2, HBM-bound
  search_kernel the fused search kernel's own device time / fp64 utilisation
  cpu_baseline  the oracle's Python port of the reference path on the host cores (bounded sample)
  extra         the >=100x target shape (m=150, k=7) searched once
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU = 25251156  # candidates of BASELINE configs[1]; the per-GPU unit count of the weak-scaling runs

WORKLOADS = {
    "cfg2": dict(m=150, k=5, seed=1, desc="synthetic m=150 n=2 k=5 seed=1 (BASELINE configs[1]), full space"),
    "cfg2w": dict(m=150, k=5, seed=1, widen=2,
                  desc="synthetic m=150 n=2 k=5 seed=1 (BASELINE configs[1] data), upper bounds +2 copies (4.196e8 "
                       "candidates), first n_gpus x 25,251,156 ranks"),
    "m150k7": dict(m=150, k=7, seed=1, desc="synthetic m=150 n=2 k=7 seed=1 (>=100x target shape), full space"),
    "cfg3": dict(m=300, k=7, seed=1, desc="synthetic m=300 n=2 k=7 seed=1 (BASELINE configs[2]), full space (1.698e11 candidates)"),
    "cfg3win": dict(m=300, k=7, seed=1, window=1 << 33,
                    desc="synthetic m=300 n=2 k=7 seed=1 (BASELINE configs[2]), first 2^33 ranks"),
}


def problem(name):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import theta_oracle as O  # input generator only; not on the measured path
    w = WORKLOADS[name]
    P = O.synthetic_problem_n2(w["m"], w["k"], seed=w["seed"])
    if w.get("widen"):
        P["upper"] = [min(w["k"], u + w["widen"]) for u in P["upper"]]
    P["window"] = w.get("window")
    P["desc"] = w["desc"]
    return P


# ----------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm (oracle port of the reference)
def _cpu_worker(args):
    name, begin, count = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import theta_oracle as O
    P = problem(name)
    col = O.unrank_n2(P["lower"], P["upper"], begin)
    # continue the reference odometer (Enumerator.py:119-152) from `begin`
    lo, up = P["lower"], P["upper"]
    m = len(col)
    t0 = time.perf_counter()
    best = float("inf")
    for _ in range(count):
        s = O.solve_n2(O.column_to_matrix(col), P["r"], P["rN"], 1.0)
        if s is not None and s[1] < best:
            best = s[1]
        for i in range(m - 1):
            if col[i] < col[i + 1] and col[i] < up[i]:
                col[i] += 1
                for j in range(i):
                    col[j] = lo[j]
                break
        else:
            if col[m - 1] < up[m - 1]:
                col[m - 1] += 1
                for j in range(m - 1):
                    col[j] = lo[j]
            else:
                break
    return time.perf_counter() - t0, best


def cpu_reference_rate(name, per_core, cores=None):
    """Oracle port of do_optimization on `cores` worker processes (RunTHetA.py:124-171 farms
    candidates to processes the same way): first per_core*cores candidates in enumeration order."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    jobs = [(name, i * per_core, per_core) for i in range(cores)]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_worker, [(name, 0, 8)] * cores)  # import / warm-up, untimed
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
    total = per_core * cores
    return {"value": total / wall, "unit": "candidates/s", "cores": cores, "kind": "port",
            "sample": "first %d candidates of the workload in enumeration order, %d per worker process, "
                      "oracle/theta_oracle.py (numpy+scipy restatement of Optimizer.solve), %.1f s wall"
                      % (total, per_core, wall)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.workload == "cfg2":
        args.workload = "cfg2w"  # same config as our arm at this N
    per_core = 40000
    vals = []
    for _ in range(args.warmup and 1 or 0):
        cpu_reference_rate(args.workload, 100)
    res = None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = cpu_reference_rate(args.workload, per_core)
        vals.append(res["value"])
    wall = time.perf_counter() - t0
    v = sum(vals) / len(vals)
    res["value"] = v
    P = problem(args.workload)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import theta_oracle as O
    total = O.count_n2(P["lower"], P["upper"])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "cfg2w":
        total = min(total, PER_GPU * world)
    line = {"impl": "reference", "metric": "C-candidates/sec", "value": v, "unit": "candidates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": P["desc"], "n": 2, "m": P["m"], "k": P["k"], "candidates": int(total)},
            "cpu_baseline": res,
            "e2e": {"value": v, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from theta_b200 import _lib, search

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    _lib.init(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream().cuda_stream

    if world > 1 and args.workload == "cfg2":
        args.workload = "cfg2w"  # weak scaling: same data and K, wider bounds, PER_GPU ranks per GPU
    P = problem(args.workload)
    plan = _lib.Plan(2, P["r"], P["rN"], P["lower"], P["upper"])
    total = plan.count()
    if args.workload == "cfg2w":
        total = min(total, PER_GPU * world)
    if P["window"]:
        total = min(total, P["window"])
    begin, end = search.shard_range(total, rank, world)
    _, _, D = search._dist_state()

    launches = [0]

    def step():
        # search kernel, min reduction [, NCCL all-reduce], stripe scan, emit kernel -- one host synchronisation
        recs, g, _ = search.search_sharded(plan, begin, end, D, world, stream)
        launches[0] += 4
        return recs, g

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p1_ms = []
    e0.record()
    for _ in range(args.steps):
        recs, g = step()
        p1_ms.append(plan.stats()["ms_pass1"])
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    st = plan.stats()
    value = total * args.steps / (ms_total * 1e-3)

    # ---- e2e: the public call (host lists in, the reference-shaped Python `best` list out), every step from
    # scratch: one C-ABI call at N=1 (theta_optimize_n2: tables built on the host, staged in pinned memory, one H2D;
    # both passes; D2H of the status block and records; the fold; winners' columns and vals).
    def e2e_step():
        return search.do_optimization_single(2, P["m"], P["k"], 2, list(P["lower"]), list(P["upper"]), P["r"], P["rN"],
                                             1.0, P["sorted_index"], True, False, rank_limit=total)
    for _ in range(max(3, args.warmup)):
        best = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        best = e2e_step()
    barrier()
    dt = time.perf_counter() - t0
    tdt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
    # bytes that cross PCIe per call: H2D = the plan arena (count tables, r, rN, their prefix sums, status block);
    # D2H = the status block + the speculative record window of pass 2 (theta_n2.cu: 1024 x 48 B)
    up256 = lambda v: (v + 255) & ~255
    h2d = (up256(plan.table_bytes()) + 2 * up256(8 * plan.m) + 2 * up256(8 * (plan.pitch + 1)) + up256(4 * (plan.pitch + 1))
           + up256(8 * 32))
    d2h = 8 * 32 + 48 * 1024
    e2e = {"value": total * args.steps / float(tdt.item()), "unit": "candidates/s",
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(tdt.item()) / args.steps,
           "api": "theta_b200.search.do_optimization_single (RunTHetA.py:173 signature), host lists in / best list out",
           "best_nll": float(best[0][2]), "best_mu0": float(best[0][1][0]), "n_best": len(best)}

    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the streaming NLL kernel (rank 0 only; tiles > L2 so every launch streams from HBM)
    roof = None
    extra = {}
    if rank == 0:
        lib = _lib.load()
        cnt = min(plan.count(), 4 * 1024 * 1024)
        pitch = plan.pitch
        tiles = torch.empty(cnt * pitch, dtype=torch.int8, device="cuda")
        mu = torch.empty(cnt, dtype=torch.float64, device="cuda")
        nll = torch.empty(cnt, dtype=torch.float64, device="cuda")
        stt = torch.empty(cnt, dtype=torch.int32, device="cuda")
        _lib.check(lib.theta_enumerate_n2_dev(plan.handle, plan.count() // 3, cnt, tiles.data_ptr(), pitch, stream))
        _lib.check(lib.theta_solve_tiles_n2_dev(plan.handle, tiles.data_ptr(), pitch, cnt, mu.data_ptr(), nll.data_ptr(),
                                                stt.data_ptr(), stream))
        torch.cuda.synchronize()
        nll_ref = nll.clone()
        for _ in range(3):
            _lib.check(lib.theta_nll_tiles_n2_dev(plan.handle, tiles.data_ptr(), pitch, cnt, mu.data_ptr(), nll.data_ptr(), stream))
        torch.cuda.synchronize()
        reps = 10
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            _lib.check(lib.theta_nll_tiles_n2_dev(plan.handle, tiles.data_ptr(), pitch, cnt, mu.data_ptr(), nll.data_ptr(), stream))
            ev[i + 1].record()
        torch.cuda.synchronize()
        t_nll = sum(ev[i].elapsed_time(ev[i + 1]) for i in range(reps)) / reps
        okm = stt < 8
        maxrel = float(((nll[okm] - nll_ref[okm]).abs() / nll_ref[okm].abs()).max().item()) if bool(okm.any()) else 0.0
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["hbm_gbs"])
                peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
        bytes_alg = cnt * (plan.m + 8 + 8)
        ach = bytes_alg / (t_nll * 1e-3) / 1e9
        traffic = None  # dram__bytes_read+write per launch from the committed ncu --set full capture, scaled to this launch
        try:
            with open(os.path.join(ROOT, "profiles", "nll_traffic.json")) as f:
                tj = json.load(f)
            if tj["m"] == plan.m:
                traffic = tj["dram_bytes_per_candidate"] * cnt
        except Exception:
            pass
        roof = {"kernel": "nll_tiles_kernel (theta_nll_tiles_n2_dev: Optimizer.L2 over int8 C tiles, TMA-staged)",
                "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src, "bytes_per_candidate": plan.m + 16, "candidates_per_launch": cnt,
                "ms_per_launch": t_nll, "tile_bytes": cnt * pitch, "check_max_rel_err_vs_solve": maxrel}
        del tiles, mu, nll, stt
        # fused search kernel: device time and fp64 rate
        fp64 = _lib.fp64_peak_gflops(stream)
        extra["fp64_peak_gflops_measured"] = fp64
    # search kernel own numbers (all ranks ran it; report rank 0's)
    search_kernel = {"kernel": "n2_search_kernel<KT,false>", "ms_pass1_mean": sum(p1_ms) / len(p1_ms),
                     "ms_pass2_last": st["ms_pass2"], "candidates_per_launch": int(end - begin),
                     "cand_per_s_kernel": (end - begin) / (sum(p1_ms) / len(p1_ms) * 1e-3),
                     "newton_iters_mean": st["newton_iters"], "stripes": int(st["n_stripes"]),
                     "rescanned_stripes": int(st["n_rescanned"])}

    if not args.no_extra and args.workload == "cfg2" and world == 1:
        # (a) the >=100x target shape of BASELINE.json (m=150, k=7, 6.07e8 candidates), sharded over all ranks
        P7 = problem("m150k7")
        plan7 = _lib.Plan(2, P7["r"], P7["rN"], P7["lower"], P7["upper"])
        n7 = plan7.count()
        b7, e7 = search.shard_range(n7, rank, world)

        def step7():
            r_, g_, _ = search.search_sharded(plan7, b7, e7, D, world, stream)
            return r_, g_
        for _ in range(2):
            step7()
        barrier()
        t0 = time.perf_counter()
        r7, g7 = step7()
        barrier()
        dt7 = time.perf_counter() - t0
        extra["m150_k7_full_space"] = {"candidates": n7, "seconds": dt7, "cand_per_s": n7 / dt7, "min_nll": g7,
                                       "ms_pass1_rank0": plan7.stats()["ms_pass1"], "n_gpus": world}
        plan7.close()
        # (a2) BASELINE configs[2]: m=300, k=7, the whole 1.698e11-candidate space (SURVEY 8d allowed a rank budget;
        # at this rate the full space takes seconds)
        P3f = problem("cfg3")
        plan3f = _lib.Plan(2, P3f["r"], P3f["rN"], P3f["lower"], P3f["upper"])
        n3f = plan3f.count()
        b3f, e3f = search.shard_range(n3f, rank, world)
        barrier()
        t0 = time.perf_counter()
        r3f, g3f, _ = search.search_sharded(plan3f, b3f, e3f, D, world, stream)
        barrier()
        dt3f = time.perf_counter() - t0
        extra["m300_k7_full_space"] = {"candidates": n3f, "seconds": dt3f, "cand_per_s": n3f / dt3f, "min_nll": g3f,
                                       "ms_pass1_rank0": plan3f.stats()["ms_pass1"], "n_records": int(len(r3f)), "n_gpus": world}
        plan3f.close()
        # (b) n = 3: BASELINE configs[3] shape (m=150, k=5, explicit bounds, F=12), whole space, sharded by items
        import theta_oracle as O
        from theta_b200 import search_n3
        P3 = O.synthetic_problem_n3(150, 5, 1, 12, False)
        plan3 = _lib.Plan(3, P3["r"], P3["rN"], P3["lower"], P3["upper"])
        search_n3.search_n3(plan3, leaf_cap=512)
        barrier()
        t0 = time.perf_counter()
        res3 = search_n3.search_n3(plan3, leaf_cap=512)
        barrier()
        dt3 = time.perf_counter() - t0
        extra["n3_m150_k5_full_space"] = {"candidates": res3.count, "seconds": dt3, "cand_per_s": res3.count / dt3,
                                          "min_nll": res3.min_nll, "items": res3.n_items, "n_gpus": world,
                                          "ms_pass1_rank0": res3.stats["ms_pass1"], "hybrj_fevals_mean": res3.stats["newton_iters"]}
        plan3.close()

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        cpu = cpu_reference_rate(args.workload, 40000)

    if rank == 0:
        line = {"metric": "C-candidates/sec", "value": value, "unit": "candidates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": P["desc"], "n": 2, "m": P["m"], "k": P["k"], "candidates": int(total),
                           "sharding": "contiguous rank ranges, %d per GPU; one all-reduce(min) + record gather per step"
                                       % ((total + world - 1) // world),
                           "l2_policy": "search kernel streams nothing from HBM (tables in shared memory); the "
                                        "roofline kernel uses tiles of %d MiB > L2" % (roof["tile_bytes"] >> 20 if roof else 0)},
                "e2e": e2e, "time_to_best_nll_ms": e2e["ms_per_step"], "gpu_launches": launches[0], "clocks": clocks,
                "roofline": roof,
                "search_kernel": search_kernel, "cpu_baseline": cpu, "extra": extra,
                "result": {"min_nll": g, "n_records": int(len(recs)), "n_solved": int(st["n_solved"])}}
        print(json.dumps(line))
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
