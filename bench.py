#!/usr/bin/env python
"""bench.py -- candidates/second of THetA's combinatorial likelihood search on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|m150k7|cfg3win]

One "step" = one complete search (enumerate every admissible C, solve mu, NLL, keep the
best with the reference's tie fold) of the workload's candidate space.  At N=1 the workload is
BASELINE.json configs[1]: synthetic m=150, n=2, k=5 (seed 1) -> 25,251,156 candidates.
The scaling series is WEAK and uses ONE workload family at every N (including N=1): the data of configs[1] with the
upper bounds widened by two copies (same r, rN, m, K=5 -> same cost per candidate; 4.196e8 candidates), first
N x 25,251,156 ranks, cut into contiguous rank ranges -- 25,251,156 candidates (configs[1]'s count) per GPU per
step.  The exact configs[1] space, the >=100x target shape and configs[2..4] are measured at every N too (`extra`,
strong scaling).  The only communication is one all-reduce(min) of a double + one all-gather of the filtered record blocks per step
(remote stores over NVLink peer memory in two kernels of ours; ncclAllReduce / ncclAllGather as the fallback).

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline         the streaming NLL kernel (theta_nll_tiles_n2) on tiles larger than L2, HBM-bound
  roofline_search  the kernel that binds the timed step (n2_search_kernel): fp64 flops executed / DFMA peak
  dist_check       N > 1: the sharded answers equal the unsharded ones (rank 0 re-runs them on one GPU)
  cpu_baseline     the reference's own multiprocessing path on the host cores (oracle/_ref snapshot; else the port)
  extra            configs[1] exact, the >=100x target shape (m=150, k=7), configs[2], [3], [4] at this N; the
                   single-process multi-GPU driver (theta_mgpu_*) on the same GPUs
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


def reference_available():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refshim
    return os.path.isdir(refshim.REF), refshim.REF


def live_reference_rate(name, seconds=15.0, cores=None):
    """The reference's OWN multiprocessing path -- RunTHetA.do_optimization (python/RunTHetA.py:124-171: one producer,
    NUM_PROCESSES-1 forked workers, pickled matrices through a Queue), unmodified arithmetic, imported through
    oracle/refshim.py from /root/reference/python or the oracle/_ref snapshot -- on the first P candidates of the
    workload in enumeration order (the Enumerator it instantiates is wrapped to stop after P; nothing else is touched).
    P is sized from a short calibration run so that the timed run takes about `seconds`."""
    import contextlib
    import io
    import warnings
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refshim
    warnings.simplefilter("ignore")
    R = refshim.modules()
    P = problem(name)
    cores = cores or os.cpu_count() or 1
    base = R.Enumerator.Enumerator

    def run(limit):
        class Bounded(base):
            def __init__(self, *a, **k):
                base.__init__(self, *a, **k)
                self._left = limit

            def generate_next_C(self):
                if self._left <= 0:
                    return False
                self._left -= 1
                return base.generate_next_C(self)
        R.RunTHetA.Enumerator = Bounded
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            best = R.RunTHetA.do_optimization(2, P["m"], P["k"], 2, list(P["lower"]), list(P["upper"]), P["r"], P["rN"], 1.0,
                                              P["sorted_index"], cores, True, False)
        dt = time.perf_counter() - t0
        R.RunTHetA.Enumerator = base
        return (limit + 1) / dt, dt, float(best[0][2])

    rate0, _, _ = run(40 * cores)
    limit = max(200, int(rate0 * seconds))
    rate, dt, nll = run(limit)
    return {"value": rate, "unit": "candidates/s", "cores": cores, "kind": "reference",
            "sample": "first %d candidates of the workload in enumeration order through the reference's own "
                      "RunTHetA.do_optimization with NUM_PROCESSES=%d (%d workers + producer), %.1f s wall; modules from %s"
                      % (limit + 1, cores, cores - 1, dt, refshim.REF), "best_nll_of_sample": nll}


def cpu_arm(name, seconds=15.0):
    """cpu_baseline: the live reference when its modules are present (build container: /root/reference; GPU box: the
    oracle/_ref snapshot), else the oracle port -- and says which."""
    ok, _ = reference_available()
    if ok and os.environ.get("THETA_CPU_ARM", "reference") != "port":
        try:
            return live_reference_rate(name, seconds)
        except Exception as exc:  # fall back to the port, keep the reason
            res = cpu_reference_rate(name, 40000)
            res["sample"] += " [live reference failed: %s]" % (str(exc)[:120],)
            return res
    return cpu_reference_rate(name, 40000)


def headline(world):
    """The workload of `value` at every N: BASELINE configs[1]'s data (m=150, n=2, k=5, seed 1), upper bounds widened by
    two copies, the first world x 25,251,156 ranks -- 25,251,156 candidates (configs[1]'s count) per GPU per step."""
    P = problem("cfg2w")
    P["window"] = PER_GPU * world
    return P


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = "cfg2w" if args.workload == "cfg2" else args.workload
    vals = []
    res = None
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        res = cpu_arm(name, seconds=12.0)
        vals.append(res["value"])
    wall = time.perf_counter() - t0
    v = sum(vals) / len(vals)
    res["value"] = v
    P = headline(world) if name == "cfg2w" else problem(name)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import theta_oracle as O
    total = O.count_n2(P["lower"], P["upper"])
    if P.get("window"):
        total = min(total, P["window"])
    line = {"impl": "reference", "metric": "C-candidates/sec", "value": v, "unit": "candidates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(P, total, world),
            "cpu_baseline": res,
            "e2e": {"value": v, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def config_of(P, total, world):
    """The `config` object of the JSON line -- the SAME dict in both arms (ours and --impl reference), so that the driver's
    same-config check compares like with like.  `sharding` and `l2_policy` describe how our arm runs this workload."""
    pitch = 8
    while pitch < P["m"] or (pitch // 8) % 2 == 0:  # theta_tile_pitch: smallest 8 * odd >= m
        pitch += 8
    tile_mib = (min(int(total), 4 * 1024 * 1024) * pitch) >> 20
    return {"workload": P["desc"], "n": 2, "m": P["m"], "k": P["k"], "candidates": int(total),
            "candidates_per_gpu": PER_GPU,
            "sharding": "contiguous rank ranges, %d per GPU; one all-reduce(min) of a double + one all-gather of the "
                        "device-filtered record blocks per step (NVLink peer-memory kernels, NCCL fallback)"
                        % ((int(total) + world - 1) // world),
            "l2_policy": "search kernel streams nothing from HBM (tables in shared memory); the roofline kernel uses "
                         "tiles of %d MiB > L2" % tile_mib}


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
    from theta_b200 import _lib, search, search_n3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    _lib.init(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream().cuda_stream
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import theta_oracle as O  # input generators only

    P = headline(world) if args.workload == "cfg2" else problem(args.workload)
    plan = _lib.Plan(2, P["r"], P["rN"], P["lower"], P["upper"])
    total = plan.count()
    if P["window"]:
        total = min(total, P["window"])
    begin, end = search.shard_range(total, rank, world)
    _, _, D = search._dist_state()

    launches = [0]

    def step():
        # per rank: search kernel, min reduction [, peer all-reduce(min)], stripe scan, emit kernel, record filter [, peer all-gather],
        # one host synchronisation; at N > 1 all of it inside ONE C call (theta_search_n2_dist)
        recs, g, _ = search.search_sharded(plan, begin, end, D, world, stream)
        launches[0] += 5 if world == 1 else 7  # search, min reduction, stripe scan, pass 2, record filter [+ peer all-reduce, peer all-gather]
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
    # both passes; D2H of the status block and records; the fold; winners' columns and vals); at N>1 every rank builds
    # its plan and runs the one-call sharded step.
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
    # bytes that cross PCIe per call and rank: H2D = the plan arena (count tables, r, rN, their prefix sums, status
    # block); D2H = the status block + the device-filtered records (256 B + 256 x 48 B per shard; N>1: the gathered blocks of all shards)
    up256 = lambda v: (v + 255) & ~255
    h2d = (up256(plan.table_bytes()) + 2 * up256(8 * plan.m) + 2 * up256(8 * (plan.pitch + 1)) + up256(4 * (plan.pitch + 1))
           + up256(8 * 32))
    d2h = 8 * 32 + 48 * 256 if world == 1 else world * (8 * 32 + 48 * 256)
    e2e = {"value": total * args.steps / float(tdt.item()), "unit": "candidates/s",
           "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
           "ms_per_step": 1e3 * float(tdt.item()) / args.steps,
           "api": "theta_b200.search.do_optimization_single (RunTHetA.py:173 signature), host lists in / best list out",
           "best_nll": float(best[0][2]), "best_mu0": float(best[0][1][0]), "n_best": len(best)}

    clocks = sampler.stop() if rank == 0 else None

    # ---- dist_check: the N-shard answer == the unsharded answer on the same window (rank 0 re-runs it on its GPU)
    dist_check = None
    sharded_best = search.fold_replay(recs["rank"], recs["nll"], dup_first=True)
    sharded_key = [(int(recs["rank"][i]), float(recs["mu"][i][0]), float(recs["nll"][i])) for i in sharded_best]
    if world > 1:
        ok = True
        if rank == 0:
            plan.pass1_begin(0, total, None, stream)
            r1, g1 = plan.pass2_dev(None, float("inf"), stream)
            one = search.fold_replay(r1["rank"], r1["nll"], dup_first=True)
            one_key = [(int(r1["rank"][i]), float(r1["mu"][i][0]), float(r1["nll"][i])) for i in one]
            ok = (one_key == sharded_key) and (g1 == g)
            dist_check = {"ok": bool(ok), "what": "n=2 headline window: best list (rank, mu, NLL) and minimum of the %d-shard "
                                                  "search identical to the unsharded search on one GPU" % world,
                          "n_best": len(one_key)}
        barrier()

    # ---- roofline of the streaming NLL kernel (rank 0 only; tiles > L2 so every launch streams from HBM)
    roof = None
    roof_search = None
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
                "ms_per_launch": t_nll, "tile_bytes": cnt * pitch, "check_max_rel_err_vs_solve": maxrel,
                "note": "the north-star's NLL kernel; the kernel that binds the timed step is in roofline_search"}
        del tiles, mu, nll, stt
        # ---- the kernel that binds the timed step: n2_search_kernel.  It streams nothing from HBM and (since round 2)
        # runs one fp64 and one fp32 evaluation of the root-finder per candidate, so what binds it is INSTRUCTION ISSUE:
        # achieved = warp instructions per launch (ncu smsp__inst_executed.sum of the committed capture of this very
        # window, profiles/search_flops.json) / the event time measured here; peak = SMs x 4 schedulers x 1 instruction
        # per cycle at the SM clock sampled during the run.  The fp64 view (2 DFMA + DMUL + DADD executed per candidate
        # against the DFMA peak measured by theta_fp64_peak in this process) is kept next to it.
        fp64 = _lib.fp64_peak_gflops(stream)
        extra["fp64_peak_gflops_measured"] = fp64
        try:
            with open(os.path.join(ROOT, "profiles", "search_flops.json")) as f:
                sf = json.load(f)
            fl = float(sf["flops_per_candidate"])
            t_k = sum(p1_ms) / len(p1_ms) * 1e-3
            n_k = int(end - begin)
            ach_f = fl * n_k / t_k / 1e9
            sms = torch.cuda.get_device_properties(local).multi_processor_count
            mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
            issue_peak = sms * 4 * float(mhz) * 1e6 / 1e9                      # G warp-instructions / s
            winst = float(sf["warp_inst_total"]) * n_k / float(sf["candidates"])  # warp instructions of one launch
            ach_i = winst / t_k / 1e9
            roof_search = {"kernel": "n2_search_kernel<KT,false> (pass 1 of the timed step: %.0f %% of the step)"
                                     % (100.0 * (sum(p1_ms) / len(p1_ms)) / (ms_total / args.steps)),
                           "bound": "issue", "achieved": ach_i, "peak": issue_peak, "unit": "G warp-inst/s",
                           "frac": ach_i / issue_peak,
                           "warp_inst_per_candidate": winst / n_k, "inst_source": "ncu smsp__inst_executed.sum of one launch on "
                           "this window (profiles/r2_search_flops.csv); ncu smsp__issue_active = %s %% of active cycles"
                           % sf.get("issue_active_pct"),
                           "peak_source": "%d SMs x 4 schedulers x %.0f MHz (SM clock sampled during the run)" % (sms, float(mhz)),
                           "fp64": {"achieved": ach_f, "peak": fp64, "unit": "GFLOP/s", "frac": ach_f / fp64,
                                    "flops_per_candidate": fl, "fp32_flops_per_candidate": sf.get("fp32_flops_per_candidate"),
                                    "flops_source": sf.get("source"),
                                    "ncu_sm__pipe_fp64_cycles_active_pct": sf.get("pipe_fp64_cycles_active_pct"),
                                    "peak_source": "theta_fp64_peak: dependent-free DFMA streams, measured in this process"},
                           "hbm_bytes_per_launch": 63488, "candidates_per_launch": n_k, "ms_per_launch": t_k * 1e3}
        except Exception:
            roof_search = None
    # search kernel own numbers (all ranks ran it; report rank 0's)
    search_kernel = {"kernel": "n2_search_kernel<KT,false>", "ms_pass1_mean": sum(p1_ms) / len(p1_ms),
                     "ms_pass2_last": st["ms_pass2"], "candidates_per_launch": int(end - begin),
                     "cand_per_s_kernel": (end - begin) / (sum(p1_ms) / len(p1_ms) * 1e-3),
                     "newton_iters_mean": st["newton_iters"], "stripes": int(st["n_stripes"]),
                     "rescanned_stripes": int(st["n_rescanned"])}

    if not args.no_extra and args.workload == "cfg2":
        def timed(fn, warm=1):
            for _ in range(warm):
                fn()
            barrier()
            t0_ = time.perf_counter()
            out = fn()
            barrier()
            d = torch.tensor([time.perf_counter() - t0_], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(d, op=dist.ReduceOp.MAX)
            return out, float(d.item())

        def n2_strong(name):
            Pn = problem(name)
            pl = _lib.Plan(2, Pn["r"], Pn["rN"], Pn["lower"], Pn["upper"])
            n_ = pl.count()
            b_, e_ = search.shard_range(n_, rank, world)
            (r_, g_, _s), dt_ = timed(lambda: search.search_sharded(pl, b_, e_, D, world, stream), warm=0 if n_ > 1e10 else 2)
            out = {"workload": Pn["desc"], "candidates": n_, "seconds": dt_, "cand_per_s": n_ / dt_, "min_nll": g_,
                   "ms_pass1_rank0": pl.stats()["ms_pass1"], "n_records": int(len(r_)), "n_gpus": world, "scaling": "strong"}
            pl.close()
            return out
        # BASELINE configs[1] exactly (the 25,251,156-candidate space of the unwidened bounds), the >=100x target shape
        # (m=150, k=7, 6.07e8 candidates) and configs[2] (m=300, k=7: the WHOLE 1.698e11-candidate space), each cut into
        # `world` contiguous rank ranges: strong scaling
        extra["cfg2_exact_space"] = n2_strong("cfg2")
        extra["m150_k7_full_space"] = n2_strong("m150k7")
        extra["cfg3_strong"] = n2_strong("cfg3")

        # n = 3: BASELINE configs[3] (m=150, k=5, explicit bounds, F=12: whole space) and configs[4] (m=500, k=7,
        # TAU-style bounds: a 200,000-item prefix of the enumeration order), sharded by item weight, one
        # ncclAllReduce(min) between the minimum search and the window records
        def n3_case(m_, k_, F_, tb_, budget):
            P3 = O.synthetic_problem_n3(m_, k_, 1, F_, tb_)
            pl = _lib.Plan(3, P3["r"], P3["rN"], P3["lower"], P3["upper"])
            tp = time.perf_counter()
            pl.prepare(4096 if budget else 0, budget or (1 << 22))  # host: cut the DFS forest into work items (once per plan)
            tp = time.perf_counter() - tp
            res3, dt3 = timed(lambda: search_n3.search_n3(pl, leaf_cap=4096 if budget else 0, max_items=budget or None), warm=1)
            out = {"candidates": res3.count, "seconds": dt3, "prepare_seconds": tp, "cand_per_s": res3.count / dt3, "min_nll": res3.min_nll,
                   "items": res3.n_items, "n_gpus": world, "ms_search_rank0": res3.stats["ms_pass1"],
                   "survivors_solved_rank0": int(res3.stats["n_exact"]), "n_best": len(res3.best), "scaling": "strong",
                   "truncated_to_item_budget": bool(getattr(pl, "truncated", False))}
            key3 = [(None if b[0] is None else tuple(int(x) for x in b[0]), b[2]) for b in res3.best]
            if world > 1 and rank == 0:  # G == 1 check on this space too
                saved = search_n3._dist_state
                search_n3._dist_state = lambda: (0, 1, None)
                try:
                    one3 = search_n3.search_n3(pl, leaf_cap=4096 if budget else 0, max_items=budget or None)
                finally:
                    search_n3._dist_state = saved
                out["equals_single_gpu"] = [(None if b[0] is None else tuple(int(x) for x in b[0]), b[2]) for b in one3.best] == key3
            barrier()
            pl.close()
            return out
        extra["cfg4_n3_m150_k5"] = n3_case(150, 5, 12, False, 0)
        extra["cfg5_n3_m500_k7_tau_bounds"] = n3_case(500, 7, 12, True, 200000)
        if dist_check is not None:
            dist_check["n3_cfg4_equals_single_gpu"] = extra["cfg4_n3_m150_k5"].get("equals_single_gpu")
            dist_check["n3_cfg5_equals_single_gpu"] = extra["cfg5_n3_m500_k7_tau_bounds"].get("equals_single_gpu")
            dist_check["ok"] = bool(dist_check["ok"] and dist_check["n3_cfg4_equals_single_gpu"]
                                    and dist_check["n3_cfg5_equals_single_gpu"])

        # ONE process driving all `world` GPUs (theta_mgpu_*: what do_optimization(max_processes) maps to), measured by
        # rank 0 in a child process while the other ranks wait: the headline workload and configs[2]
        # (the other ranks wait on the rendezvous store, i.e. on the CPU: an NCCL barrier would keep a spinning kernel on
        # their GPUs and time-slice them with the child process)
        torch.cuda.synchronize()
        store = dist.distributed_c10d._get_default_store() if world > 1 else None
        if rank == 0:
            try:
                out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mgpu_bench.py"), "--gpus", str(world),
                                      "--steps", str(args.steps)], capture_output=True, text=True, timeout=300)
                extra["single_process_driver"] = json.loads(out.stdout.strip().split("\n")[-1])
            except Exception as exc:
                extra["single_process_driver"] = {"error": str(exc)[:200]}
            if store is not None:
                store.set("theta_mgpu_done", "1")
        elif store is not None:
            store.wait(["theta_mgpu_done"])
        barrier()

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        cpu = cpu_arm("cfg2w" if args.workload == "cfg2" else args.workload)

    if rank == 0:
        cfg = config_of(P, total, world)
        line = {"metric": "C-candidates/sec", "value": value, "unit": "candidates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg,
                "e2e": e2e, "time_to_best_nll_ms": e2e["ms_per_step"], "gpu_launches": launches[0], "clocks": clocks,
                "roofline": roof, "roofline_search": roof_search, "dist_check": dist_check,
                "search_kernel": search_kernel, "cpu_baseline": cpu, "extra": extra,
                "result": {"min_nll": g, "n_records": int(len(recs)), "n_solved": int(st["n_solved"])}}
        print(json.dumps(line))
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
