"""TimeEstimate.time_estimate (python/TimeEstimate.py:40-142) for the GPU search.

The reference times 100 (n=2) / 20 (n=3) CPU solves and multiplies by a candidate count: exact for n=2
(count_number_matrices_2, TimeEstimate.py:91-111), an over-estimate for n=3 (count_number_matrices_3, :113-142).  Here
the counts are the same DPs (Python integers, so they never overflow) and the per-candidate time is MEASURED on the GPU by
searching a short prefix of the space; the messages, the n=3 / m > 30 guard and the > 200 h abort without --FORCE are the
reference's (TimeEstimate.py:48-50, 74-85).
"""
import sys

from . import _lib
from .search import n3_row_table

_EXCESSIVE = ("WARNING: With the current settings, the runtime is likely excessive. To reduce runtime, try:\n\t1) Increase "
              "the number of processes used with the --NUM_PROCESSES flag.\n\t2) Reduce the number of intervals chosen using "
              "the --NUM_INTERVALS flag.\n\t3) Disable automatic interval selection using --NO_INTERVAL_SELECTION, and "
              "hand-select a smaller number of intervals, or set tighter bounds on the current intervals.\n\t Run with "
              "--FORCE to continue with current settings.")


def count_number_matrices_2(m, upper_bounds, lower_bounds):
    """Exact number of candidates for n=2 (TimeEstimate.py:91-111): non-decreasing columns within the bounds."""
    ub = [int(v) for v in upper_bounds]
    lb = [int(v) for v in lower_bounds]
    K = max(ub)
    poss = [0] * (K + 1)
    for i in range(lb[0], ub[0] + 1):
        poss[i] += 1
    for row in range(1, m):
        new = [0] * (K + 1)
        for j, v in enumerate(poss):
            if v > 0:
                for c in range(max(j, lb[row]), ub[row] + 1):
                    new[c] += v
        poss = new
    return sum(poss)


def count_number_matrices_3(m, upper_bounds, lower_bounds, tau=2):
    """The reference's over-estimate for n=3 (TimeEstimate.py:113-142): paths in the row graph within the bounds, halved
    for the column symmetry, mu-bound pruning ignored (Python 2 integer division)."""
    ub = [int(v) for v in upper_bounds]
    lb = [int(v) for v in lower_bounds]
    rows = n3_row_table(tau, max(ub))
    edges = [[w for w, rw in enumerate(rows) if w == v or rw[0] > rv[0] or rw[1] > rv[1]] for v, rv in enumerate(rows)]
    ok = lambda row, i: all(lb[i] <= a <= ub[i] for a in row)
    poss = [1 if ok(row, 0) else 0 for row in rows]
    for i in range(m - 1):
        new = [0] * len(rows)
        for j, v in enumerate(poss):
            if v > 0:
                for w in edges[j]:
                    if ok(rows[w], i + 1):
                        new[w] += v
        poss = new
    return sum(poss) // 2


def measure_rate(n, r, rN, lower_bounds, upper_bounds, tau, max_normal):
    """Candidates per second of one GPU on a prefix of this very search (after one warm-up run)."""
    plan = _lib.Plan(n, r, rN, lower_bounds, upper_bounds, tau=tau, max_normal=max_normal)
    try:
        if n == 2:
            total = min(plan.count(), 1 << 24)
            if total == 0:
                return None
            for _ in range(2):
                plan.pass1(0, total)
            ms = plan.stats()["ms_pass1"]
            return total / (ms * 1e-3) if ms > 0 else float("inf")
        ni = plan.prepare(256, 8192)
        if ni == 0:
            return None
        for _ in range(2):
            plan.search_min_n3(0, ni)
        st = plan.stats()
        return st["n_candidates"] / (st["ms_pass1"] * 1e-3) if st["ms_pass1"] > 0 else float("inf")
    finally:
        plan.close()


def time_estimate(n, m, k, tau, lower_bounds, upper_bounds, r, rN, max_normal, sorted_index, num_processes, multi_event,
                  force, out=None):
    """Prints what the reference prints and exits where it exits; returns {'count', 'rate', 'seconds'}.
    num_processes is the number of GPUs the search will use (RunTHetA's --NUM_PROCESSES)."""
    say = (lambda *a, **kw: print(*a, **kw)) if out is None else (lambda *a, **kw: print(*a, file=out, **kw))
    say("Estimating time...")
    if n == 3 and m > 30 and not force:
        say("\tWARNING: With n=3 and", m, "intervals, the runtime would likely be excessive. Try reducing the number of "
            "intervals below 25. Run with --FORCE to continue.")
        sys.exit(1)
    try:
        lo, up = _monotone(lower_bounds, upper_bounds)
    except ValueError:
        lo, up = list(lower_bounds), list(upper_bounds)
    if n == 2:
        count = count_number_matrices_2(m, up, lo)
    else:
        count = count_number_matrices_3(m, up, lo, tau)
    try:
        rate = measure_rate(n, r, rN, lower_bounds, upper_bounds, tau, max_normal) if count > 0 or n == 3 else None
    except _lib.EmptySpaceError:
        rate = None
    if rate is None:
        say("ERROR: No valid Copy Number Profiles exist for these intervals within the bounds specified. Exiting...")
        sys.exit(1)
    gpus = max(1, min(int(num_processes or 1), max(1, _lib.device_count())))
    seconds = count / rate / gpus
    say("\tCandidate copy number profiles:", count, "(exact)" if n == 2 else "(upper estimate)")
    say("\tEstimated Total Time:", end=" ")
    if seconds < 60:
        say(int(seconds + .5), "second(s)")
    elif seconds < 3600:
        say(int((seconds / 60) + .5), "minute(s)")
    else:
        hours = int((seconds / 3600) + .5)
        say(hours, "hour(s)")
        if hours > 200 and not force:
            say(_EXCESSIVE)
            sys.exit(1)
    return {"count": count, "rate": rate, "seconds": seconds}


def _monotone(lower, upper):
    """Enumerator._check_bound_order (python/Enumerator.py:90-113) on copies."""
    lo, up = [int(v) for v in lower], [int(v) for v in upper]
    for i in range(1, len(lo)):
        if lo[i] < lo[i - 1]:
            lo[i] = lo[i - 1]
    for i in range(len(up) - 2, -1, -1):
        if up[i] > up[i + 1]:
            up[i] = up[i + 1]
    return lo, up
