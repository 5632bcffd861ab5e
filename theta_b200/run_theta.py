"""RunTHetA-compatible command line (python/RunTHetA.py:278-509, python/FileIO.py:45-227) on the GPU search.

    python -m theta_b200.run_theta QUERY_FILE [-n N] [-k K] [-t TAU] [-d DIR] [-p PREFIX] [-m MAX_NORMAL]
           [--NUM_PROCESSES P] [--NUM_INTERVALS I] [--BOUND_HEURISTIC BH] [--NORMAL_BOUND_HEURISTIC NBH]
           [--HEURISTIC_LB LB] [--HEURISTIC_UB UB] [--BOUNDS_ONLY] [--NO_MULTI_EVENT] [--RESULTS file] [--FORCE]
           [--GET_VALUES] [--NO_INTERVAL_SELECTION] [--RATIO_DEV x] [--MIN_FRAC x]

Same flags, same <prefix>.n<N>.withBounds / <prefix>.n<N>.results / <prefix>.RunN3.bash / <prefix>.BEST.results
outputs; without -n the reference's default flow (n=2, n=3 seeded by the n=2 files, ModelSelection).  Differences,
all outside the search path: K may be 0..7 (the reference's parser stops at 6 although it says 0-7, Q4);
bounds read from the file are integers (Q11); floats print with Python-3 repr (Q12); the CPU timing loop of
TimeEstimate is replaced by the exact candidate count; plotting, the BAF model and SNP clustering are not
part of this build (out of scope), so --BAF / --TUMOR_FILE / --NORMAL_FILE are rejected.
"""
import argparse
import os
import sys

from . import _lib, pipeline
from .search import do_optimization, do_optimization_single
from .time_estimate import time_estimate


def parse_arguments(argv=None):
    p = argparse.ArgumentParser(prog="RunTHetA")
    p.add_argument("QUERY_FILE")
    p.add_argument("-n", "--N", type=int, default=None)
    p.add_argument("-k", "--MAX_K", type=int, default=3)
    p.add_argument("-t", "--TAU", type=int, default=2)
    p.add_argument("-d", "--DIR", default="./")
    p.add_argument("-p", "--OUTPUT_PREFIX", default=None)
    p.add_argument("-m", "--MAX_NORMAL", type=float, default=1.0)
    p.add_argument("--NUM_PROCESSES", type=int, default=1)
    p.add_argument("--NUM_INTERVALS", type=int, default=100)
    p.add_argument("--BOUND_HEURISTIC", default=False)
    p.add_argument("--NORMAL_BOUND_HEURISTIC", type=int, default=False)
    p.add_argument("--HEURISTIC_LB", type=float, default=0.9)
    p.add_argument("--HEURISTIC_UB", type=float, default=1.1)
    p.add_argument("--BOUNDS_ONLY", action="store_true")
    p.add_argument("--NO_MULTI_EVENT", action="store_true")
    p.add_argument("--RESULTS", default=None)
    p.add_argument("--FORCE", action="store_true")
    p.add_argument("--GET_VALUES", action="store_true")
    p.add_argument("--NO_INTERVAL_SELECTION", action="store_true")
    p.add_argument("--RATIO_DEV", type=float, default=0.1)
    p.add_argument("--MIN_FRAC", type=float, default=0.05)
    p.add_argument("--BAF", action="store_true")
    p.add_argument("--TUMOR_FILE", default=None)
    p.add_argument("--NORMAL_FILE", default=None)
    a = p.parse_args(argv)
    if not os.path.isfile(a.QUERY_FILE):
        p.error("The specified file does not exist: " + a.QUERY_FILE)
    if a.N is not None and a.N not in (2, 3):
        p.error("Invalid value for n. n must be 2 or 3")  # FileIO.py:42
    if a.MAX_K not in range(8):
        p.error("Invalid value for k: " + str(a.MAX_K) + ". Supported values of k are 0-7")
    if a.MAX_NORMAL < 0 or a.MAX_NORMAL > 1:
        p.error("Invalid value for max_normal: " + str(a.MAX_NORMAL) + ". Max_normal must be between 0 and 1")
    if a.BAF or a.TUMOR_FILE or a.NORMAL_FILE:
        p.error("the BAF model and SNP clustering are outside this build (no bnpy / matplotlib path)")
    if a.OUTPUT_PREFIX is None:
        a.OUTPUT_PREFIX = os.path.basename(a.QUERY_FILE).split(".")[0]
    if a.NUM_INTERVALS == 100 and a.N == 3:
        a.NUM_INTERVALS = 20  # FileIO.py:170: the n=3 default
    return a


def run_fixed_N(n, a, intervals, resultsfile=None, argv=None):
    """python/RunTHetA.py:296-509 without the clustering / BAF / plotting branches."""
    lengths, tumor, normal, m, upper_bounds, lower_bounds = intervals
    k, tau, max_normal = a.MAX_K, a.TAU, a.MAX_NORMAL
    prefix, directory = a.OUTPUT_PREFIX, a.DIR
    frac = pipeline.determine_frac_copy_num(normal, tumor, lengths, a.RATIO_DEV)
    print("Frac with potential copy numbers:", frac)
    if frac < a.MIN_FRAC:
        print("ERROR: This sample does not have enough large copy number aberrations to be a good candidate for tumor "
              "composition estimation using THetA.  See --RATIO_DEVIATION and --MIN_FRAC flags to modify how the "
              "potential presence of large copy number aberrations is determined.  Exiting...")
        sys.exit(1)
    choose = not a.NO_INTERVAL_SELECTION
    all_tumor, all_normal = tumor, normal
    order = None
    if choose:
        print("Selecting intervals...")
        if n == 2:
            if lower_bounds is None or upper_bounds is None:
                order, lengths, tumor, normal = pipeline.select_intervals_n2(lengths, tumor, normal, m, k, a.FORCE,
                                                                             a.NUM_INTERVALS)
                upper_bounds = lower_bounds = None
            else:
                order, lengths, tumor, normal, lower_bounds, upper_bounds = pipeline.select_intervals_n2(
                    lengths, tumor, normal, m, k, a.FORCE, a.NUM_INTERVALS, lower_bounds, upper_bounds)
        else:
            if resultsfile is None:
                print("ERROR: No results file supplied. Unable to automatically select intervals for n=3 without results "
                      "of n=2 analysis. See --RESULTS flag, or --NO_INTERVAL_SELECTION to disable interval selection. "
                      "Exiting...")
                sys.exit(1)
            copy = pipeline.read_results_file(resultsfile)
            order, lengths, tumor, normal, upper_bounds, lower_bounds, copy = pipeline.select_intervals_n3(
                lengths, tumor, normal, m, upper_bounds, lower_bounds, copy, tau, a.FORCE, a.NUM_INTERVALS)
        m = len(order)
    print("Preprocessing data...")
    r, rN, sorted_index = pipeline.sort_r(normal, tumor)
    if a.NORMAL_BOUND_HEURISTIC is not False:
        upper_bounds, lower_bounds = pipeline.calculate_bounds_normal_heuristic(
            a.NORMAL_BOUND_HEURISTIC, a.HEURISTIC_LB, a.HEURISTIC_UB, r, rN, m, k)
    elif a.BOUND_HEURISTIC is not False or (upper_bounds is None and lower_bounds is None):
        bh = 0.5 if a.BOUND_HEURISTIC is False else float(a.BOUND_HEURISTIC)
        upper_bounds, lower_bounds = pipeline.calculate_bounds_heuristic(bh, r, rN, m, tau, k)
    else:
        upper_bounds = pipeline.sort_by_sorted_index(upper_bounds, sorted_index)
        lower_bounds = pipeline.sort_by_sorted_index(lower_bounds, sorted_index)
    ub_out = pipeline.reverse_sort_list(upper_bounds, sorted_index)
    lb_out = pipeline.reverse_sort_list(lower_bounds, sorted_index)
    boundsfile = pipeline.write_out_bounds(directory, prefix, a.QUERY_FILE, ub_out, lb_out, n, order if choose else None)
    if a.BOUNDS_ONLY:
        sys.exit(0)
    # TimeEstimate.time_estimate (python/TimeEstimate.py:40-89): same counts, messages and aborts; the per-candidate
    # time is measured on the GPU instead of on 100 / 20 CPU solves
    time_estimate(n, m, k, tau, lower_bounds, upper_bounds, r, rN, max_normal, sorted_index, a.NUM_PROCESSES,
                  not a.NO_MULTI_EVENT, a.FORCE)
    print("Performing optimization...")
    pre = prefix  # RunTHetA.py:307-308: --GET_VALUES writes <prefix>.likelihoods in the working directory
    if a.NUM_PROCESSES == 1:
        best = do_optimization_single(n, m, k, tau, lower_bounds, upper_bounds, r, rN, max_normal, sorted_index,
                                      not a.NO_MULTI_EVENT, a.GET_VALUES, prefix=pre)
    else:
        best = do_optimization(n, m, k, tau, lower_bounds, upper_bounds, r, rN, max_normal, sorted_index,
                               a.NUM_PROCESSES, not a.NO_MULTI_EVENT, a.GET_VALUES, prefix=pre)
    if best == []:
        print("ERROR: Maximum Likelihood Solution not found within given bounds.")
        sys.exit(1)
    if n == 2 and pipeline.best_near_max_contamination(best, max_normal):
        print("WARNING: At least one of the top solutions is near the upper bound on normal contamination. Further "
              "analysis may required (see --MAX_NORMAL).")
    r = pipeline.reverse_sort_list(r, sorted_index)
    rN = pipeline.reverse_sort_list(rN, sorted_index)
    if choose:
        if n == 2:
            best = pipeline.calc_all_c_2(best, r, rN, all_tumor, all_normal, order)
            best = pipeline.find_mins(best)
        else:
            # RunTHetA.py:470-474: --NO_MULTI_EVENT only selects which extension routine runs (Q3)
            if a.NO_MULTI_EVENT:
                best = pipeline.calc_all_c_3(best, r, rN, all_tumor, all_normal, order)
            else:
                best = pipeline.calc_all_c_3_multi_event(best, r, rN, all_tumor, all_normal, order)
            best = pipeline.find_mins(best)
    resultsfile = pipeline.write_out_result(directory, prefix, best, n)
    if n == 2:
        pipeline.write_out_N3_script(directory, prefix, a.QUERY_FILE, argv)
    return resultsfile, boundsfile


def main(argv=None):
    a = parse_arguments(argv)
    _lib.init()
    intervals = pipeline.read_interval_file(a.QUERY_FILE)
    full_argv = ["RunTHetA.py"] + list(argv if argv is not None else sys.argv[1:])
    if a.N is None:
        # RunTHetA.py:291-295: n=2, then n=3 seeded by the n=2 bounds / results files, then ModelSelection
        res2, bounds2 = run_fixed_N(2, a, intervals, argv=full_argv)
        intervals = pipeline.read_interval_file(bounds2)
        res3, bounds3 = run_fixed_N(3, a, intervals, res2, argv=full_argv)
        pipeline.model_selection(a.QUERY_FILE, res2, res3)
        return res3, bounds3
    return run_fixed_N(a.N, a, intervals, a.RESULTS, argv=full_argv)


if __name__ == "__main__":
    main()
