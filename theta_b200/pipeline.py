"""Host-side steps either side of the search, written for Python 3.

These are the O(m) callers of the hot path in the reference's run_fixed_N
(python/RunTHetA.py:296-509): input/output files, interval selection, sorting, the bound
heuristics, and the post-search extension of the winning C to the intervals that were not
searched.  They are plain host code (text I/O and a few float operations per interval); the
search itself and every NLL evaluation go through the C ABI.
"""
import math
import os
import sys

import numpy as np

from . import calc_all_c

MIN_LENGTH_N2 = 1000000  # python/SelectIntervals.py:45
MIN_LENGTH_N3 = 5000000  # python/SelectIntervals.py:46


# --------------------------------------------------------------------------- FileIO
def read_interval_file(filename):
    """python/FileIO.py:386-446.  Bounds columns are cast to int here (the reference keeps strings, Q11).

    Returns [lengths, tumor_counts, norm_counts, m, upper_bounds | None, lower_bounds | None];
    a missing bound is the string 'X'.
    """
    lengths, tumor, normal, upper, lower = [], [], [], [], []
    n_line = 0
    with open(filename) as f:
        for raw in f:
            if raw.startswith("#"):
                continue
            cols = raw.strip().replace(" ", "\t").split()
            n_line += 1
            if len(cols) < 6 or len(cols) > 8:
                sys.stderr.write("Invalid input file format in interval file line #" + str(n_line) + ":\n" + str(cols) +
                                 "\nToo few/many columns. Exiting...\n")
                sys.exit(1)
            lengths.append(int(cols[3]) - int(cols[2]))
            tumor.append(int(cols[4]))
            normal.append(int(cols[5]))
            upper.append(_bound(cols[6]) if len(cols) > 6 else "X")
            lower.append(_bound(cols[7]) if len(cols) > 7 else "X")
    if n_line == 1:
        sys.stderr.write("Number of intervals must be greater than 1. Exiting...\n")
        sys.exit(1)
    if all(x == "X" for x in upper):
        upper = None
    if all(x == "X" for x in lower):
        lower = None
    return [lengths, tumor, normal, len(lengths), upper, lower]


def _bound(tok):
    return "X" if tok == "X" else int(float(tok))


def read_results_file(filename):
    """python/FileIO.py:484-503: the C column of the first solution of a .results file."""
    with open(filename) as f:
        lines = [ln for ln in f.readlines() if not ln.startswith("#")]
    if not lines:
        print("ERROR: The result file provided appears to be empty. Exiting...")
        sys.exit(1)
    if len(lines) > 1:
        print("WARNING: The results file contains more than one solution. THetA will use the first provided solution.")
    return lines[0].strip().split("\t")[2].split(":")


def write_out_bounds(directory, prefix, input_file, upper_bounds, lower_bounds, n, order=None):
    """python/FileIO.py:733-784: the input file with UpperBound / LowerBound columns ('X' where unused)."""
    with open(input_file) as f:
        lines = f.readlines()
    out = os.path.join(directory, prefix + ".n" + str(n) + ".withBounds")
    print("Writing bounds file to", out)
    if "#" in lines[0]:
        lines = lines[1:]
    where = None if order is None else {v: i for i, v in enumerate(order)}
    with open(out, "w") as f:
        f.write("#ID\tchrm\tstart\tend\ttumorCount\tnormalCount\tUpperBound\tLowerBound\n")
        for i, line in enumerate(lines):
            f.write("\t".join(line.strip().split("\t")[:6]).strip())
            j = i if where is None else where.get(i)
            if j is None:
                f.write("\tX\tX\n")
            else:
                f.write("\t" + str(int(upper_bounds[j])) + "\t" + str(int(lower_bounds[j])) + "\n")
    return out


def write_out_result(directory, prefix, results, n):
    """python/FileIO.py:620-664: '#NLL mu C p*' lines; floats print with Python-3 repr (Q12)."""
    path = os.path.join(directory, prefix + ".n" + str(n) + ".results")
    print("Writing results file to", path)
    with open(path, "w") as f:
        f.write("#NLL\tmu\tC\tp*\n")
        for C, mu, L, vals in results:
            rows = []
            for i in range(C.shape[0]):
                rows.append(",".join("X" if int(C[i][j]) == -1 else str(int(C[i][j])) for j in range(1, C.shape[1])))
            f.write(str(L) + "\t" + ",".join(str(x) for x in mu) + "\t" + ":".join(rows) + "\t" +
                    ",".join(str(v) for v in vals) + "\n")
    return path


def write_out_N3_script(directory, prefix, input_file, argv=None):
    """python/FileIO.py:786-798: a bash line that restarts at n=3 from the n=2 outputs."""
    path = os.path.join(directory, prefix + ".RunN3.bash")
    print("Writing script to run N=3 to ", path)
    args = " ".join(argv if argv is not None else sys.argv)
    bounds = os.path.join(directory, prefix + ".n2.withBounds")
    results = os.path.join(directory, prefix + ".n2.results")
    with open(path, "w") as f:
        f.write("#!/bin/bash\n")
        f.write("python " + args.replace("-n 2", "").replace(input_file, bounds) + " -n 3" + " --RESULTS " + results)
    return path


# --------------------------------------------------------------------------- DataTools
def determine_frac_copy_num(rN, r, lengths, dev):
    """python/DataTools.py:162-200: fraction of the genome whose normalised tumour/normal ratio is off 1 by > dev."""
    sum_r, sum_rN = sum(r), sum(rN)
    off = 0
    for t, n_, ln in zip(r, rN, lengths):
        if n_ == 0:
            continue
        ratio = (t * 1.0 / n_) * (1.0 * sum_rN / sum_r)
        if ratio > 1.0 + dev or ratio < 1.0 - dev:
            off += ln
    return float(off) / float(sum(lengths))


def sort_r(rN, r):
    """python/DataTools.py:95-118: ascending normalised r_i/rN_i, stable.  Returns (r, rN, sorted_index)."""
    sum_r, sum_rN = float(sum(r)), float(sum(rN))
    ratio = [(t * 1.0 / n_) * (sum_rN / sum_r) for n_, t in zip(rN, r)]
    order = sorted(range(len(ratio)), key=lambda i: ratio[i])
    return [r[i] for i in order], [rN[i] for i in order], order


def sort_by_sorted_index(vec, sorted_index):
    return [vec[i] for i in sorted_index]


def reverse_sort_list(vec, sorted_index):
    out = [0] * len(sorted_index)
    for i, s in enumerate(sorted_index):
        out[s] = vec[i]
    return out


def calculate_bounds_heuristic(x, r, rN, m, tau, k):
    """python/DataTools.py:47-67.  Returns (upper_bounds, lower_bounds)."""
    print("Calculating bounds using bound heuristic...")
    sum_r, sum_rN = float(sum(r)), float(sum(rN))
    ratios = [(float(t) / sum_r) / (float(n_) / sum_rN) for t, n_ in zip(r, rN)]
    mean = (1.0 / m) * sum(ratios)
    std = ((1.0 / (m - 1)) * sum([(mean - q) ** 2 for q in ratios])) ** .5
    c = mean + x * std
    lower, upper = [0] * m, [tau] * m
    for i, q in enumerate(ratios):
        if q > c:
            y = round(tau * q)
            lower[i] = max(tau, y - 1)
            upper[i] = max(k, y + 1)
    return upper, lower


def calculate_bounds_normal_heuristic(nbh, heuristic_lb, heuristic_ub, r, rN, m, k):
    """python/DataTools.py:69-93.  Returns (upper_bounds, lower_bounds)."""
    print("Calculating bounds using normal bound heuristic...")
    sum_r, sum_rN = float(sum(r)), float(sum(rN))
    ratios = [(float(t) / sum_r) / (float(n_) / sum_rN) for t, n_ in zip(r, rN)]
    upper, lower = [nbh] * m, [nbh] * m
    for j, q in enumerate(ratios):
        if q < heuristic_lb:
            lower[j], upper[j] = 0, nbh
        elif q > heuristic_ub:
            if q > 2:
                y = round(nbh * q)
                lower[j], upper[j] = y - 1, max(k, y + 1)
            else:
                lower[j], upper[j] = nbh, k
    return upper, lower


# --------------------------------------------------------------------------- SelectIntervals
def select_intervals_n2(lengths, tumor, normal, m, k, force, num_intervals, lower=None, upper=None):
    """python/SelectIntervals.py:126-160 (+ filter_intervals_n2 :206-218): the longest usable intervals,
    returned in input order as columns [index, length, tumor, normal(, lower, upper)]."""
    total_len = float(sum(lengths))
    tot_t, tot_n = float(sum(tumor)), float(sum(normal))
    idx = [i for i in range(m) if tumor[i] > 0 and normal[i] > 0 and lengths[i] >= MIN_LENGTH_N2]
    idx = [i for i in idx if ((tumor[i] / tot_t) / (normal[i] / tot_n)) < float(k + 1) / 2]
    if lower is None or upper is None:
        lines = [[i, lengths[i], tumor[i], normal[i]] for i in idx]
    else:
        lines = [[i, lengths[i], tumor[i], normal[i], lower[i], upper[i]] for i in idx]
    lines.sort(key=lambda x: x[1])
    lim = min(num_intervals, len(idx))
    top = lines[-lim:] if lim else []
    if float(sum(t[1] for t in top)) < 0.1 * total_len:
        msg = ("WARNING: This sample isn't a good candidate for THetA analysis. The longest  %d intervals chosen for "
               "analysis represent <10%% of the combined length of all provided intervals." % lim)
        if not force:
            print(msg + " Run with --FORCE flag to ignore this warning. Exiting...")
            sys.exit(1)
        print(msg)
    top.sort(key=lambda x: x[0])
    print("\tSelected", len(top), "intervals for analysis.")
    return [[t[c] for t in top] for c in range(len(top[0]))]


def select_intervals_n3(lengths, tumor, normal, m, upper, lower, copy, tau, force, num_intervals):
    """python/SelectIntervals.py:50-124: intervals for the n=3 run, chosen from the n=2 answer `copy`."""
    if tau != 2:
        print("ERROR: For automatic interval selection with 3 subpopulations, the default copy number (--TAU) must be 2. "
              "To run with other values, bounds must be provided in the input file.")
        sys.exit(1)
    used = [i for i in range(m) if upper[i] != "X"]
    rows = [[i, lengths[i], tumor[i], normal[i], int(upper[i]), int(lower[i]), int(copy[i])] for i in used]
    b = int(math.ceil(num_intervals * .75))
    c = int(num_intervals - b)
    lines = [row for row in rows if row[1] >= MIN_LENGTH_N3]
    lines.sort(key=lambda x: -x[1])
    chosen = []
    for i, line in enumerate(lines):
        if c > 0 and line[6] == 2 and line[4] == 2:
            chosen.append(i)
            c -= 1
        elif b > 0 and line[6] in (0, 1, 3):
            chosen.append(i)
            b -= 1
    for i, line in enumerate(lines):
        if c > 0 and line[6] == 2 and line[4] > 2:
            chosen.append(i)
            c -= 1
    if c > 0 or b > 0:
        msg = ("WARNING: This sample isn't a good candidate for THetA analysis with 3 subpopulations: There aren't a "
               "sufficient number of intervals that fit the criteria for interval selection.")
        if not force:
            print(msg + " Run with --FORCE flag to ignore this warning. Exiting...")
            sys.exit(1)
        print(msg)
    top = [lines[i] for i in chosen]
    for line in top:
        cn = line[6]
        if cn == 1:
            line[5] = 1
        elif cn == 2:
            line[5] = 1
            line[4] = min(3, line[4])
        elif cn != 0:
            line[4] = 3
    top.sort(key=lambda x: x[0])
    print("\tSelected", len(chosen), "intervals for analysis.")
    return [[t[c_] for t in top] for c_ in range(7)]


# --------------------------------------------------------------------------- CalcAllC (post-search)
def _calculate_x(tumor_i, normal_i, sum_r, sum_all, mu, n, row, h):
    """python/CalcAllC.py:78-90: the real-valued copy number of population h that zeroes dL/dx for one interval."""
    row = [v * normal_i for v in row]
    nR = float(tumor_i) / (sum_r + tumor_i)
    s = sum([row[i] * mu[i] for i in range(n) if i != h])
    return float(nR * (sum_all + s) - s) / ((1 - nR) * mu[h])


def calc_all_c_2(best, r, rN, all_tumor, all_normal, intervals_used):
    """python/CalcAllC.py:92-146: extend each winning n=2 C to the intervals left out of the search by comparing
    the NLL at floor/ceil of the closed-form optimum.  L2 is the GPU evaluator (theta_b200.calc_all_c.L2) and, as
    in the reference, rescales the matrix it is handed in place, call after call (Q10)."""
    out = []
    n_all = len(all_tumor)
    used = set(intervals_used)
    for c, mu, likelihood, vals in best:
        m, n = c.shape
        c_new = np.zeros((m + 1, n))
        c_new[:m] = c
        c_new = c_new * np.asarray(list(rN) + [0], dtype=np.float64)[:, None]
        c_all = np.zeros((n_all, n))
        for i, v in enumerate(intervals_used):
            c_all[v] = c[i]
        sum_all = sum([c_new[j][k] * mu[k] for j in range(m) for k in range(n)])
        sum_r = sum(r)
        for i in range(n_all):
            if i in used:
                continue
            c_all[i][0] = 2
            if all_normal[i] == 0:
                c_all[i][1] = -1
                continue
            x = _calculate_x(all_tumor[i], all_normal[i], sum_r, sum_all, mu, n, [2, 0], 1) / all_normal[i]
            if x < 0:
                c_all[i][1] = 0
                continue
            bot, top = math.floor(x), math.ceil(x)
            c_new[m][0] = 2 * all_normal[i]
            c_new[m][1] = bot * all_normal[i]
            l_bot = calc_all_c.L2(mu[0], c_new, m + 1, list(r) + [all_tumor[i]])
            c_new[m][1] = top * all_normal[i]
            l_top = calc_all_c.L2(mu[0], c_new, m + 1, list(r) + [all_tumor[i]])
            c_all[i][1] = int(bot) if l_bot[0] < l_top[0] else int(top)
        c_all_w = c_all * np.asarray(all_normal, dtype=np.float64)[:, None]
        likelihood, vals = calc_all_c.L2(mu[0], c_all_w, n_all, all_tumor)
        out.append([(c_all, mu, likelihood, vals)])
    return out


class _RowProbe(object):
    """One winning n=3 solution extended by a trial row for an interval that was not searched.

    Holds the weighted matrix of the searched intervals plus one spare last row (python/CalcAllC.py:149-167,
    253-271) and evaluates L3 with that row set to (2, x, y) * normal_i."""

    def __init__(self, c, mu, r, rN):
        self.m, self.n = c.shape
        self.mu = mu
        self.r = list(r)
        self.cw = np.zeros((self.m + 1, self.n))
        self.cw[:self.m] = c * np.asarray(rN, dtype=np.float64)[:, None]
        # |C_w mu| over the searched intervals and the searched tumour reads (CalcAllC.py:165-167)
        self.sum_all = sum([self.cw[j][k] * mu[k] for j in range(self.m) for k in range(self.n)])
        self.sum_r = sum(r)

    def optimum(self, tumor_i, normal_i, row, h):
        """Real-valued copy number of population h that zeroes dL/dx, in units of copies."""
        return _calculate_x(tumor_i, normal_i, self.sum_r, self.sum_all, self.mu, self.n, row, h) / normal_i

    def nll(self, tumor_i, normal_i, x, y):
        last = self.cw[self.m]
        last[0], last[1], last[2] = 2 * normal_i, x * normal_i, y * normal_i
        return calc_all_c.L3(self.mu, self.cw, self.m + 1, self.r + [tumor_i], self.n)[0]


def _floor_ceil(v):
    return int(max(0, math.floor(v))), int(max(0, math.ceil(v)))


def _extend_n3(best, r, rN, all_tumor, all_normal, intervals_used, choose_row):
    """Shared driver of calc_all_c_3 / calc_all_c_3_multi_event: fill in every unsearched interval with the row
    `choose_row` picks, then re-evaluate L3 over all intervals."""
    out = []
    n_all = len(all_tumor)
    searched = set(intervals_used)
    weights = np.asarray(all_normal, dtype=np.float64)[:, None]
    for c, mu, _likelihood, _vals in best:
        probe = _RowProbe(c, mu, r, rN)
        c_all = np.zeros((n_all, probe.n))
        c_all[np.asarray(intervals_used, dtype=np.int64)] = c
        for i in (j for j in range(n_all) if j not in searched):
            c_all[i][0] = 2
            if all_normal[i] == 0:  # no normal reads: no optimum exists, printed as X
                c_all[i][1] = c_all[i][2] = -1
            else:
                c_all[i][1], c_all[i][2] = choose_row(probe, all_tumor[i], all_normal[i])
        likelihood, vals = calc_all_c.L3(mu, c_all * weights, n_all, all_tumor, probe.n)
        out.append([(c_all, mu, likelihood, vals)])
    return out


def _row_single_event(probe, t_i, n_i):
    """python/CalcAllC.py:183-243: candidates (floor/ceil x, 2), (2, floor/ceil y) and the diagonal (j, j) walked
    upwards until the NLL rises; the smallest (NLL, row) pair wins (list sort: NLL first, then the row)."""
    cands = []
    for x in _floor_ceil(probe.optimum(t_i, n_i, [2, 0, 2], 1)):
        cands.append((probe.nll(t_i, n_i, x, 2), [x, 2]))
    for y in _floor_ceil(probe.optimum(t_i, n_i, [2, 2, 0], 2)):
        cands.append((probe.nll(t_i, n_i, 2, y), [2, y]))
    prev, j = float("inf"), 0
    while True:
        l = probe.nll(t_i, n_i, j, j)
        cands.append((l, [j, j]))
        j += 1
        if l > prev:
            break
        prev = l
    cands.sort()
    return cands[0][1]


def _row_multi_event(probe, t_i, n_i):
    """python/CalcAllC.py:286-321: x = 0..ceil(x*), y = floor/ceil of the conditional optimum kept on the same side
    of 2 as x; the first strictly smaller NLL wins."""
    max_x = max(0, math.ceil(probe.optimum(t_i, n_i, [2, 0, 0], 1)))
    best_l, best_row = float("inf"), None
    for x in range(int(max_x) + 1):
        bot, top = _floor_ceil(probe.optimum(t_i, n_i, [2, x, 0], 2))
        if x < 2:
            bot, top = min(bot, 2), min(top, 2)
        elif x > 2:
            bot, top = max(2, bot), max(2, top)
        for y in (bot, top):
            l = probe.nll(t_i, n_i, x, y)
            if l < best_l:
                best_l, best_row = l, (x, y)
    return best_row


def calc_all_c_3(best, r, rN, all_tumor, all_normal, intervals_used):
    """CalcAllC.calc_all_c_3 (python/CalcAllC.py:147-249), the --NO_MULTI_EVENT extension."""
    return _extend_n3(best, r, rN, all_tumor, all_normal, intervals_used, _row_single_event)


def calc_all_c_3_multi_event(best, r, rN, all_tumor, all_normal, intervals_used):
    """CalcAllC.calc_all_c_3_multi_event (python/CalcAllC.py:251-328), the default extension."""
    return _extend_n3(best, r, rN, all_tumor, all_normal, intervals_used, _row_multi_event)


def find_mins(best):
    """python/RunTHetA.py:107-122: merge per-solution lists, comparing only each list's first NLL (Misc.isClose)."""
    m_ = float("inf")
    keep = []
    for solns in best:
        if len(solns) == 0:
            continue
        L = solns[0][2]
        if not (abs(m_ - L) > 10e-4):
            keep += solns
        elif L < m_:
            m_ = L
            keep = list(solns)
    return keep


def best_near_max_contamination(best, max_normal):
    """python/RunTHetA.py:222-225."""
    return any(abs(max_normal - mu[0]) < .01 for _c, mu, _l, _v in best)


# --------------------------------------------------------------------------- ModelSelection
def load_results(filename, tau=2):
    """python/FileIO.py:801-833: [(nll, C as lists of strings with tau prepended, mu)] of a .results file."""
    with open(filename) as f:
        lines = f.readlines()
    if lines and lines[0].startswith("#"):
        lines = lines[1:]
    if len(lines) == 0:
        sys.stderr.write("ERROR: The result file provided appears to be empty. Exiting...\n")
        sys.exit(1)
    results = []
    for l in lines:
        soln = l.strip().split("\t")
        mu = [float(i) for i in soln[1].split(",")]
        C = [[tau] + [i for i in row.split(",")] for row in soln[2].split(":")]
        results.append((float(soln[0]), C, mu))
    return results


def get_frac_breakdown(C, lengths):
    """python/ModelSelection.py:151-188: genome fractions with no aberration / clonal / subclonal aberrations."""
    tot_len = tot_norm = tot_clonal = tot_subclonal = 0
    for i, row in enumerate(C):
        cur = lengths[i]
        tot_len += cur
        row_str = [str(x) for x in row[1:]]
        if all(v == "X" for v in row_str):
            continue
        elif all(v == "2" for v in row_str):
            tot_norm += cur
        elif all(v == row_str[0] for v in row_str):
            tot_clonal += cur
        else:
            tot_subclonal += cur
    return float(tot_norm) / float(tot_len), float(tot_clonal) / float(tot_len), float(tot_subclonal) / float(tot_len)


def additional_criteria(n2_result, n3_result, input_file, min_pop=0.05, min_clonal=0.0, max_ratio=5, min_ratio=0.05):
    """python/ModelSelection.py:109-149: n=3 is kept only if one of its solutions has all tumour populations above
    min_pop, some clonal aberration, and a subclonal/clonal length ratio within [min_ratio, max_ratio)."""
    lengths = read_interval_file(input_file)[0]
    is_valid = False
    for _nll, C, mu in load_results(n3_result):
        _no_cna, clonal, subclonal = get_frac_breakdown(C, lengths)
        pop_is_big = all(i > min_pop for i in mu[1:])
        clonal_is_big = clonal > min_clonal
        small_enough, big_enough = False, True
        if clonal > 0:
            ratio = float(subclonal) / float(clonal)
            small_enough = ratio < max_ratio
            if ratio < min_ratio:
                big_enough = False
        if pop_is_big and clonal_is_big and small_enough and big_enough:
            is_valid = True
    return (3, n3_result) if is_valid else (2, n2_result)


def _min_nll(path):
    best = float("inf")
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            best = min(best, float(line.strip().split("\t")[0]))
    return best


def model_selection(input_file, n2_result, n3_result):
    """python/ModelSelection.py:43-107: penalised NLL  2 L + (m+1)(n-1) log(T+N)  decides between the n=2 and the n=3
    result files; the winner is copied to <prefix>.BEST.results."""
    num_tumor = num_normal = num_intervals = 0
    with open(input_file) as f:
        for line in f:
            if line.startswith("#"):
                continue
            tumor, normal = line.strip().split("\t")[4:6]
            if int(normal) > 0:
                num_tumor += int(tumor)
                num_normal += int(normal)
                num_intervals += 1
    p2 = 2 * _min_nll(n2_result) + (num_intervals + 1) * math.log(num_tumor + num_normal)
    p3 = 2 * _min_nll(n3_result) + (num_intervals + 1) * 2 * math.log(num_tumor + num_normal)
    selected_num, selected_res = 2, n2_result
    if p3 <= p2:
        selected_num, selected_res = additional_criteria(n2_result, n3_result, input_file)
    filename = selected_res.replace(".n" + str(selected_num) + ".results", ".BEST.results")
    print("Selected n=" + str(selected_num) + " solution.  Writing to", filename)
    with open(filename, "w") as out, open(selected_res) as src:
        for line in src:
            out.write(line)
    return selected_num, filename
