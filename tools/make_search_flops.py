"""profiles/search_flops.json (read by bench.py for `roofline_search`) from an ncu --csv metrics log of ONE launch of the
pass-1 search kernel on the headline window:
    ncu --metrics smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd,ffma,fmul,fadd}_pred_on.sum,... -k regex:n2_search_kernel
        -s 2 -c 1 --csv --log-file profiles/r2_search_flops.csv python tools/prof_run.py search headline
usage: make_search_flops.py profiles/r2_search_flops.csv candidates"""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
mi, vi, ki = h.index("Metric Name"), h.index("Metric Value"), h.index("Kernel Name")
m = {r[mi]: float(r[vi].replace(",", "")) for r in rows[1:]}
kern = rows[1][ki]
n = int(sys.argv[2])
g = lambda k: m.get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % k, 0.0)
out = {"kernel": "%s (pass 1, bench.py headline window: %d candidates)" % (kern, n),
       "source": "ncu smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum of one launch (%s)" % sys.argv[1],
       "candidates": n, "dfma": g("dfma"), "dmul": g("dmul"), "dadd": g("dadd"),
       "ffma": g("ffma"), "fmul": g("fmul"), "fadd": g("fadd"),
       "flops_per_candidate": (2 * g("dfma") + g("dmul") + g("dadd")) / n,
       "fp32_flops_per_candidate": (2 * g("ffma") + g("fmul") + g("fadd")) / n,
       "fp64_pipe_warp_inst": m.get("sm__inst_executed_pipe_fp64.sum"), "warp_inst_total": m.get("smsp__inst_executed.sum"),
       "pipe_fp64_cycles_active_pct": m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
       "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "gpu_time_ns": m.get("gpu__time_duration.sum")}
json.dump(out, open("profiles/search_flops.json", "w"), indent=1)
print(json.dumps(out, indent=1))
