"""Development aid: join an `ncu --page source --csv` export (SASS view) with `nvdisasm --print-line-info` of the same
kernel and print executed warp instructions per CUDA source line.  usage: sass_by_line.py src.csv lines.txt [n_units]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    ins.append((int(r[idx['Address']], 16), r[idx['Source']], int(r[idx['Instructions Executed']]), int(r[idx['Thread Instructions Executed']]), int(r[idx['# Samples']])))
base = ins[0][0]
line_of = {}
cur = None
stack = []
for l in open(sys.argv[2]):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), m.group(3).strip()); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m: line_of[int(m.group(1), 16)] = cur
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
agg = collections.Counter(); thr = collections.Counter(); smp = collections.Counter()
tot = 0
for a, s, n, t, sm in ins:
    k = line_of.get(a - base)
    key = (k[0], k[1]) if k else ('?', 0)
    agg[key] += n; thr[key] += t; smp[key] += sm; tot += n
print("total warp inst %d  (%.1f per unit)" % (tot, tot / units))
for k, n in agg.most_common(45):
    print("%-22s %5d  %9.1f per unit  %4.1f%%  act %.1f  samples %d" % (k[0], k[1], n / units, 100.0 * n / tot, thr[k] / max(n, 1), smp[k]))
