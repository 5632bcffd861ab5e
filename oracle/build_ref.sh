#!/bin/sh
# Snapshot of the reference's own hot-path modules for timing on the GPU box (bench.py --impl reference,
# cpu_baseline.kind = "reference").  /root/reference exists only in the build container; the GPU box receives
# oracle/_ref/ with the rest of the working tree (it is git-ignored, NOT gpurun-ignored), so this copy never enters the
# repository's history.  The files are executed through oracle/refshim.py (py2 -> py3 print rewrites at import time,
# no arithmetic line touched).  TEST / MEASUREMENT INFRASTRUCTURE: the product never imports it.
set -e
SRC="${THETA_REFERENCE_SRC:-/root/reference/python}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref/python"
if [ ! -d "$SRC" ]; then
    echo "build_ref: $SRC not present (GPU box: using the prebuilt oracle/_ref if any)"
    exit 0
fi
mkdir -p "$DST"
for f in Misc Enumerator Optimizer CalcAllC DataTools TimeEstimate SelectIntervals FileIO RunTHetA ModelSelection; do
    cp "$SRC/$f.py" "$DST/$f.py"
done
[ -f "$SRC/../doc/LICENSE.txt" ] && cp "$SRC/../doc/LICENSE.txt" "$HERE/_ref/LICENSE.txt"
echo "build_ref: $(ls "$DST" | wc -l) reference modules -> $DST"
