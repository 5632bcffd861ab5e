#!/bin/bash
# Development aid: per-source-line executed instructions of one kernel of an .ncu-rep
#   tools/prof_lines.sh report.ncu-rep object.o 'mangled-kernel-prefix' units
set -e
rep=$1; obj=$2; kern=$3; units=${4:-1}
tmp=$(mktemp -d)
ncu -i "$rep" --page raw --csv > $tmp/raw.csv 2>/dev/null
python "$(dirname $0)/ncu_keys.py" $tmp/raw.csv | grep -E "time_dur|fp64|issue_active|thread_inst|pipe_|stalled|registers|warps_active" || true
ncu -i "$rep" --page source --csv > $tmp/src.csv 2>/dev/null
(cd $tmp && cuobjdump -xelf all "$obj" >/dev/null && nvdisasm --print-line-info *.cubin 2>/dev/null | awk -v k="^\\.text\\.$kern" '$0 ~ k {on=1} on&&/^\t\.section/{if(seen)exit} on{seen=1;print}' > lines.txt)
python "$(dirname $0)/sass_by_line.py" $tmp/src.csv $tmp/lines.txt $units
echo "tmp: $tmp"
