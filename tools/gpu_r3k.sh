#!/bin/bash
TAG=${1:-r3k}
mkdir -p gpurun_out
ORBX_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_one_pair_$TAG.csv python tools/one_pair.py 3 > gpurun_out/one_pair_$TAG.log 2>&1
tail -2 gpurun_out/one_pair_$TAG.log
python - $TAG <<'PY'
import csv, sys
rows = [r for r in csv.reader(open("gpurun_out/launches_one_pair_%s.csv" % sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
n = len(rows) // 3
last = rows[-n:]
tot = 0
for r in last:
    name = r[4].split("(")[0].replace("void ", "").replace("orbx::", "")
    us = float(r[-1]) / 1e3
    tot += us
    print("%-26s grid %-16s block %-12s %8.1f us" % (name[:26], r[7] if len(r) > 7 else "", r[6] if len(r) > 6 else "", us))
print("kernels of one call: %d, sum %.1f us" % (n, tot))
PY
