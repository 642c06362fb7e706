#!/bin/bash
# Quick GPU round: parity tests + short bench (no CPU baseline) + launch list. Usage: bash tools/gpu_quick.sh <tag> [pytest-args]
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -x ${2:-} 2>&1 | tail -25 > gpurun_out/pytest_$TAG.log
tail -8 gpurun_out/pytest_$TAG.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("value %.0f fps  e2e %.0f fps  ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
    print("stages", {k: round(v, 3) for k, v in d["roofline"]["stage_ms_per_step"].items()})
    print("roofline", d["roofline"]["kernel"], round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 4))
except Exception as e:
    print("bench failed", e)
PY
tail -3 gpurun_out/bench_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --pairs 64 --no-cpu > gpurun_out/ncu_list_$TAG.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += float(r[-1])
tot = sum(v[1] for v in agg.values()) or 1
for k, v in agg.items():
    print("%-28s n=%3d %9.1f us %5.1f%%" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
PY
# optional: PROF=<kernel regex> captures ONE launch with the full set + source counters (read here with ncu -i)
if [ -n "$PROF" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$PROF" -s ${PROF_SKIP:-2} -c 1 -f \
      -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --pairs 64 --no-cpu > gpurun_out/ncu_prof_$TAG.log 2>&1
  ls -la gpurun_out/prof_$TAG.ncu-rep
fi
# optional: UBENCH=1 runs the stand-alone prototypes of tools/ubench (self-checking, watchdog-bounded)
if [ -n "$UBENCH" ]; then
  for args in "1000 1000" "10000 10000" "100000 100000" "100000 100000 0 0"; do
    timeout 120 tools/ubench/knn2_tc $args 2>&1 | tail -12
  done > gpurun_out/knn2_tc_$TAG.log
  tail -20 gpurun_out/knn2_tc_$TAG.log
  # loads-in-flight variants of the epilogue, and one full-set capture of the tensor-core kernel (stand-alone binary:
  # no Python start-up under ncu)
  for g in 2 4; do timeout 120 tools/ubench/knn2_tc 100000 100000 0 1 $g 2>&1 | tail -2; done >> gpurun_out/knn2_tc_$TAG.log
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_knn2_tc" -c 1 -f \
      -o gpurun_out/prof_knn2_tc_$TAG tools/ubench/knn2_tc 100000 100000 > gpurun_out/ncu_knn2_tc_$TAG.log 2>&1
  ls -la gpurun_out/prof_knn2_tc_$TAG.ncu-rep
fi
