#!/bin/bash
# One gpurun call = tests + bench + ncu evidence. Usage (GPU box, repo root): bash tools/gpu_round.sh <tag> [full]
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -40 > gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
BENCH="python bench.py --steps 2 --warmup 3 --pairs 64 --no-cpu"
# launch list of the same command (shares of the step; cold-cache / serialised)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/ncu_list_$TAG.log 2>&1
if [ "$2" = "full" ]; then
  # one full-set capture per kernel family (ncu replays every kernel ~40x: keep the counts small)
  timeout 600 ncu --set full --clock-control none --import-source on \
      -k 'regex:k_fast|k_quadtree|k_describe|k_stereo_match|k_stereo_median|k_blur7' -s 20 -c 6 -f -o /tmp/prof_a_$TAG \
      $BENCH > gpurun_out/ncu_a_$TAG.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_resize' -s 28 -c 7 -f \
      -o /tmp/prof_b_$TAG $BENCH > gpurun_out/ncu_b_$TAG.log 2>&1
  for x in a b; do
    ncu -i /tmp/prof_${x}_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${x}_${TAG}_raw.csv 2>/dev/null
    sz=$(stat -c %s /tmp/prof_${x}_$TAG.ncu-rep 2>/dev/null || echo 0)
    echo "prof_$x size $sz"
    if [ "$sz" -gt 0 ] && [ "$sz" -lt 30000000 ]; then cp /tmp/prof_${x}_$TAG.ncu-rep gpurun_out/; fi
  done
fi
du -sh gpurun_out
