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
  # one full-set capture of every launch of ONE step (the first 26 launches are the 2-pair parity check): 2 extract
  # calls x (7 k_resize_tma + 2 k_fast + k_quadtree + k_blur7 + k_describe) + k_stereo_match + k_stereo_median
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k 'regex:k_fast|k_quadtree|k_describe|k_stereo_match|k_stereo_median|k_blur7|k_resize' -s 26 -c 26 -f \
      -o /tmp/prof_$TAG $BENCH > gpurun_out/ncu_full_$TAG.log 2>&1
  ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
  sz=$(stat -c %s /tmp/prof_$TAG.ncu-rep 2>/dev/null || echo 0)
  echo "prof size $sz"
  if [ "$sz" -gt 0 ] && [ "$sz" -lt 40000000 ]; then cp /tmp/prof_$TAG.ncu-rep gpurun_out/; fi
fi
du -sh gpurun_out
