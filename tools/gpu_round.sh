#!/bin/bash
# One gpurun call = tests + bench + ncu evidence. Usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag> [full]
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -40 > gpurun_out/pytest_$TAG.log
tail -15 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
# launch list of the same command (shares of the step, cold-cache / serialised)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 80 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --pairs 64 --no-cpu > gpurun_out/ncu_list_$TAG.log 2>&1
if [ "$2" = "full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -s 120 -c 40 -f -o gpurun_out/prof_$TAG \
      python bench.py --steps 2 --warmup 3 --pairs 64 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
  tail -3 gpurun_out/ncu_full_$TAG.log
fi
