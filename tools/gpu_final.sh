#!/bin/bash
# Round-end evidence in one gpurun call: parity tests, the default bench line (+ reference arm), configs[2] / [4], the ncu
# launch list of the bench command and one full-set capture of a configs[3] step at the benched launch size.
# Usage: bash tools/gpu_final.sh <tag>
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("value %.0f fps  e2e %.0f fps  ms/batch p50 %.3f  cpu %.0f" % (d["value"], d["e2e"]["value"], d["config"]["ms_per_batch"]["p50"], d["cpu_baseline"]["value"]))
    print("stages", d["roofline"]["stage_ms_per_batch"])
    print("e2e ms/call", d["e2e"]["ms_per_call"], "latency", d.get("latency", {}).get("p50"))
except Exception as e:
    print("bench failed", e)
PY
tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
head -c 600 gpurun_out/bench_${TAG}_ref.json; echo
timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_config4.json 2> gpurun_out/bench_${TAG}_config4.err
timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_config2.json 2> gpurun_out/bench_${TAG}_config2.err
head -c 400 gpurun_out/bench_${TAG}_config2.json; echo
# launch list of the bench command (device-resident leg: grids over 1024 pairs)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --reps 1 --no-cpu --no-second --parity-pairs 0 > gpurun_out/ncu_list_$TAG.log 2>&1
# full-set capture of one step at the benched launch size (stand-alone harness: no Python start-up under ncu)
timeout 900 ncu --set full --clock-control none --import-source on -c 32 -f -o /tmp/prof_$TAG \
    tools/ubench/track_check 1024 1 0 > gpurun_out/prof_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ls -la /tmp/prof_$TAG.ncu-rep
du -sh gpurun_out
