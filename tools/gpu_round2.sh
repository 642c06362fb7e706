#!/bin/bash
# GPU round: parity tests, default bench (short), configs[4] (knn2 sweep). Usage: bash tools/gpu_round2.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -25 > gpurun_out/pytest_$TAG.log
tail -8 gpurun_out/pytest_$TAG.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("value %.0f fps  e2e %.0f fps  ms/batch p50 %.3f" % (d["value"], d["e2e"]["value"], d["config"]["ms_per_batch"]["p50"]))
    print("stages", d["roofline"]["stage_ms_per_batch"])
    print("e2e ms/call", d["e2e"]["ms_per_call"], "latency", d.get("latency", {}).get("p50"))
except Exception as e:
    print("bench failed", e)
PY
tail -3 gpurun_out/bench_$TAG.err
timeout 300 python bench.py --config 4 --steps 5 --warmup 2 --no-cpu > gpurun_out/bench4_$TAG.json 2> gpurun_out/bench4_$TAG.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench4_$TAG.json"))
    print("config4", d["value"], d["unit"], json.dumps(d.get("config", {}).get("sweep", d.get("sweep", "")))[:600])
except Exception as e:
    print("bench4 failed", e)
PY
tail -3 gpurun_out/bench4_$TAG.err
