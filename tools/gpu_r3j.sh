#!/bin/bash
# Recorded (CUDA graph) small stereo calls: the replay test, the track / shim suites, and the single-pair latency with and without.
TAG=${1:-r3j}
mkdir -p gpurun_out
true
true
for g in 1 0; do
  ORBX_GRAPH=$g timeout 300 python bench.py --pairs 64 --reps 2 --steps 3 --warmup 3 --no-cpu --no-second --parity-pairs 2 --e2e-steps 2 > /tmp/b_$g.json 2>/tmp/b_$g.err
  python - $g <<'PY'
import json, sys
g = sys.argv[1]
try:
    d = json.load(open("/tmp/b_%s.json" % g))
    print("ORBX_GRAPH=%s latency" % g, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["latency"].items() if k in ("p50", "p90", "p99", "n")})
except Exception as e:
    print("failed", e); print(open("/tmp/b_%s.err" % g).read()[-1500:])
PY
done 2>&1 | tee gpurun_out/graph_latency_$TAG.log
