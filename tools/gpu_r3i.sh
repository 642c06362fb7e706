#!/bin/bash
# Multi-warp quadtree: parity (extractor / golden / full-size tests), stage times of the 1024-pair batch and of small batches.
TAG=${1:-r3i}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_extractor.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -x 2>&1 | tail -8 > gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
for p in 1024 1 8 128; do
  timeout 300 python bench.py --pairs $p --steps 4 --warmup 3 --no-cpu --no-second --parity-pairs 4 --e2e-steps 2 > /tmp/b_$p.json 2>/tmp/b_$p.err
  python - $p <<'PY'
import json, sys
p = sys.argv[1]
try:
    d = json.load(open("/tmp/b_%s.json" % p))
    print("pairs=%s value %.0f e2e %.0f batch p50 %.4f ms" % (p, d["value"], d["e2e"]["value"], d["config"]["ms_per_batch"]["p50"]),
          {k: round(v, 4) for k, v in d["roofline"]["stage_ms_per_batch"].items()}, "latency p50", (d.get("latency") or {}).get("p50"))
except Exception as e:
    print("failed", e); print(open("/tmp/b_%s.err" % p).read()[-800:])
PY
done 2>&1 | tee gpurun_out/quadtree_team_$TAG.log
