#!/bin/bash
# Round-2 late A/B: single-pass k_track_enum (always on) and the compass pre-test in the minThFAST retry (ORBX_FAST_PRETEST=3).
TAG=${1:-r3a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
ORBX_FAST_PRETEST=3 timeout 600 python -m pytest tests/test_gpu_extractor.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -x 2>&1 | tail -8 > gpurun_out/pytest_${TAG}_pretest3.log
tail -3 gpurun_out/pytest_${TAG}_pretest3.log
for v in 1 3 1 3; do
  ORBX_FAST_PRETEST=$v timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-second --parity-pairs 4 --e2e-steps 2 > /tmp/b_$v.json 2>/tmp/b_$v.err
  python - $v <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open("/tmp/b_%s.json" % v))
    print("pretest=%s value %.0f e2e %.0f batch p50 %.3f" % (v, d["value"], d["e2e"]["value"], d["config"]["ms_per_batch"]["p50"]), d["roofline"]["stage_ms_per_batch"])
except Exception as e:
    print("failed", e); print(open("/tmp/b_%s.err" % v).read()[-800:])
PY
done 2>&1 | tee gpurun_out/ab_$TAG.log
