#!/bin/bash
# A/B of a variant library (ORBX_SO_PATH) against the built one: stage times of the 1024-pair batch + single-pair latency.
TAG=${1:-r3m}; VAR=${2:-orb_slam3_fast_b200/liborbx_stereo8.so}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for so in "" /root/repo/$VAR; do
  ORBX_SO_PATH=$so timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-second --parity-pairs 4 --e2e-steps 2 > /tmp/b.json 2>/tmp/b.err
  python - "$so" <<'PY'
import json, sys
try:
    d = json.load(open("/tmp/b.json"))
    print("lib=%s value %.0f e2e %.0f batch p50 %.4f ms" % (sys.argv[1] or "default", d["value"], d["e2e"]["value"], d["config"]["ms_per_batch"]["p50"]),
          {k: round(v, 4) for k, v in d["roofline"]["stage_ms_per_batch"].items()}, "latency p50", d["latency"]["p50"])
except Exception as e:
    print("failed", e); print(open("/tmp/b.err").read()[-800:])
PY
done 2>&1 | tee gpurun_out/variant_ab_$TAG.log
