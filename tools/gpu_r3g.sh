#!/bin/bash
TAG=${1:-r3g}
mkdir -p gpurun_out
timeout 300 tools/ubench/knn2_abi_check 2>&1 | tail -3
for c in 20 1000; do
  ORBM_KNN2_C0=$c timeout 300 python bench.py --config 4 --steps 10 --warmup 3 > /tmp/c4_$c.json 2>/tmp/c4.err || tail -5 /tmp/c4.err
  python - $c <<'PY'
import json, sys
d = json.loads(open("/tmp/c4_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("c0=%s" % sys.argv[1], [(r['n'], round(r['ms'], 4)) for r in d['config']['rows']])
PY
done 2>&1 | tee gpurun_out/knn2_c0_$TAG.log
