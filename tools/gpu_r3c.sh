#!/bin/bash
# knn2 with the query tile in tensor memory (k_knn2_tc_ts) against the shared-memory form and the POPC kernel.
TAG=${1:-r3c}
mkdir -p gpurun_out
{
  echo "== ORBM_KNN2_TS=1 (query tile in tensor memory, N = 224)"; timeout 300 tools/ubench/knn2_abi_check
  echo "== ORBM_KNN2_TS=0 (query tile in shared memory, N = 256)"; ORBM_KNN2_TS=0 timeout 300 tools/ubench/knn2_abi_check
} 2>&1 | tee gpurun_out/knn2_ts_$TAG.log
timeout 600 python -m pytest tests/test_gpu_matcher.py -m gpu -q --tb=short -x -k knn2 2>&1 | tail -5 | tee gpurun_out/pytest_$TAG.log
for v in 1 0; do
  ORBM_KNN2_TS=$v timeout 300 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/bench_config4_ts${v}_$TAG.json 2>/tmp/c4.err || tail -5 /tmp/c4.err
  python - $v $TAG <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_config4_ts%s_%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
print("TS=%s" % sys.argv[1], d.get("value"), d.get("unit"), json.dumps(d.get("config", {}).get("sweep", d.get("sweep")))[:600])
PY
done 2>&1 | tee -a gpurun_out/knn2_ts_$TAG.log

