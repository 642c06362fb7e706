timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-second --parity-pairs 4 --e2e-steps 2 > /tmp/b.json 2>/tmp/b.err
python - <<'PY'
import json
try:
    d = json.load(open("/tmp/b.json"))
    print("value %.0f e2e %.0f batch p50 %.3f" % (d["value"], d["e2e"]["value"], d["config"]["ms_per_batch"]["p50"]), d["roofline"]["stage_ms_per_batch"])
except Exception as e:
    print("failed", e); print(open("/tmp/b.err").read()[-800:])
PY
