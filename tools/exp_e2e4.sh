run() {
  echo "== $*"
  env "$@" timeout 300 python bench.py --steps 2 --warmup 2 --e2e-steps 4 --no-cpu --no-second --parity-pairs 16 $EXTRA > /tmp/b.json 2>/tmp/b.err
  python - <<'PY'
import json
try:
    d = json.load(open("/tmp/b.json"))
    e = d["e2e"]
    print("e2e %.0f fps  ms/call p50 %.2f  group %s  h2d %.1f GB/s of %.1f  latency p50 %.3f  parity: %s" % (e["value"], e["ms_per_call"]["p50"], e.get("group_pairs"), e["h2d_gbs_per_rank"], e["h2d_ceiling_gbs_per_rank"], d["latency"]["p50"], d["config"]["parity"][:40]))
except Exception as ex:
    print("failed", ex); print(open("/tmp/b.err").read()[-900:])
PY
}
EXTRA=""
run X=1
run ORBX_LANE_UPLOADS=1
run ORBX_LANES=4
EXTRA="--e2e-group 128"; run X=1
EXTRA="--e2e-group 32"; run X=1
ORBX_TRACE=2 timeout 300 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu --no-second --parity-pairs 0 > /tmp/b.json 2>/tmp/b.err
grep -A 22 "^group pairs" /tmp/b.err | tail -24
