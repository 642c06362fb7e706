run() {
  echo "== $* $EXTRA"
  env "$@" timeout 300 python bench.py --steps 2 --warmup 2 --e2e-steps 4 --no-cpu --no-second --parity-pairs 0 $EXTRA > /tmp/b.json 2>/tmp/b.err
  python - <<'PY'
import json
try:
    d = json.load(open("/tmp/b.json"))
    e = d["e2e"]
    print("e2e %.0f fps  ms/call p50 %.2f  group %s  h2d %.1f GB/s of %.1f" % (e["value"], e["ms_per_call"]["p50"], e.get("group_pairs"), e["h2d_gbs_per_rank"], e["h2d_ceiling_gbs_per_rank"]))
except Exception as ex:
    print("failed", ex); print(open("/tmp/b.err").read()[-900:])
PY
}
for g in 96 128 192 256; do EXTRA="--e2e-group $g"; run X=1; done
EXTRA="--e2e-group 128"; run ORBX_LANES=4
EXTRA="--e2e-group 128"; run ORBX_LANES=6
EXTRA="--e2e-group 256"; run ORBX_LANES=4
EXTRA="--e2e-group 128"; run ORBX_SERIAL_EYES=1
