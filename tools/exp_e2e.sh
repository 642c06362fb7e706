# e2e pipeline accounting for the default workload: what bounds the host-facing call
run() {
  echo "== $*"
  env "$@" timeout 300 python bench.py --steps 2 --warmup 2 --e2e-steps 4 --no-cpu --no-second --parity-pairs 0 $EXTRA > /tmp/b.json 2>/tmp/b.err
  python - <<'PY'
import json
try:
    d = json.load(open("/tmp/b.json"))
    e = d["e2e"]
    print("e2e %.0f fps  ms/call p50 %.2f  group %s  h2d %.1f GB/s of %.1f" % (e["value"], e["ms_per_call"]["p50"], e.get("group_pairs"), e["h2d_gbs_per_rank"], e["h2d_ceiling_gbs_per_rank"]))
except Exception as ex:
    print("failed", ex); print(open("/tmp/b.err").read()[-600:])
PY
  grep "orbm_stereo_frames_batch" /tmp/b.err | tail -2
}
EXTRA=""
run ORBX_TRACE=1
run ORBX_TRACE=1 ORBX_DEBUG_SKIP_H2D=1
run ORBX_TRACE=1 ORBX_DEBUG_SKIP_KERNELS=1
run ORBX_TRACE=1 ORBX_SERIAL_EYES=1
run ORBX_TRACE=1 ORBX_BLUR_TC=0
EXTRA="--e2e-group 32";  run X=1
EXTRA="--e2e-group 128"; run X=1
EXTRA="--e2e-group 256"; run X=1
