# builder-side scaling check of the default bench line on one multi-GPU box: N = 2 and N = 4 (the driver runs 1/2/4/8 itself)
for N in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 6 --warmup 3 --no-cpu --no-second > gpurun_out/bench_scale_n$N.json 2> gpurun_out/bench_scale_n$N.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_scale_n$N.json"))
    e = d["e2e"]
    print("N=$N value %.0f  e2e %.0f  ms/call p50 %.2f  h2d %.1f / %.1f GB/s per rank" % (d["value"], e["value"], e["ms_per_call"]["p50"], e["h2d_gbs_per_rank"], e["h2d_ceiling_gbs_per_rank"]))
except Exception as ex:
    print("N=$N failed", ex); print(open("gpurun_out/bench_scale_n$N.err").read()[-800:])
PY
done
