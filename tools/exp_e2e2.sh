ORBX_TRACE=2 timeout 300 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu --no-second --parity-pairs 0 > /tmp/b.json 2>/tmp/b.err
grep -A 30 "^group pairs" /tmp/b.err | tail -64
