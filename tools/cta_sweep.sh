#!/bin/bash
# usage (GPU box): tools/cta_sweep.sh  -> pipelined step (MP/s, ms) against the CTAs per analysis launch
for n in 37 36 35 34 32 29; do
  timeout 120 python bench.py --steps 400 --warmup 10 --no-cpu --no-other --e2e-steps 0 --no-r1 --roofline-steps 0 --analysis-ctas $n 2>gpurun_out/sweep.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($n, d['value'], d['ms_per_step'])" 2>&1 | tail -1
done
