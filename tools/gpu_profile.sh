#!/bin/bash
# Usage (on the GPU box, from the repo root): tools/gpu_profile.sh TAG [kernel-regex]
# Plain bench run first (must exit 0), then the ncu launch list and full captures, as B200_PROFILING.md asks.
# The bench command issues 3 warm-up + 8 pipelined steps (quarter-SM analysis launches) and then 4 steps of its roofline
# pass (one CTA per SM, the kernel alone); 3 kernels per step.
TAG=${1:-rXX}; KRE=${2:-yk_k_}
B="python bench.py --steps 8 --warmup 3 --e2e-steps 0 --no-cpu --no-other --no-prewarm --roofline-steps 4"
mkdir -p gpurun_out
$B > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_list_$TAG.log 2>&1
# pipelined region (quarter-SM launch of yk_k_analyze + owner + emit)
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 9 -c 3 -o gpurun_out/prof_$TAG $B > gpurun_out/ncu_full_$TAG.log 2>&1
# roofline pass (one CTA per SM): launches 33.. of the same command
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 36 -c 3 -o gpurun_out/prof_${TAG}_alone $B > gpurun_out/ncu_full_${TAG}_alone.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}_alone.log
