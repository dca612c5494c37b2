#!/bin/bash
# Usage (GPU box, repo root): tools/gpu_profile2.sh TAG [kernel-regex]
# Plain bench run first (must exit 0), then the ncu launch list and ONE full capture of the dominant kernel as launched
# alone (one CTA per SM) with source correlation.  The bench command issues 3 warm-up + 1 region of 8 pipelined steps
# (3 kernels per step), a parity pass of 8 images and then the roofline pass (the kernel alone).
TAG=${1:-rXX}; KRE=${2:-yk_k_analyze}
B="python bench.py --steps 8 --warmup 3 --e2e-steps 0 --no-cpu --no-other --no-r1 --no-prewarm --roofline-steps 4"
mkdir -p gpurun_out
$B > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_list_$TAG.log 2>&1
# the last launches of the command belong to the roofline pass: take the kernel there (skip all but the last few matches)
N=$(grep -c "$KRE" gpurun_out/launches_$TAG.csv)
SKIP=$((N > 3 ? N - 3 : 0))
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -o gpurun_out/prof_${TAG}_alone $B > gpurun_out/ncu_full_${TAG}_alone.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}_alone.log
# DynamicTileEncode's two kernels as launched behind the analysis (three planes per launch)
BR="python bench.py --steps 8 --warmup 3 --e2e-steps 0 --no-cpu --no-other --no-prewarm --roofline-steps 4"
ncu --set full --clock-control none --import-source on -k regex:yk_k_r1_ -s 40 -c 2 -o gpurun_out/prof_${TAG}_r1 $BR > gpurun_out/ncu_full_${TAG}_r1.log 2>&1
tail -1 gpurun_out/ncu_full_${TAG}_r1.log
