#!/bin/bash
# usage (GPU box): tools/ab_run_r1.sh name1 name2 ...   -> step with DynamicTileEncode (MP/s, ms) and the R1 launches alone, per variant library
for v in "$@"; do
  YK_LIB=yaik_b200/csrc/variants/libyaik_b200_$v.so timeout 180 python bench.py --steps 100 --warmup 10 --no-cpu --no-other --e2e-steps 0 --roofline-steps 8 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); w=d['with_r1']; print('$v', d['value'], w['value'], w['ms_per_step'], w['r1_kernel']['ms_alone'])"
done
