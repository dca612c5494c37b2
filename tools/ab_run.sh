#!/bin/bash
# usage (GPU box): tools/ab_run.sh name1 name2 ...   -> MP/s, ms/step and kernel times of each variant library
for v in "$@"; do
  YK_LIB=yaik_b200/csrc/variants/libyaik_b200_$v.so timeout 120 python bench.py --steps 400 --warmup 10 --no-cpu --no-other --e2e-steps 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
