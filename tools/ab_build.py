#!/usr/bin/env python
"""Developer tool: build variants of the CUDA library with extra -D flags for A/B timing on the GPU box.
usage: tools/ab_build.py name1:-DFOO=1,-DBAR name2:...   ->  yaik_b200/csrc/variants/libyaik_b200_<name>.so"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yaik_b200 import build as B
out_dir = os.path.join(B.CSRC, "variants"); os.makedirs(out_dir, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    out = os.path.join(out_dir, f"libyaik_b200_{name}.so")
    cmd = [B.NVCC, *B.FLAGS, *[f for f in flags.split(",") if f], "-o", out, *[os.path.join(B.CSRC, s) for s in B.SOURCES]]
    subprocess.run(cmd, check=True, capture_output=True)
    print(out)
