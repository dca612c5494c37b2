#!/usr/bin/env python
"""Developer tool: build yaik_b200/csrc/libyaik_b200_timing.so (the product sources with -DYK_TIMING: clock64 buckets in
the producer and one consumer warp of yk_k_analyze), to be run on the GPU box with tools/timing_run.py."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yaik_b200 import build as B
out = os.path.join(B.CSRC, "libyaik_b200_timing.so")
subprocess.run([B.NVCC, *B.FLAGS, "-DYK_TIMING", "-o", out, *[os.path.join(B.CSRC, s) for s in B.SOURCES]], check=True)
print(out)
