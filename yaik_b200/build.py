"""Build yaik_b200/csrc/libyaik_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m yaik_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libyaik_b200.so")
SOURCES = ["yk_analyze.cu", "yk_emit.cu", "yk_kernels.cu", "yk_api.cu", "yk_hostpack.cpp", "yk_hosttail.cpp"]
DEPS = SOURCES + ["yk_internal.h", "yk_device.h", os.path.join("..", "..", "include", "yaik_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-O2,-fvisibility=default", "-shared", "-cudart", "static"]


LAST = {"nvcc_ran": False, "seconds": 0.0}      # what the last build() call did (the driver's build check reads the printed line)


def build(force: bool = False, verbose: bool = False) -> str:
    import time
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = [os.path.join(CSRC, d) for d in DEPS]
    force = force or os.environ.get("YK_FORCE_BUILD") == "1"
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        LAST.update(nvcc_ran=False, seconds=0.0)
        return OUT
    cmd = [NVCC, *FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", OUT, *srcs]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True)
    LAST.update(nvcc_ran=True, seconds=round(time.perf_counter() - t0, 1))
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
