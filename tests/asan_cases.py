"""Run by tests/test_emulated_asan.py in a subprocess with libasan preloaded: a few analysis / R1 / chroma cases through the
AddressSanitizer build of the emulated library.  Any out-of-bounds access of the kernels (as emulated) or of the C-ABI
layer aborts the process."""
import sys

import cases
from parity import check_chroma, check_image
from yaik_b200 import capi

lib = capi.load_library(sys.argv[1])
for name in ("patchy_72x40", "alpha_island128", "r1_signed96", "mip8_rgb"):
    planes, stages = cases.SMALL_CASES[name]()
    ctx = capi.Context(128, 128, planes=4, slots=1, lib=lib)
    check_image(ctx, planes, stages, fused=True)
    ctx.close()
for name, pre, cfg, modes in (cases.CHROMA_CASES[0], cases.CHROMA_CASES[2]):
    planes, _ = cases.SMALL_CASES[name]()
    ctx = capi.Context(128, 128, planes=4, slots=1, lib=lib)
    check_chroma(ctx, planes, pre, cfg, modes)
    ctx.close()
print("asan cases ok")
