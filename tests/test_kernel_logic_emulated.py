"""Kernel LOGIC on the CPU: the product's .cu sources compiled against tests/emu/cuda_emu.h (OS threads standing in
for CUDA threads) and driven through the same C ABI, checked against the oracle.  This does not replace the
GPU parity tests (tests/test_gpu_parity.py, -m gpu) — it finds indexing / ordering mistakes before GPU time is
spent.  The emulated library is test infrastructure; the product has no CPU path."""
import os
import subprocess

import pytest

import cases
from parity import check_image
from yaik_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "_build", "libyaik_b200_emu.so")


@pytest.fixture(scope="module")
def emu_lib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
    return capi.load_library(EMU)


EMU_CASES = ["synth256_rgb_3bit", "alpha_island256", "ramp64_a2", "patchy_72x40", "patchy128", "mip32_rgba", "mip16_rgba", "mip8_rgb", "mip4_rgb",
             "alpha_island128", "alpha_corner_only", "noise_delta1", "noise_hi", "flat64"]


@pytest.mark.parametrize("name", EMU_CASES)
def test_emulated_kernels_match_oracle(emu_lib, name):
    planes, stages = cases.SMALL_CASES[name]()
    ctx = capi.Context(256, 256, planes=4, slots=1, lib=emu_lib)
    try:
        check_image(ctx, planes, stages, fused=True)
    finally:
        ctx.close()


def test_emulated_stage_by_stage_calls(emu_lib):
    planes, stages = cases.SMALL_CASES["patchy_72x40"]()
    ctx = capi.Context(128, 64, planes=4, slots=1, lib=emu_lib)
    try:
        check_image(ctx, planes, ("grad", "r2"), fused=False)
    finally:
        ctx.close()
