"""Kernel LOGIC on the CPU: the product's .cu sources compiled against tests/emu/cuda_emu.h (OS threads standing in
for CUDA threads) and driven through the same C ABI, checked against the oracle.  This does not replace the
GPU parity tests (tests/test_gpu_parity.py, -m gpu) — it finds indexing / ordering mistakes before GPU time is
spent.  The emulated library is test infrastructure; the product has no CPU path."""
import os
import subprocess

import pytest

import cases
from parity import check_chroma, check_image
from yaik_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "_build", "libyaik_b200_emu.so")


@pytest.fixture(scope="module")
def emu_lib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
    return capi.load_library(EMU)


EMU_CASES = ["r1_signed96", "r1_signed_edge64", "synth256_rgb_3bit", "alpha_island256", "ramp64_a2", "patchy_72x40", "patchy128", "mip32_rgba", "mip16_rgba", "mip8_rgb", "mip4_rgb",
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


def test_emulated_alternative_data_paths(emu_lib):
    """Upload formats (packed bytes / int32), device-resident planes, yk_fetch_all, the range check of both uploads."""
    import numpy as np
    import paths_check
    ctx = capi.Context(256, 256, planes=4, slots=1, lib=emu_lib)
    keep = []
    try:
        paths_check.check_upload_formats(ctx)
        paths_check.check_out_of_range(ctx)
        paths_check.check_fetch_all(ctx)

        def to_device(planes):       # the emulated "device" is host memory
            keep.append(np.ascontiguousarray(planes, dtype=np.int32))
            return [keep[-1][i].ctypes.data for i in range(planes.shape[0])]
        paths_check.check_device_resident_planes(ctx, to_device)
    finally:
        ctx.close()


@pytest.mark.parametrize("name,pre,cfg,modes", cases.CHROMA_CASES)
def test_emulated_chroma_front_end(emu_lib, name, pre, cfg, modes):
    planes, _ = cases.SMALL_CASES[name]()
    ctx = capi.Context(256, 256, planes=4, slots=1, lib=emu_lib)
    try:
        check_chroma(ctx, planes, pre, cfg, modes)
    finally:
        ctx.close()
