"""Randomised shapes and contents through the CPU-emulated kernels against the oracle: widths / heights that are any
multiple of 4 (partial regions, odd numbers of region columns, units whose second region does not exist, partial macro
tiles), content from flat to noisy, RGB and RGBA.  Seeds are fixed: the cases are the same on every run."""
import os
import subprocess

import numpy as np
import pytest

import cases
from parity import check_image
from yaik_b200 import capi
from yaik_b200.synth import make_image, SEED_BASE, _rand

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "_build", "libyaik_b200_emu.so")


@pytest.fixture(scope="module")
def emu_lib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
    return capi.load_library(EMU)


def _case(i):
    r = _rand(777, i, np.arange(8, dtype=np.uint64))
    w = 4 * (1 + int(r[0] % np.uint64(48)))          # 4 .. 192
    h = 4 * (1 + int(r[1] % np.uint64(40)))          # 4 .. 160
    kind = int(r[2] % np.uint64(4))
    if kind == 0:
        planes = cases._patchy(w, h, 1000 + i, 4, 1 + int(r[3] % np.uint64(4)))
    elif kind == 1:
        planes = cases._smooth_noisy(w, h, 2000 + i, int(r[3] % np.uint64(3)))
    elif kind == 2:
        planes = make_image(w, h, 3, SEED_BASE + 500 + i)
    else:
        planes = cases._noise(w, h, 3, 3000 + i, 90, 90 + int(r[3] % np.uint64(12)))
    stages = ["grad"]
    if w % 8 == 0 and h % 8 == 0:
        stages.append("r2")
    return planes, tuple(stages)


@pytest.mark.parametrize("i", range(24))
def test_random_shape_matches_oracle(emu_lib, i):
    planes, stages = _case(i)
    c, h, w = planes.shape
    ctx = capi.Context(w, h, planes=4, slots=1, lib=emu_lib)
    try:
        ctx.set_upload_format(i % 2 == 0)
        check_image(ctx, planes, stages, fused=(i % 3 != 0))
    finally:
        ctx.close()


@pytest.mark.parametrize("i", range(16))
def test_random_chroma_pipeline_matches_oracle(emu_lib, i):
    from parity import check_chroma
    planes, pre, cfg, modes = cases.random_chroma_case(i)
    c, h, w = planes.shape
    ctx = capi.Context(w, h, planes=4, slots=1, lib=emu_lib)
    try:
        check_chroma(ctx, planes, pre, cfg, modes)
    finally:
        ctx.close()
