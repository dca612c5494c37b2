"""Decoder-side cross-check of the gradient streams (as a test): the oracle's and — on the emulated
kernels — the product's (bitmap, rgbStream) pairs must be consumable by the reference decoder's corner rule."""
import os
import subprocess

import numpy as np
import pytest

import cases
import decoder_walk
from oracle_py import Oracle, PASS_ORDER
from yaik_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "_build", "libyaik_b200_emu.so")
NAMES = ["synth128_rgb", "patchy_192x136", "ramp64_a2", "patchy_72x40", "noise_lowamp96", "mip8_rgb"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_streams_decode(name):
    planes, _ = cases.SMALL_CASES[name]()
    o = Oracle(planes)
    passes = [o.gradient_pass(sx, sy) for sx, sy in PASS_ORDER]
    n = decoder_walk.check(planes, passes, o.state(5))
    o.close()
    assert n >= 0


@pytest.fixture(scope="module")
def emu_lib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
    return capi.load_library(EMU)


@pytest.mark.parametrize("name", NAMES[:3])
def test_emulated_kernel_streams_decode(emu_lib, name):
    planes, _ = cases.SMALL_CASES[name]()
    c, h, w = planes.shape
    ctx = capi.Context(w, h, planes=4, slots=1, lib=emu_lib)
    try:
        ctx.set_image(planes)
        ctx.analyze(capi.STAGE_GRADIENT)
        passes = [ctx.gradient_pass(sx, sy) for sx, sy in PASS_ORDER]
        decoder_walk.check(planes, passes, ctx.download_state(recon=False)["mappedRGB"][0])
    finally:
        ctx.close()


@pytest.mark.parametrize("name", ["synth128_rgb", "patchy_192x136", "noise_lowamp96", "noise_delta1", "noise_hi"])
def test_oracle_range_streams_decode(name):
    """DynamicTileCompressor's streams through the reference decoder's consumption rule (Decompress1D)."""
    planes, _ = cases.SMALL_CASES[name]()
    c, h, w = planes.shape
    o = Oracle(planes)
    for sx, sy in PASS_ORDER:
        o.gradient_pass(sx, sy)
    claimed = o.state(0).reshape(h, w)[::4, ::4] != 0
    for n in range(3):
        r = o.range1d(n)
        decoder_walk.check_range1d(planes[n], claimed, r["idx"], r["type"])
    o.close()
