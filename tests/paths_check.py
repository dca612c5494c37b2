"""Checks shared by the CPU (emulated kernels) and GPU test files for the alternative data paths of the C ABI:
upload formats, device-resident int32 planes, yk_fetch_all."""
import ctypes as C

import numpy as np

import cases
from oracle_py import Oracle, PASS_ORDER
from parity import check_image, _eq
from yaik_b200 import capi


def check_upload_formats(ctx):
    """yk_set_image with the packed (bytes) and the plain int32 upload give the oracle's results."""
    for packed in (False, True):
        ctx.set_upload_format(packed)
        for name in ("patchy_72x40", "synth256_rgba", "mip8_rgb"):
            planes, stages = cases.SMALL_CASES[name]()
            check_image(ctx, planes, stages, fused=True)
    ctx.set_upload_format(True)


def check_out_of_range(ctx):
    """A sample outside 0..255 is reported by both upload paths (host-side check when packing, device-side otherwise)."""
    planes = cases.SMALL_CASES["flat64"]()[0].copy()
    planes[1, 10, 10] = 300
    for packed in (False, True):
        ctx.set_upload_format(packed)
        ctx.set_image(planes)
        ctx.analyze(capi.STAGE_GRADIENT)
        try:
            ctx.gradient_pass(4, 4)
        except capi.YaikError as e:
            assert e.code == -4
        else:
            raise AssertionError("out-of-range sample not reported (packed=%s)" % packed)
    ctx.set_upload_format(True)


def check_device_resident_planes(ctx, to_device):
    """yk_set_image_device on int32 planes the caller put in device memory (the path bench.py's `value` measures)."""
    planes, stages = cases.SMALL_CASES["synth256_rgba"]()
    c, h, w = planes.shape
    ptrs = to_device(planes)
    ctx.set_image_device(ptrs, c, w, h)
    ctx.analyze(capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D)
    o = Oracle(planes)
    for k, (sx, sy) in enumerate(PASS_ORDER):
        want, got = o.gradient_pass(sx, sy), ctx.gradient_pass(sx, sy)
        assert got["tiledone"] == want["tiledone"] and got["bbox"] == want["bbox"]
        _eq(got["bitmap"], want["bitmap"], f"pass {k} bitmap"); _eq(got["rgb"], want["rgb"], f"pass {k} rgb")
    for n in range(3):
        want, got = o.range1d(n), ctx.range1d(n)
        _eq(got["idx"], want["idx"], f"R2 idx {n}"); _eq(got["type"], want["type"], f"R2 type {n}")
    o.close()


def check_fetch_all(ctx):
    """yk_fetch_all returns what the per-stage getters return."""
    planes, stages = cases.SMALL_CASES["synth256_rgba"]()
    ctx.set_image(planes)
    ctx.analyze(capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D)
    allr = ctx.fetch_all()
    for k, (sx, sy) in enumerate(PASS_ORDER):
        g = ctx.gradient_pass(sx, sy)
        a = allr["passes"][k]
        assert a["tiledone"] == g["tiledone"] and a["bbox"] == g["bbox"]
        _eq(a["bitmap"], g["bitmap"], f"pass {k} bitmap"); _eq(a["rgb"], g["rgb"], f"pass {k} rgb")
    for n in range(3):
        g = ctx.range1d(n)
        _eq(allr["r2"][n]["idx"], g["idx"], f"R2 idx {n}"); _eq(allr["r2"][n]["type"], g["type"], f"R2 type {n}")
    al = ctx.alpha_reject()
    assert allr["alpha"]["bound"] == al["bound"] and allr["alpha"]["remaining"] == al["remaining"] and allr["alpha"]["wrote"] == al["wrote"]
    _eq(allr["alpha"]["bitmap"], al["bitmap"], "alpha bitmap")
