"""Developer tool: device time of yk_k_analyze fused (alpha + gradient + R2 in one launch) against two launches (alpha +
gradient, then R2 alone on the claim masks) on the bench texture - does a shorter instruction stream per launch pay?"""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yaik_b200 import capi
from yaik_b200.synth import make_image, SEED_BASE
lib = capi.load_library()
lib.yk_profile.argtypes = [C.c_void_p, C.c_int]
lib.yk_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
ctx = capi.Context(2048, 2048, planes=4, slots=1, lib=lib)
ctx.set_upload_format(False)
img = make_image(2048, 2048, 4, SEED_BASE + 1)
ctx.set_image(img, 0)
def run(seq, n=30):
    for it in range(n + 5):
        if it == 5: lib.yk_profile(ctx.ctx, 1)
        ctx.reset_state(0)
        for st in seq: ctx.analyze(st)
        ctx.sync()
    a = (C.c_double * 8)(); b = (C.c_longlong * 8)()
    lib.yk_profile_read(ctx.ctx, a, b); lib.yk_profile(ctx.ctx, 0)
    return [round(a[k] / n * 1e3, 2) for k in range(3)], [int(b[k]) // n for k in range(3)]
A, G, R = capi.STAGE_ALPHA, capi.STAGE_GRADIENT, capi.STAGE_RANGE1D
print("fused            analyze/emit/owner us per image:", *run([A | G | R]))
print("gradient, then R2 analyze/emit/owner us per image:", *run([A | G, R]))
print("gradient only                                   :", *run([A | G]))
