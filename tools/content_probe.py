"""Developer tool: kernel-sequence time of one 2048x2048 RGBA texture by content type (all flat / all gradient / all noise /
the bench mix), int32 planes resident in HBM, to see which phase of the analysis kernel bounds which content."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from yaik_b200 import capi
from yaik_b200.synth import make_image, SEED_BASE
lib = capi.load_library()
ctx = capi.Context(2048, 2048, planes=4, slots=1, lib=lib)
ctx.set_upload_format(False)
st = capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D
for name, mix in (("flat", (1, 0, 0, 0)), ("gradient", (0, 1, 0, 0)), ("edge", (0, 0, 1, 0)), ("noise", (0, 0, 0, 1)), ("bench mix", (0.30, 0.40, 0.15, 0.15))):
    img = make_image(2048, 2048, 4, SEED_BASE + 1, mix=mix)
    ctx.set_image(img, 0)
    for _ in range(3):
        ctx.reset_state(0); ctx.analyze(st); ctx.sync()
    N = 30
    t0 = time.perf_counter()
    for _ in range(N):
        ctx.reset_state(0); ctx.analyze(st)
    ctx.sync()
    dt = (time.perf_counter() - t0) / N
    print(f"{name:10s} {dt * 1e6:8.1f} us per texture (memset + analyze + owner + emit, back to back on one stream)")
