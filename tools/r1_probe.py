"""Developer tool: time of the DynamicTileEncode (R1, 3/4 bits per pixel) calls on the bench texture after the fused analysis."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yaik_b200 import capi
from yaik_b200.synth import make_image, SEED_BASE
lib = capi.load_library()
ctx = capi.Context(2048, 2048, planes=4, slots=1, lib=lib)
ctx.set_upload_format(False)
img = make_image(2048, 2048, 4, SEED_BASE + 1)
ctx.set_image(img, 0)
ctx.analyze(capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D); ctx.sync()
for mode3 in (False, True):
    for _ in range(2):
        for p in range(3):
            r = ctx.range_dyn(p, mode3=mode3)
    t0 = time.perf_counter(); N = 5
    for _ in range(N):
        for p in range(3):
            r = ctx.range_dyn(p, mode3=mode3)
    dt = (time.perf_counter() - t0) / N
    print(f"mode3BitOnly={mode3}: 3 planes {dt * 1e3:.3f} ms wall (kernels + D2H of {r['nibbles'].size} nibble bytes and {r['defs'].size} defs per plane)")
# chroma front-end: yk_chroma_prepare + DynamicTileEncode of Y / half-width Co / half-width Cg (the CLI's configuration)
for _ in range(2):
    ctx.chroma((1, 0, 1, 0), (2, 2), planes=False)
t0 = time.perf_counter()
for _ in range(5):
    ctx.chroma((1, 0, 1, 0), (2, 2), planes=False)
print(f"chroma pipeline (prepare + 3 encodes, streams and dst planes downloaded): {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms wall")
