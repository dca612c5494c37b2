"""Developer tool: yk_k_analyze / yk_k_emit / yk_k_owner timed alone (yk_profile) for batch launches of n textures
(2048x2048 RGBA bench textures, one kernel launch over n slots): per-texture times."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yaik_b200 import capi
from yaik_b200.synth import make_image, SEED_BASE
lib = capi.load_library(os.environ.get('YK_LIB'))
NMAX = 8
ctx = capi.Context(2048, 2048, planes=4, slots=NMAX, lib=lib)
ctx.set_upload_format(False)
for s in range(NMAX):
    ctx.set_image(make_image(2048, 2048, 4, SEED_BASE + 1 + s), s)
st = capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D
for n in (1, 2, 4, 8):
    for _ in range(3):
        ctx.reset_states(0, n); ctx.analyze(st, 0, n); ctx.sync()
    lib.yk_profile(ctx.ctx, 1)
    for _ in range(20):
        ctx.reset_states(0, n); ctx.analyze(st, 0, n); ctx.sync()
    a = (C.c_double * 8)(); b = (C.c_longlong * 8)()
    lib.yk_profile_read(ctx.ctx, a, b); lib.yk_profile(ctx.ctx, 0)
    print(f"batch of {n}:", {k: round(a[i] / max(b[i], 1) * 1e3 / n, 2) for i, k in enumerate(["analyze_us_per_texture", "emit_us_per_texture", "owner_us_per_texture"])})
