import sys, os, time
sys.path.insert(0, '/root/repo')
import torch
from yaik_b200 import capi
from yaik_b200.synth import make_image, SEED_BASE
lib = capi.load_library(os.environ['YK_LIB'])
ctx = capi.Context(2048, 2048, planes=4, slots=1, lib=lib)
ctx.set_upload_format(False)
img = make_image(2048, 2048, 4, SEED_BASE + 1)
ctx.set_image(img, 0)
st = capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D
for _ in range(5):
    ctx.reset_state(0); ctx.analyze(st); ctx.sync()
ts = []
for _ in range(20):
    ctx.reset_state(0); ctx.sync()
    t0 = time.perf_counter(); ctx.analyze(st); ctx.sync(); ts.append(time.perf_counter() - t0)
ts.sort()
print(os.environ['YK_LIB'].split('_')[-1], "wall per analyze call (3 kernels + sync): median %.1f us min %.1f us" % (ts[10] * 1e6, ts[0] * 1e6))
