"""Developer tool: the pipelined step of bench.py (8 contexts / streams, quarter-GPU analysis launches) with and without the
ownership / emission kernels behind each analysis (yk_strip_phase(0) = reset + analysis only on a one-strip image): what the
small kernels cost the pipeline, and what four analysis launches side by side achieve by themselves."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from yaik_b200 import capi
from yaik_b200.synth import make_image, SEED_BASE
lib = capi.load_library(os.environ.get('YK_LIB'))
N = int(os.environ.get('YK_PROBE_CTXS', '8'))
ctxs = [capi.Context(2048, 2048, planes=4, slots=1, lib=lib) for _ in range(N)]
sts = [torch.cuda.Stream() for _ in range(N)]
for i, (c, s) in enumerate(zip(ctxs, sts)):
    c.set_stream(s.cuda_stream); c.set_upload_format(False)
    c.set_image(make_image(2048, 2048, 4, SEED_BASE + 1 + i), 0)
    c.strip_config(2048, 0)
ST = capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D
def full(i):
    c = ctxs[i % N]; c.reset_state(0); c.analyze(ST)
def only(i):
    c = ctxs[i % N]; c.reset_state(0); c.strip_phase(0)
for ctas in (37, 0):     # a quarter of the SMs per launch (the bench's pipelined step) / one CTA per SM
    for c in ctxs: c.set_analysis_ctas(ctas)
    for name, fn in (("analysis + owner + emit", full), ("analysis only", only)):
        for i in range(40): fn(i)
        torch.cuda.synchronize()
        K = 400
        t0 = time.perf_counter()
        for i in range(K): fn(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{ctas or 148:4d} CTAs per analysis launch, {name:24s}: {dt / K * 1e6:6.2f} us per texture")

