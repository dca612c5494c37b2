import sys, os, time
sys.path.insert(0, '/root/repo')
import torch
from yaik_b200 import capi
from yaik_b200.synth import make_image, SEED_BASE
lib = capi.load_library()
N=8
ctxs=[capi.Context(2048,2048,planes=4,slots=1,lib=lib) for _ in range(N)]
sts=[torch.cuda.Stream() for _ in range(N)]
img=make_image(2048,2048,4,SEED_BASE+1)
for c,s in zip(ctxs,sts):
    c.set_stream(s.cuda_stream); c.set_upload_format(False); c.set_image(img,0); c.set_analysis_ctas(37)
st=capi.STAGE_ALPHA|capi.STAGE_GRADIENT|capi.STAGE_RANGE1D
def step(i):
    c=ctxs[i%N]; c.reset_state(0); c.analyze(st)
for i in range(40): step(i)
torch.cuda.synchronize()
for rep in range(3):
    t0=time.perf_counter()
    for i in range(20): step(i)
    t1=time.perf_counter()
    torch.cuda.synchronize()
    t2=time.perf_counter()
    print(f"host enqueue of 20 steps {1e6*(t1-t0):.0f} us ({1e6*(t1-t0)/20:.1f} per step), until done {1e6*(t2-t0):.0f} us")
# split
t0=time.perf_counter()
for i in range(20): ctxs[i%N].reset_state(0)
t1=time.perf_counter()
torch.cuda.synchronize()
print(f"reset_state alone: {1e6*(t1-t0)/20:.1f} us per call")
