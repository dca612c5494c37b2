"""Developer tool: run the bench texture through the -DYK_TIMING build and print where producer / consumer cycles go."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from yaik_b200 import capi
from yaik_b200.synth import make_image, SEED_BASE
lib = capi.load_library(os.path.join(ROOT, 'yaik_b200', 'csrc', 'libyaik_b200_timing.so'))
ctx = capi.Context(2048, 2048, planes=4, slots=1, lib=lib)
img = make_image(2048, 2048, 4, SEED_BASE + 1)
ctx.set_image(img, 0)
st = capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D
for _ in range(3):
    ctx.reset_state(0); ctx.analyze(st); ctx.sync()
out = (C.c_ulonglong * 32)()
lib.yk_debug_timing(out, 1)
N = 5
for _ in range(N):
    ctx.reset_state(0); ctx.analyze(st); ctx.sync()
lib.yk_debug_timing(out, 0)
names = {0: "prod: lookahead+free wait", 6: "prod: shfl ticket", 7: "prod: issue ticket atomic", 3: "prod: decode", 2: "prod: fence + expect_tx + 4x TMA issue",
         8: "cons(w0): queue+wait raw", 9: "cons(w0): pack", 11: "cons(w0):   cascade passes (incl. pretest)", 12: "cons(w0):   cells + touch + latRGB", 13: "cons(w0):   range stage", 10: "cons(w0): rest of the item (incl. raw 16x16 pass)"}
print(f'all consumer warps, cycles per warp per launch: wait for the first item {out[14] / N / 3404:.0f}, waits between items {out[5] / N / 3404:.0f} ({out[4] / N / 3404:.1f} items), wait at the end {out[15] / N / 3404:.0f}')
for i, n in names.items():
    print(f"{n:34s} {out[i] / N / 148:12.0f} cycles per CTA per launch")

# wall-clock shape of the launches (globaltimer, ns; low 32 bits summed, so differences of averages are exact enough over 5 launches)
# only meaningful for the LAST launch alone: rerun one launch
lib.yk_debug_timing(out, 1)
ctx.reset_state(0); ctx.sync(); ctx.analyze(st); ctx.sync()
lib.yk_debug_timing(out, 0)
t0 = (~out[16]) & 0xFFFFFFFFFFFFFFFF
lo = t0 & 0xFFFFFFFF
def avg(s, n): return (s / max(n, 1)) - lo
print(f"one launch: CTA set-up done avg {avg(out[17], out[23]) / 1e3:7.2f} us after the first CTA;  consumer warps: first item avg {avg(out[18], out[19]) / 1e3:7.2f} us, exit avg {avg(out[20], out[21]) / 1e3:7.2f} us, last exit {(out[22] - t0) / 1e3:7.2f} us  ({out[19]} warps with work, {out[21]} warps)")
n = out[23]
print("producer: loads of unit 0 / 1 / 2 issued avg " + " / ".join(f"{avg(out[24 + k], n) / 1e3:6.2f}" for k in range(3)) + " us;  first item of unit 0 / 1 / 2 taken avg " + " / ".join(f"{avg(out[27 + k], n) / 1e3:6.2f}" for k in range(3)) + " us")
te = (~out[30]) & 0xFFFFFFFFFFFFFFFF
print(f"kernel entry: first CTA {(te - t0) / 1e3:7.2f} us, avg {(out[31] / max(n, 1) - (t0 & 0xFFFFFFFF)) / 1e3:7.2f} us relative to the first CTA's set-up done")
import torch
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
