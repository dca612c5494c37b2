#!/usr/bin/env python
"""Throughput of the batch shapes of BASELINE.json on one GPU (planes resident in HBM as int32, results left in HBM):
configs[2] one context's share of 256 x 1024x1024 RGBA textures as batch launches over slots, configs[4] RGBA mip chains
4096 -> 4 analysed level by level (alpha for levels >= 16, range stage for levels >= 8).  Prints one JSON line each."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from yaik_b200 import capi
from yaik_b200.synth import make_image, mip_chain, SEED_BASE

lib = capi.load_library()
ST = capi.STAGE_ALPHA | capi.STAGE_GRADIENT | capi.STAGE_RANGE1D

# ---- configs[2]: 32 textures of 1024x1024 RGBA per launch
NS = 32
ctx = capi.Context(1024, 1024, planes=4, slots=NS, lib=lib)
ctx.set_upload_format(False)
imgs = [make_image(1024, 1024, 4, SEED_BASE + 2 + i) for i in range(4)]
for s in range(NS):
    ctx.set_image(imgs[s % 4], s)
def batch():
    ctx.reset_states(0, NS)
    ctx.analyze(ST, slot0=0, n_slots=NS)
for _ in range(3):
    batch()
ctx.sync()
R = 20
t0 = time.perf_counter()
for _ in range(R):
    batch()
ctx.sync()
dt = (time.perf_counter() - t0) / R
print(json.dumps({"workload": f"batch launches of {NS} synthetic 1024x1024 RGBA textures (BASELINE configs[2] shape), one GPU", "value": round(NS * 1.048576 / dt, 1),
                  "unit": "MP/s", "ms_per_batch": round(dt * 1e3, 3), "note": "256 MiB of planes per launch (> L2), wall clock over 20 launches"}), flush=True)
ctx.close()

# ---- configs[4]: one RGBA mip chain 4096 -> 4, level by level
chain = mip_chain(4096, SEED_BASE + 4, 4)
ctxs = []
for lvl in chain:
    c, h, w = lvl.shape
    cx = capi.Context(w, h, planes=4, slots=1, lib=lib)
    cx.set_upload_format(False)
    cx.set_image(lvl, 0)
    ctxs.append((cx, (capi.STAGE_ALPHA if w >= 16 else 0) | capi.STAGE_GRADIENT | (capi.STAGE_RANGE1D if w >= 8 else 0)))
def run_chain():
    for cx, st in ctxs:
        cx.reset_state(0)
        cx.analyze(st)
for _ in range(3):
    run_chain()
for cx, _ in ctxs:
    cx.sync()
R = 20
t0 = time.perf_counter()
for _ in range(R):
    run_chain()
for cx, _ in ctxs:
    cx.sync()
dt = (time.perf_counter() - t0) / R
px = sum(l.shape[1] * l.shape[2] for l in chain)
print(json.dumps({"workload": "one synthetic RGBA mip chain 4096x4096 ... 4x4 (BASELINE configs[4] shape), every level its own context / stream, one GPU",
                  "value": round(px / 1e6 / dt, 1), "unit": "MP/s", "ms_per_chain": round(dt * 1e3, 3), "levels": len(chain)}), flush=True)
for cx, _ in ctxs:
    cx.close()
