#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled from /root/reference (oracle/_ref/yaik_ref,
built by oracle/Makefile).  Run in the build container only (the GPU box has no /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

Each fixture holds the seeded input (uint8 planes; int16 for the signed R1 case) and every observable result of the reference's stage functions
for it: MipPrefilter (bound, bitmap), the seven FittingQuadSmooth passes (TileDone, bitmap, bbox header,
rgbStream captured at PaletteCompressor), DynamicTileCompressor (idx/type per plane), DynamicTileEncode (tile defs,
nibbles, constraint) and SHA-256 digests of the int32 state planes.  tests/test_golden.py pins the C oracle to them
(and, on the GPU box, tests/test_gpu_parity.py pins the CUDA path to them) without needing the reference."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

import cases  # noqa: E402
from refrun import have_ref, run_ref  # noqa: E402

GOLDEN_CASES = ["ramp64_a2", "patchy128", "patchy_72x40", "noise_delta1", "noise_hi", "mip32_rgba", "mip16_rgba",
                "mip8_rgb", "mip4_rgb", "alpha_island128", "alpha_corner_only", "synth256_rgba", "synth256_rgb_3bit", "r1_signed96"]


# chroma front-end fixtures: indices into cases.CHROMA_CASES (the CLI configuration, quarter size with alpha, one axis each)
CHROMA_GOLDEN = [0, 1, 2, 7]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert have_ref(), "build oracle/_ref first (make -C oracle ref)"
    for name in GOLDEN_CASES:
        planes, stages = cases.SMALL_CASES[name]()
        ref = run_ref(planes, stages)
        in_bytes = int(planes.min()) >= 0 and int(planes.max()) <= 255
        out = {"input": planes.astype(np.uint8 if in_bytes else np.int16), "stages": np.array(list(stages))}
        for k, v in ref.items():
            if k.startswith("time."):
                continue
            if k.startswith("state.") or k.startswith("r2.debug") or k.startswith("r1.dst") or k == "alpha.mask":
                out["sha256:" + k] = np.array(digest(v.astype(np.int32)))
            else:
                out[k] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, planes.shape, sorted(k for k in out if not k.startswith("sha256"))[:6], "...")
    for k in CHROMA_GOLDEN:
        name, pre, cfg, modes = cases.CHROMA_CASES[k]
        planes, _ = cases.SMALL_CASES[name]()
        stages = (*pre, "chroma=%d%d%d%d:%d%d" % (*cfg, *modes))
        ref = run_ref(planes, stages)
        out = {"input": planes.astype(np.uint8), "stages": np.array(list(stages))}
        for key, v in ref.items():
            if key.startswith("time."):
                continue
            if key.startswith("state.") or key == "alpha.mask" or key.startswith("yc.dst") or key in ("yc.Y", "yc.Co", "yc.Cg", "yc.workCo", "yc.workCg"):
                out["sha256:" + key] = np.array(digest(v.astype(np.int32)))
            else:
                out[key] = v
        np.savez_compressed(os.path.join(HERE, "chroma_" + name + ".npz"), **out)
        print("chroma_" + name, planes.shape, stages)


if __name__ == "__main__":
    main()
