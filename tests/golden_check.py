"""Compare an implementation (the C oracle, or a yaik_b200 Context) with a golden fixture made from the reference."""
import glob
import hashlib
import os

import numpy as np

from oracle_py import PASS_ORDER

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixtures():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return g, g["input"].astype(np.int32), tuple(str(s) for s in g["stages"])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(np.asarray(a, dtype=np.int32)).tobytes()).hexdigest()


def chroma_args(stages):
    """(cfg, modes) of a "chroma=XYXY:MM" stage, or None."""
    for s in stages:
        if s.startswith("chroma="):
            return tuple(int(ch) for ch in s[7:11]), tuple(int(ch) for ch in s[12:14])
    return None


def check(g, stages, *, alpha, gradient_pass, range1d, range_dyn, state=None, chroma=None):
    """The callables return dicts shaped like oracle_py.Oracle's methods."""
    if "alpha" in stages:
        a = alpha()
        assert a["bound"] == list(g["alpha.bound"])
        assert a["remaining"] == int(g["alpha.remaining"][0])
        assert np.array_equal(a["bitmap"], g["alpha.bitmap"])
        assert a["chunk_bbox"] == list(g["alpha.chunk_bbox"][:4])
    if "grad" in stages:
        for k, (sx, sy) in enumerate(PASS_ORDER):
            r = gradient_pass(sx, sy)
            assert r["tiledone"] == int(g[f"grad{k}.tiledone"][0]), (k, "TileDone")
            assert np.array_equal(r["rgb"], g[f"grad{k}.rgb"]), (k, "rgbStream")
            if int(g[f"grad{k}.tiledone"][1]):          # the reference wrote a GTIL chunk: bitmap + header bbox recoverable
                assert np.array_equal(r["bitmap"], g[f"grad{k}.bitmap"]), (k, "bitmap")
                mnx, mny, mxx, mxy = r["bbox"]
                assert [mnx, mny, mxx - mnx, mxy - mnx] == list(g[f"grad{k}.bbox"]), (k, "bbox")   # EC.cpp:4258 (maxY - minX)
        if state is not None:
            st = state()
            assert sha(st["smoothMap"]) == str(g["sha256:state.smoothMap"])
            assert sha(st["mipmapMask"]) == str(g["sha256:state.mipmapMask"])
            for n in range(3):
                assert sha(st["mapSmoothTile"][n]) == str(g[f"sha256:state.mapSmoothTile{n}"])
                assert sha(st["mappedRGB"][n]) == str(g[f"sha256:state.mappedRGB{n}"])
                assert sha(st["recon"][n]) == str(g[f"sha256:state.recon{n}"])
    if "r2" in stages:
        for n in range(3):
            r = range1d(n)
            assert np.array_equal(r["idx"], g[f"r2.idx{n}"]), n
            assert np.array_equal(r["type"], g[f"r2.type{n}"]), n
    if "r1" in stages or "r1_3bit" in stages:
        for n in range(3):
            r = range_dyn(n, "r1_3bit" in stages)
            assert np.array_equal(r["defs"], g[f"r1.defs{n}"]), n
            assert np.array_equal(r["nibbles"], g[f"r1.nibbles{n}"]), n
            assert r["constraint"] == list(g[f"r1.hdr{n}"][:4])
            assert sha(r["dst"]) == str(g[f"sha256:r1.dst{n}"])
    ca = chroma_args(stages)
    if ca is not None:
        assert chroma is not None, "fixture holds a chroma stage"
        r = chroma(*ca)
        for k in ("Y", "workCo", "workCg"):
            assert sha(r[k]) == str(g["sha256:yc." + k]), k
        for n in range(3):
            d = r["coded"][n]
            assert d["constraint"] == list(g[f"yc.hdr{n}"][:4]), n
            assert np.array_equal(d["defs"], g[f"yc.defs{n}"]), n
            assert np.array_equal(d["nibbles"], g[f"yc.nibbles{n}"]), n
            assert sha(d["dst"]) == str(g[f"sha256:yc.dst{n}"]), n
