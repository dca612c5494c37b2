"""Shared by the strip tests: merged strip results against the oracle run on the whole image."""
import numpy as np

from oracle_py import Oracle, PASS_ORDER
from parity import _eq


def check_against_oracle(merged, planes, r2=True):
    o = Oracle(planes)
    if planes.shape[0] == 4 and merged.get("alpha") is not None:     # the alpha stage of the strip set (merged on the host)
        want = o.alpha()
        got = merged["alpha"]
        if want is not None:
            assert got["bound"] == want["bound"] and got["remaining"] == want["remaining"] and got["wrote"] == want["wrote"]
            assert got["chunk_bbox"] == want["chunk_bbox"]
            _eq(got["bitmap"], want["bitmap"], "alpha bitmap of the strip set")
    for k, (sx, sy) in enumerate(PASS_ORDER):
        want = o.gradient_pass(sx, sy)
        got = merged["passes"][k]
        assert got["tiledone"] == want["tiledone"], (f"pass {k} tileDone", got["tiledone"], want["tiledone"])
        _eq(got["bitmap"], want["bitmap"], f"pass {k} bitmap")
        assert list(got["bbox"]) == list(want["bbox"]), (f"pass {k} bbox", got["bbox"], want["bbox"])
        _eq(got["rgb"], want["rgb"], f"pass {k} rgbStream")
    if r2:
        for n in range(3):
            want = o.range1d(n)
            _eq(merged["r2"][n]["type"], want["type"], f"R2 type plane {n}")
            _eq(merged["r2"][n]["idx"], want["idx"], f"R2 idx plane {n}")
    o.close()
