"""TEST INFRASTRUCTURE: run the compiled, unmodified reference (oracle/_ref/yaik_ref) on an image and
parse the named records it dumps (see oracle/ref_harness.cpp).  Never imported by the product."""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "yaik_ref")

_DT = {"i": "<i4", "B": "u1", "H": "<u2", "d": "<f8"}


def have_ref() -> bool:
    return os.path.exists(REF_BIN)


def parse_records(blob: bytes) -> dict:
    out, off = {}, 0
    while off < len(blob):
        name = blob[off:off + 32].split(b"\0", 1)[0].decode()
        dt = chr(blob[off + 32])
        (cnt,) = struct.unpack_from("<Q", blob, off + 33)
        off += 41
        a = np.frombuffer(blob, dtype=_DT[dt], count=cnt, offset=off).copy()
        off += a.nbytes
        out[name] = a
    return out


def run_ref(planes: np.ndarray, stages=("alpha", "grad", "r2"), reps: int = 1, timeout: int = 600) -> dict:
    from yaik_b200.synth import to_ykin
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.ykin"), os.path.join(td, "out.ykout")
        with open(fin, "wb") as f:
            f.write(to_ykin(planes))
        cmd = [REF_BIN, fin, fout, *stages]
        if reps > 1:
            cmd.append(f"reps={reps}")
        subprocess.run(cmd, check=True, timeout=timeout, stdout=subprocess.DEVNULL)
        with open(fout, "rb") as f:
            return parse_records(f.read())
