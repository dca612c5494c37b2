"""The host-side C++ mirror of the reference's EncoderContext (yaik_b200/host) driven the way Convert() drives the
reference, compared with the golden vectors of the reference.  CPU: the mirror linked against the emulated library
(logic of the mirror itself); -m gpu: the real binary on the CUDA library."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import golden_check
from refrun import parse_records
from yaik_b200.synth import to_ykin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_BIN = os.path.join(ROOT, "tests", "emu", "_build", "host_mirror_test_emu")
GPU_BIN = os.path.join(ROOT, "yaik_b200", "host", "_build", "host_mirror_test")


def run_mirror(binary, planes, stages, extra=()):
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.ykin"), os.path.join(td, "out.ykout")
        open(fin, "wb").write(to_ykin(planes))
        subprocess.run([binary, fin, fout, *stages, *extra], check=True, timeout=600)
        return parse_records(open(fout, "rb").read())


def compare(name, binary, extra=()):
    g, planes, stages = golden_check.load(name)
    r = run_mirror(binary, planes, stages, extra)
    h, w = planes.shape[1:]

    def grad(sx, sy):
        k = golden_check.PASS_ORDER.index((sx, sy))
        bb = list(r[f"grad{k}.bbox"])
        return dict(tiledone=int(r[f"grad{k}.tiledone"][0]), rgb=r[f"grad{k}.rgb"], bitmap=r[f"grad{k}.bitmap"],
                    bbox=[bb[0], bb[1], bb[0] + bb[2], bb[0] + bb[3]])         # undo the header form (maxY - minX, EC.cpp:4258)
    golden_check.check(
        g, stages,
        alpha=lambda: dict(bound=list(r["alpha.bound"]), remaining=int(r["alpha.remaining"][0]), bitmap=r["alpha.bitmap"],
                           chunk_bbox=list(r["alpha.chunk_bbox"][:4])),
        gradient_pass=grad,
        range1d=lambda n: dict(idx=r[f"r2.idx{n}"], type=r[f"r2.type{n}"]),
        range_dyn=lambda n, m3: dict(defs=r[f"r1.defs{n}"], nibbles=r[f"r1.nibbles{n}"], constraint=list(r[f"r1.hdr{n}"][:4]), dst=r[f"r1.dst{n}"]),
        chroma=lambda cfg, modes: dict(Y=r["yc.Y"], workCo=r["yc.workCo"], workCg=r["yc.workCg"],
                                       coded=[dict(defs=r[f"yc.defs{n}"], nibbles=r[f"yc.nibbles{n}"], constraint=list(r[f"yc.hdr{n}"][:4]),
                                                   dst=r[f"yc.dst{n}"]) for n in range(3)]),
        state=lambda: dict(smoothMap=r["state.smoothMap"], mipmapMask=r["state.mipmapMask"],
                           mapSmoothTile=[r[f"state.mapSmoothTile{i}"] for i in range(3)],
                           mappedRGB=[r[f"state.mappedRGB{i}"] for i in range(3)], recon=[r[f"state.recon{i}"] for i in range(3)]))
    assert int(r["meta"][3]) == 0


@pytest.fixture(scope="module")
def emu_bin():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu"), "mirror"], check=True)
    return EMU_BIN


@pytest.mark.parametrize("name", ["patchy_72x40", "mip32_rgba", "alpha_island128", "r1_signed96", "chroma_synth128_rgb", "chroma_alpha_island128"])
def test_mirror_logic_on_emulated_library(emu_bin, name):
    compare(name, emu_bin)


def test_mirror_without_prepare_runs_pass_by_pass(emu_bin):
    compare("patchy_72x40", emu_bin, extra=("noprepare",))


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_check.fixtures())
def test_mirror_on_gpu_matches_reference_vectors(name):
    assert os.path.exists(GPU_BIN), "build it: make -C yaik_b200/host"
    compare(name, GPU_BIN)


@pytest.mark.gpu
def test_mirror_on_gpu_pass_by_pass():
    compare("patchy128", GPU_BIN, extra=("noprepare",))


# ---- host tails of the mirror: the chunks it writes into outFile -------------------------------
def _split_chunks(blob):
    """FileHeader (12 bytes), then HeaderBase-framed chunks, then the 4-byte end tag."""
    import struct
    out, off = [blob[:12]], 12
    while off < len(blob) - 4:
        (length,) = struct.unpack_from("<I", blob, off + 4)
        out.append(blob[off:off + 8 + length]); off += 8 + length
    out.append(blob[off:])
    return out


def _mirror_file(binary, name, extra):
    import host_tail_check as H
    g, planes, stages = golden_check.load(name)
    with tempfile.TemporaryDirectory() as td:
        fin, fout, fy = os.path.join(td, "in.ykin"), os.path.join(td, "out.ykout"), os.path.join(td, "out.yaik")
        open(fin, "wb").write(to_ykin(planes))
        subprocess.run([binary, fin, fout, *stages, "yaik=" + fy, *extra], check=True, timeout=600)
        return g, planes, stages, open(fy, "rb").read()


def _check_file_against_reference_chunks(binary, name, extra=()):
    """With the reference's own ZSTD behind the callback the file equals what the reference encoder wrote, chunk for chunk
    (bytes it never initialises masked): FileHeader, MIPM, GTIL x n, 1DTL, PLNT x 3, end tag."""
    import struct
    import host_tail_check as H
    g, planes, stages, blob = _mirror_file(binary, name, ("zstdlib=" + H.REF_LIB, *extra))
    h, w = planes.shape[1:]
    want = [b"YAIK" + struct.pack("<HHHH", 1, w, h, 1 if planes.shape[0] == 4 else 0)]
    if "alpha" in stages and "alpha.chunk" in g.files and g["alpha.chunk"].size:
        want.append(bytes(g["alpha.chunk"]))
    if "grad" in stages:
        want += [bytes(g[f"grad{k}.chunk"]) for k in range(7) if g[f"grad{k}.chunk"].size]
    if "r2" in stages and g["r2.chunk"].size:
        want.append(bytes(g["r2.chunk"]))
    if "r1" in stages or "r1_3bit" in stages:
        want += [bytes(g[f"r1.chunk{c}"]) for c in range(3)]
    want.append(struct.pack("<I", 0xDEADBEEF))
    got = _split_chunks(blob)
    assert [c[:4] for c in got] == [c[:4] for c in want]
    assert [H.masked(c) for c in got] == [H.masked(c) for c in want]


needs_ref_zstd = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libyaikref.so")), reason="oracle/_ref not built")


@needs_ref_zstd
@pytest.mark.parametrize("name", ["patchy_72x40", "mip32_rgba", "alpha_island128", "synth256_rgba"])
@pytest.mark.parametrize("mode", ["sync", "async"])
def test_mirror_writes_the_references_chunks(emu_bin, name, mode):
    _check_file_against_reference_chunks(emu_bin, name, ("async",) if mode == "async" else ())


def test_mirror_tails_on_the_worker_thread_write_the_same_file(emu_bin):
    """Without any zstd (stand-in compressor): the worker-thread form writes byte for byte what the synchronous form writes."""
    a = _mirror_file(emu_bin, "synth256_rgba", ())[3]
    b = _mirror_file(emu_bin, "synth256_rgba", ("async",))[3]
    assert a == b and len(a) > 1000 and a[:4] == b"YAIK" and a[-4:] == bytes([0xEF, 0xBE, 0xAD, 0xDE])


@pytest.mark.gpu
@needs_ref_zstd
@pytest.mark.parametrize("name", ["alpha_island128", "synth256_rgba", "patchy128"])
def test_mirror_on_gpu_writes_the_references_chunks(name):
    _check_file_against_reference_chunks(GPU_BIN, name, ("async",))
