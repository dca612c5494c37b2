"""Host tails of the product library on the CPU: PaletteCompressor and the chunk serialisers against
what the unmodified reference wrote (golden fixtures made by tests/golden/make_golden.py; a live run of oracle/_ref when
it is built).  No GPU involved: these entry points are plain host code of libyaik_b200.so."""
import zlib

import numpy as np
import pytest

import golden_check
import host_tail_check as H
from refrun import have_ref, run_ref
from yaik_b200 import capi
from yaik_b200.synth import make_image, SEED_BASE

GRAD_FIXTURES = [n for n in golden_check.fixtures() if not n.startswith("chroma_")]


@pytest.fixture(scope="module")
def lib():
    from yaik_b200 import build as ykbuild
    return capi.load_library(ykbuild.build())


def _grad(g, name):
    return "grad" in tuple(str(s) for s in g["stages"])


@pytest.mark.parametrize("name", GRAD_FIXTURES)
def test_palette_compressor_writes_the_references_bytes(lib, name):
    g, planes, stages = golden_check.load(name)
    if "grad" not in stages:
        pytest.skip("no gradient stage in this fixture")
    H.check_palette_sequence(lib, g)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("w,h,ch,seed", [(512, 512, 3, 0), (384, 256, 4, 21), (128, 128, 3, 5), (1024, 512, 4, 2)])
def test_palette_compressor_against_a_live_reference_run(lib, w, h, ch, seed):
    rec = run_ref(make_image(w, h, ch, SEED_BASE + seed), ("grad",))
    assert H.check_palette_sequence(lib, rec) >= 5


@pytest.mark.parametrize("name", GRAD_FIXTURES)
def test_decodable_mode_only_uses_its_own_code_book(lib, name):
    """The reference's decoder reads what YK_PALETTE_DECODABLE writes back to the rgbStream (remapped to 0..255), every
    index inside the code book of the stream; the bug-compatible bytes may not decode but never differ in
    length class: both modes code the same colours."""
    g, planes, stages = golden_check.load(name)
    if "grad" not in stages:
        pytest.skip("no gradient stage in this fixture")
    pal = lib.yk_palette_create(1)
    try:
        for k in range(7):
            rgb = np.ascontiguousarray(g[f"grad{k}.rgb"], np.uint8)
            if rgb.size < 3:
                continue
            raw, remapped = H.palette_decompress(H.palette_compress(lib, pal, rgb), rgb.size)
            assert np.array_equal(raw, rgb)
            assert np.array_equal(remapped, ((rgb.astype(np.int64) * ((255 << 16) // 250)) >> 16).astype(np.uint8))
    finally:
        lib.yk_palette_destroy(pal)


def test_bug_compatible_mode_reproduces_the_stale_code_book(lib):
    """SURVEY.md S10 on a sequence built to trigger it: a large code book, then a stream whose own book is small but whose
    deltas exist among the stale entries.  Bug-compatible: an index past the new book is emitted (what the reference does);
    decodable: never."""
    rng = np.random.default_rng(7)
    big = rng.integers(0, 64, size=3 * 400, dtype=np.uint8)
    small = np.tile(np.array([10, 10, 10, 10, 10, 10], np.uint8), 4)
    probe = np.concatenate([small, big[:6], small[:6]]).astype(np.uint8)
    outs = {}
    for mode in (0, 1):
        pal = lib.yk_palette_create(mode)
        H.palette_compress(lib, pal, big)
        outs[mode] = H.palette_compress(lib, pal, probe)
        lib.yk_palette_destroy(pal)
    H.palette_decompress(outs[1], probe.size)                      # decodable: reads back
    n_code = int(outs[0][0])
    codes = outs[0][1 + 3 * n_code + 3:]
    stale = [int(c) for c in codes if c < 0x80 and c >= n_code]
    if stale:                                                       # then the decoder's own rule rejects / misreads it
        with pytest.raises(AssertionError):
            H.palette_decompress(outs[0], probe.size)
    # a fresh object forgets: the same stream alone codes like the decodable mode's
    fresh = lib.yk_palette_create(0)
    alone = H.palette_compress(lib, fresh, probe)
    lib.yk_palette_destroy(fresh)
    H.palette_decompress(alone, probe.size)


def test_palette_capacity_and_arguments(lib):
    pal = lib.yk_palette_create(0)
    rgb = np.arange(30, dtype=np.uint8)
    out = np.zeros(4, np.uint8)
    import ctypes as C
    n = C.c_int()
    assert lib.yk_palette_compress(pal, rgb.ctypes.data, 30, out.ctypes.data, 4, C.byref(n)) == -3      # YK_ERR_CAPACITY
    assert lib.yk_palette_compress(pal, rgb.ctypes.data, 2, out.ctypes.data, 4, C.byref(n)) == -2       # YK_ERR_ARG
    assert lib.yk_palette_create(5) is None
    lib.yk_palette_destroy(pal)


# ---- chunks ---------------------------------------------------------------------------------------------------------
def _pass_results(g):
    out = []
    for k in range(7):
        x, y, w, hb = [int(v) for v in g[f"grad{k}.bbox"]]
        wrote = int(g[f"grad{k}.tiledone"][1])
        # header form -> minX, minY, maxX, maxY (bbox.h = maxY - minX in the reference, EC.cpp:4258)
        bbox = [x, y, x + w, hb + x] if wrote else [0, 0, 0, 0]
        out.append(dict(bitmap=np.asarray(g[f"grad{k}.bitmap"], np.uint8), rgb=np.asarray(g[f"grad{k}.rgb"], np.uint8), bbox=bbox, wrote=wrote))
    return out


@pytest.mark.parametrize("name", GRAD_FIXTURES)
def test_chunk_layouts_against_the_references_chunks(lib, name):
    """Named header fields equal the reference's; the payloads (here behind zlib: any compressor can sit behind the callback)
    decompress to the streams handed in; chunk lengths are padded to 4 bytes."""
    g, planes, stages = golden_check.load(name)
    cb = H.zlib_callback()
    if "grad" in stages:
        pal = lib.yk_palette_create(0)
        for k, ((sx, sy), res) in enumerate(zip(capi.PASS_ORDER, _pass_results(g))):
            ref = bytes(np.asarray(g[f"grad{k}.chunk"], np.uint8))
            mine = H.gtil_chunk(lib, pal, cb, sx, sy, res) if res["wrote"] else b""
            assert (len(mine) > 0) == (len(ref) > 0), k
            if not ref:
                continue
            a, b = H.parse_gtil(mine), H.parse_gtil(ref)
            for f in ("tag", "bbox", "custom", "uncompressed", "colorCompression", "format", "plane"):
                assert a[f] == b[f], (k, f, a[f], b[f])
            assert a["length"] % 4 == 0 and a["length"] == len(mine) - 8 and a["length"] >= 28 + a["zbitmap"] + a["zrgb"]
            assert zlib.decompress(a["payload_bitmap"]) == res["bitmap"].tobytes()
            assert zlib.decompress(a["payload_rgb"]) == bytes(np.asarray(g[f"grad{k}.pal"], np.uint8))
        lib.yk_palette_destroy(pal)
    if "alpha" in stages and "alpha.chunk" in g.files:
        ref = bytes(np.asarray(g["alpha.chunk"], np.uint8))
        cbx = [int(v) for v in g["alpha.chunk_bbox"][:4]]
        mine = H.mipm_chunk(lib, dict(wrote=1, chunk_bbox=cbx, bitmap=np.asarray(g["alpha.bitmap"], np.uint8)))
        assert H.masked(mine) == H.masked(ref)
    if "r2" in stages and "r2.chunk" in g.files:
        ref = bytes(np.asarray(g["r2.chunk"], np.uint8))
        r2 = [dict(idx=np.asarray(g[f"r2.idx{c}"], np.uint8), type=np.asarray(g[f"r2.type{c}"], np.uint8)) for c in range(3)]
        mine = H.tile1d_chunk(lib, cb, r2)
        assert (len(mine) > 0) == (len(ref) > 0)
        if ref:
            a, b = H.parse_1dtl(mine), H.parse_1dtl(ref)
            for f in ("tag", "upix", "utype", "color", "range", "version"):
                assert a[f] == b[f], (f, a[f], b[f])
            assert zlib.decompress(a["payload_pix"]) == b"".join(p["idx"].tobytes() for p in r2)
            assert zlib.decompress(a["payload_type"]) == b"".join(p["type"].tobytes() for p in r2)
    if ("r1" in stages or "r1_3bit" in stages) and "r1.chunk0" in g.files:
        for c in range(3):
            ref = bytes(np.asarray(g[f"r1.chunk{c}"], np.uint8))
            hdr = [int(v) for v in g[f"r1.hdr{c}"]]
            defs = np.asarray(g[f"r1.defs{c}"], np.uint16); nib = np.asarray(g[f"r1.nibbles{c}"], np.uint8)
            # the reference's stream always holds whole bytes: an odd nibble count is closed (EC.cpp:4524-4526)
            mine = H.plnt_chunk(lib, cb, dict(constraint=hdr[:4], defs=defs, nibbles=nib, n_nibbles=2 * nib.size))
            a, b = H.parse_plnt(mine), H.parse_plnt(ref)
            for f in ("tag", "bbox", "expected", "version", "format"):
                assert a[f] == b[f], (c, f, a[f], b[f])
            assert zlib.decompress(a["payload_map"]) == defs.tobytes() and zlib.decompress(a["payload_stream"]) == nib.tobytes()


@pytest.mark.skipif(H.ref_zstd() is None, reason="oracle/_ref/libyaikref.so not built (no /root/reference)")
@pytest.mark.parametrize("name", GRAD_FIXTURES)
def test_chunks_are_byte_identical_behind_the_references_zstd(lib, name):
    """With the reference's own ZSTD 1.3.4 behind the callback every chunk equals the reference's byte for byte, apart
    from the bytes the reference never initialises (HeaderGradientTile::version, MipmapHeader::streamSize, padding)."""
    g, planes, stages = golden_check.load(name)
    cb = H.ref_zstd_callback()
    if "grad" in stages:
        pal = lib.yk_palette_create(0)
        for k, ((sx, sy), res) in enumerate(zip(capi.PASS_ORDER, _pass_results(g))):
            ref = bytes(np.asarray(g[f"grad{k}.chunk"], np.uint8))
            if res["wrote"]:
                assert H.masked(H.gtil_chunk(lib, pal, cb, sx, sy, res)) == H.masked(ref), k
        lib.yk_palette_destroy(pal)
    if "r2" in stages and "r2.chunk" in g.files:
        r2 = [dict(idx=np.asarray(g[f"r2.idx{c}"], np.uint8), type=np.asarray(g[f"r2.type{c}"], np.uint8)) for c in range(3)]
        assert H.masked(H.tile1d_chunk(lib, cb, r2)) == H.masked(bytes(np.asarray(g["r2.chunk"], np.uint8)))
    if ("r1" in stages or "r1_3bit" in stages) and "r1.chunk0" in g.files:
        for c in range(3):
            hdr = [int(v) for v in g[f"r1.hdr{c}"]]
            defs = np.asarray(g[f"r1.defs{c}"], np.uint16); nib = np.asarray(g[f"r1.nibbles{c}"], np.uint8)
            mine = H.plnt_chunk(lib, cb, dict(constraint=hdr[:4], defs=defs, nibbles=nib, n_nibbles=2 * nib.size))
            assert H.masked(mine) == H.masked(bytes(np.asarray(g[f"r1.chunk{c}"], np.uint8))), c


def test_file_header_and_end_tag(lib):
    assert H.file_header(lib, 2048, 1024, True) == b"YAIK" + bytes([1, 0, 0, 8, 0, 4, 1, 0])
    assert H.end_tag(lib) == bytes([0xEF, 0xBE, 0xAD, 0xDE])
