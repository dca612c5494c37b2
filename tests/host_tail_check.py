"""Shared by the host-tail tests: yk_palette_* and yk_chunk_* of the product library against what
the reference wrote — PaletteCompressor's bytes (captured by oracle/ref_harness.cpp's interposer) and the raw MIPM / GTIL /
1DTL / PLNT chunks — from golden fixtures or a live run of the compiled reference."""
import ctypes as C
import os
import struct
import zlib

import numpy as np

from yaik_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libyaikref.so")
PASS_ORDER = capi.PASS_ORDER


# ---- compressor callbacks -----------------------------------------------------------------------------------------
def zlib_callback():
    """Any compressor can sit behind the chunk writers; the tests that only look at layout use zlib."""
    def fn(user, dst, cap, src, n, level):
        data = zlib.compress(C.string_at(src, n), 6)
        if len(data) > cap:
            return 0
        C.memmove(dst, data, len(data))
        return len(data)
    return capi.COMPRESS_FN(fn)


_ref = None


def ref_zstd():
    """ZSTD 1.3.4 as the reference links it (oracle/_ref/libyaikref.so, built from the reference's vendored sources)."""
    global _ref
    if _ref is None and os.path.exists(REF_LIB):
        L = C.CDLL(REF_LIB)
        L.ZSTD_compress.restype = C.c_size_t
        L.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.ZSTD_decompress.restype = C.c_size_t
        L.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ZSTD_isError.restype = C.c_uint
        L.ZSTD_isError.argtypes = [C.c_size_t]
        _ref = L
    return _ref


def ref_zstd_callback():
    L = ref_zstd()

    def fn(user, dst, cap, src, n, level):
        r = L.ZSTD_compress(dst, cap, src, n, level)
        return 0 if L.ZSTD_isError(r) else r
    return capi.COMPRESS_FN(fn)


def unzstd(data: bytes, cap: int) -> bytes:
    L = ref_zstd()
    out = C.create_string_buffer(max(1, cap))
    r = L.ZSTD_decompress(out, cap, data, len(data))
    assert not L.ZSTD_isError(r)
    return out.raw[:r]


# ---- PaletteCompressor --------------------------------------------------------------------------------------------
def palette_compress(lib, pal, rgb: np.ndarray) -> np.ndarray:
    out = np.zeros(rgb.size * 3 + 16, np.uint8)
    n = C.c_int()
    rc = lib.yk_palette_compress(pal, rgb.ctypes.data_as(C.c_void_p), int(rgb.size), out.ctypes.data_as(C.c_void_p), int(out.size), C.byref(n))
    assert rc == 0, rc
    return out[:n.value].copy()


def check_palette_sequence(lib, rec):
    """rec: records of one reference process (7 passes in Convert()'s order).  A fresh bug-compatible object fed the same
    rgbStreams in the same order must write the reference's bytes, stale code-book indices included."""
    pal = lib.yk_palette_create(0)
    try:
        calls = 0
        for k in range(7):
            rgb = np.ascontiguousarray(rec[f"grad{k}.rgb"], dtype=np.uint8)
            want = np.asarray(rec[f"grad{k}.pal"], dtype=np.uint8)
            if int(rec[f"grad{k}.tiledone"][1]) == 0:        # no chunk: the reference did not call PaletteCompressor
                assert want.size == 0
                continue
            got = palette_compress(lib, pal, rgb)
            assert np.array_equal(got, want), (k, got.size, want.size, int(np.argmax(got[:min(got.size, want.size)] != want[:min(got.size, want.size)])))
            calls += 1
        return calls
    finally:
        lib.yk_palette_destroy(pal)


def palette_decompress(data: np.ndarray, out_size: int, color_compression=250):
    """Python restatement of the reference decoder's PaletteDecompressor (decoder/YAIK_GenericFunctions.cpp:128-241);
    returns (6-bit colours before the range remapping, remapped colours) or raises on a stream it would reject or on an
    index outside the code book the stream carries."""
    d = [int(x) for x in data]
    n_code = d[0]
    book = d[1:1 + 3 * n_code]
    pos = 1 + 3 * n_code
    out = d[pos:pos + 3]
    pos += 3
    last = 0
    while len(out) < out_size:
        c = d[pos]; pos += 1
        if c & 0x80:
            if c & 0x40:
                last = len(out) - 3 * ((c & 0x3F) + 2)
                assert last >= 0
                continue
            kind = (c >> 3) & 7
            cur = []
            for comp in range(3):
                if kind == 0:
                    v = out[last + comp]
                    if c & (1 << comp):
                        v = (v + d[pos]) & 255; pos += 1
                elif kind == 1:
                    if c & (1 << comp):
                        v = d[pos]; pos += 1
                    else:
                        v = out[last + comp]
                else:
                    raise AssertionError("reserved code")
                cur.append(v)
        else:
            idx = c & 0x7F
            assert idx < n_code, ("code book index outside the book", idx, n_code)
            cur = [(out[last + comp] + book[3 * idx + comp]) & 255 for comp in range(3)]
        last = len(out)
        out += cur
    assert pos == len(d), ("trailing bytes", pos, len(d))
    inv = (255 << 16) // color_compression
    return np.array(out, np.uint8), np.array([(v * inv) >> 16 for v in out], np.uint8)


# ---- chunks -------------------------------------------------------------------------------------------------------
def _call_chunk(fn, *args, cap=1 << 24):
    buf = np.zeros(cap, np.uint8)
    n = C.c_size_t()
    rc = fn(buf.ctypes.data_as(C.c_void_p), cap, C.byref(n), *args)
    assert rc == 0, rc
    return buf[:n.value].tobytes()


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def gtil_chunk(lib, pal, cb, sx, sy, res, color_compression=250, plane_bits=7):
    """res: what yk_gradient_pass / the oracle returned for the pass (bitmap, rgb, bbox = minX, minY, maxX, maxY)."""
    bbox = (C.c_int * 4)(*res["bbox"])
    bm = np.ascontiguousarray(res["bitmap"], np.uint8); rgb = np.ascontiguousarray(res["rgb"], np.uint8)
    return _call_chunk(lib.yk_chunk_gtil, pal, cb, None, sx, sy, plane_bits, bbox, _p(bm), int(bm.size), _p(rgb), int(rgb.size), color_compression)


def mipm_chunk(lib, alpha):
    if not alpha["wrote"]:
        return b""
    bbox = (C.c_int * 4)(*alpha["chunk_bbox"])
    bm = np.ascontiguousarray(alpha["bitmap"], np.uint8)
    return _call_chunk(lib.yk_chunk_mipm, bbox, _p(bm), int(bm.size))


def tile1d_chunk(lib, cb, planes_r2):
    idx = np.ascontiguousarray(np.concatenate([p["idx"] for p in planes_r2]), np.uint8)
    typ = np.ascontiguousarray(np.concatenate([p["type"] for p in planes_r2]), np.uint8)
    return _call_chunk(lib.yk_chunk_1dtl, cb, None, _p(idx), int(idx.size), _p(typ), int(typ.size), 255, 15)


def plnt_chunk(lib, cb, r1, plane_type=0, half=(0, 0)):
    cons = (C.c_int * 4)(*r1["constraint"])
    defs = np.ascontiguousarray(r1["defs"], np.uint16); nib = np.ascontiguousarray(r1["nibbles"], np.uint8)
    return _call_chunk(lib.yk_chunk_plnt, cb, None, cons, _p(defs), int(defs.size), _p(nib), int(r1["n_nibbles"]), plane_type, half[0], half[1])


def file_header(lib, w, h, has_alpha):
    return _call_chunk(lib.yk_chunk_file_header, w, h, int(has_alpha))


def end_tag(lib):
    return _call_chunk(lib.yk_chunk_end)


# bytes the reference never initialises, by chunk tag: offsets inside the chunk (header base included)
GARBAGE = {b"GTIL": [8 + 25], b"MIPM": [8 + 8, 8 + 9, 8 + 10, 8 + 11, 8 + 14, 8 + 15], b"PLNT": [8 + 22, 8 + 23], b"1DTL": [8 + 19]}


def masked(chunk: bytes) -> bytes:
    b = bytearray(chunk)
    for off in GARBAGE.get(bytes(b[:4]), []):
        if off < len(b):
            b[off] = 0
    return bytes(b)


def parse_gtil(chunk: bytes):
    tag, length = chunk[:4], struct.unpack_from("<I", chunk, 4)[0]
    x, y, w, h, zb, zr, cust, unc, cc, ver, fmt, plane = struct.unpack_from("<hhhhIIIIBBBB", chunk, 8)
    return dict(tag=tag, length=length, bbox=[x, y, w, h], zbitmap=zb, zrgb=zr, custom=cust, uncompressed=unc, colorCompression=cc, format=fmt, plane=plane,
                payload_bitmap=chunk[36:36 + zb], payload_rgb=chunk[36 + zb:36 + zb + zr])


def parse_1dtl(chunk: bytes):
    zpix, upix, ztype, utype, cc, cr, ver = struct.unpack_from("<IIIIBBB", chunk, 8)
    return dict(tag=chunk[:4], length=struct.unpack_from("<I", chunk, 4)[0], zpix=zpix, upix=upix, ztype=ztype, utype=utype, color=cc, range=cr, version=ver,
                payload_type=chunk[28:28 + ztype], payload_pix=chunk[28 + ztype:28 + ztype + zpix])


def parse_plnt(chunk: bytes):
    x, y, w, h, zmap, zstream, expected, ver, fmt = struct.unpack_from("<hhhhIIIBB", chunk, 8)
    return dict(tag=chunk[:4], length=struct.unpack_from("<I", chunk, 4)[0], bbox=[x, y, w, h], zmap=zmap, zstream=zstream, expected=expected, version=ver, format=fmt,
                payload_map=chunk[32:32 + zmap], payload_stream=chunk[32 + zmap:32 + zmap + zstream])


def parse_mipm(chunk: bytes):
    x, y, w, h, ssz, ver, lvl = struct.unpack_from("<hhhhIBB", chunk, 8)
    return dict(tag=chunk[:4], length=struct.unpack_from("<I", chunk, 4)[0], bbox=[x, y, w, h], version=ver, level=lvl, bitmap=chunk[24:24 + (w * h + 7) // 8])


# ---- a whole .yaik stream and the reference decoder -----------------------------------------------------------------
DEC_BIN = os.path.join(ROOT, "oracle", "_ref", "yaik_dec")


def build_yaik(lib, cb, w, h, passes, r2, palette_mode=0):
    """FileHeader + the GTIL chunks of the seven passes + 1DTL + end tag, as Convert() orders them (EC.cpp:9007-9016,
    9057-9093, 9451-9465, 9779-9782) for an image without alpha.  passes / r2: as Context.gradient_pass / range1d return."""
    pal = lib.yk_palette_create(palette_mode)
    try:
        chunks = [file_header(lib, w, h, False)]
        for (sx, sy), res in zip(PASS_ORDER, passes):
            chunks.append(gtil_chunk(lib, pal, cb, sx, sy, res))
        chunks.append(tile1d_chunk(lib, cb, r2))
        chunks.append(end_tag(lib))
        return chunks
    finally:
        lib.yk_palette_destroy(pal)


def decode_yaik(blob: bytes) -> np.ndarray:
    """Run the unmodified reference decoder (oracle/_ref/yaik_dec) on a stream; returns [h][w][channels] uint8."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "a.yaik"), os.path.join(td, "a.raw")
        with open(fin, "wb") as f:
            f.write(blob)
        p = subprocess.run([DEC_BIN, fin, fout], capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-500:]
        raw = open(fout, "rb").read()
    w, h, ch = struct.unpack_from("<iii", raw, 0)
    return np.frombuffer(raw, np.uint8, offset=12).reshape(h, w, ch).copy()
