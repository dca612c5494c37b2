"""Decoder-side round trip with the UNMODIFIED reference decoder library (decoder/*.cpp compiled
into oracle/_ref by oracle/Makefile, driven through its public API by oracle/dec_harness.cpp): the analysis streams,
serialised by the product's own chunk writers and PaletteCompressor, are a .yaik stream the reference decodes."""
import os
import struct

import numpy as np
import pytest

import host_tail_check as H
from oracle_py import Oracle, PASS_ORDER
from refrun import have_ref, run_ref
from yaik_b200 import capi
from yaik_b200.synth import make_image, SEED_BASE

needs_decoder = pytest.mark.skipif(not (os.path.exists(H.DEC_BIN) and H.ref_zstd() is not None), reason="oracle/_ref decoder not built (no /root/reference)")


def _oracle_streams(planes):
    o = Oracle(planes)
    passes = [o.gradient_pass(sx, sy) for sx, sy in PASS_ORDER]
    r2 = [o.range1d(p) for p in range(3)]
    o.close()
    return passes, r2


@needs_decoder
@pytest.mark.skipif(not have_ref(), reason="oracle/_ref encoder not built")
@pytest.mark.parametrize("w,h,seed", [(256, 256, 12), (512, 256, 0), (128, 192, 5)])
def test_file_equals_the_reference_encoders_and_decodes_alike(w, h, seed):
    """Bug-compatible PaletteCompressor + the reference's ZSTD behind the callback: the stream equals the chunks the
    reference encoder wrote, byte for byte (uninitialised bytes masked), and so does the image its decoder makes of it."""
    from yaik_b200 import build as ykbuild
    lib = capi.load_library(ykbuild.build())
    planes = make_image(w, h, 3, SEED_BASE + seed)
    passes, r2 = _oracle_streams(planes)
    mine = H.build_yaik(lib, H.ref_zstd_callback(), w, h, passes, r2, palette_mode=0)
    rec = run_ref(planes, ("grad", "r2"))
    theirs = [b"YAIK" + struct.pack("<HHHH", 1, w, h, 0)] + [bytes(rec[f"grad{k}.chunk"]) for k in range(7)] + [bytes(rec["r2.chunk"]), struct.pack("<I", 0xDEADBEEF)]
    assert [H.masked(c) for c in mine] == [H.masked(c) for c in theirs]
    a, b = H.decode_yaik(b"".join(mine)), H.decode_yaik(b"".join(theirs))
    assert np.array_equal(a, b)
    assert a.shape == (h, w, 3)


@needs_decoder
@pytest.mark.parametrize("w,h,seed", [(256, 256, 12), (512, 512, 0)])
def test_decodable_stream_decodes_close_to_the_source(w, h, seed):
    """YK_PALETTE_DECODABLE: every pixel comes back within the codec's tolerance of the source — gradient tiles within the
    reject factor plus the 6-bit corner quantisation, range tiles within one quantisation step."""
    from yaik_b200 import build as ykbuild
    lib = capi.load_library(ykbuild.build())
    planes = make_image(w, h, 3, SEED_BASE + seed)
    passes, r2 = _oracle_streams(planes)
    img = H.decode_yaik(b"".join(H.build_yaik(lib, H.ref_zstd_callback(), w, h, passes, r2, palette_mode=1)))
    diff = np.abs(img.astype(np.int32) - planes.transpose(1, 2, 0))
    assert diff.max() <= 12, int(diff.max())
    assert diff.mean() < 3.0


@pytest.mark.gpu
@needs_decoder
def test_cuda_streams_round_trip_through_the_reference_decoder_2048():
    """configs[1]-sized RGB texture: analysis on the GPU, host tails of the product library, the reference's decoder."""
    lib = capi.load_library()
    w = h = 2048
    planes = make_image(w, h, 3, SEED_BASE + 1)
    c = capi.Context(w, h, planes=3, slots=1, lib=lib)
    try:
        c.set_image(planes)
        c.analyze(capi.STAGE_GRADIENT | capi.STAGE_RANGE1D)
        passes = [c.gradient_pass(sx, sy) for sx, sy in capi.PASS_ORDER]
        r2 = [c.range1d(p) for p in range(3)]
        st = c.download_state()
    finally:
        c.close()
    img = H.decode_yaik(b"".join(H.build_yaik(lib, H.ref_zstd_callback(), w, h, passes, r2, palette_mode=1)))
    src = planes.transpose(1, 2, 0)
    diff = np.abs(img.astype(np.int32) - src)
    assert diff.max() <= 12, int(diff.max())
    # gradient tiles: the decoder's interpolation of the dequantised corners against the encoder's own reconstruction
    claimed = st["smoothMap"] != 0
    recon = np.stack([st["recon"][k] for k in range(3)], -1)
    assert np.abs(img.astype(np.int32) - recon)[claimed].max() <= 8
