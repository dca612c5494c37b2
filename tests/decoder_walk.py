"""TEST INFRASTRUCTURE: the corner-consumption rule of the reference *decoder* (KLab/YAIK decoder/YAIK_Gradient.cpp:
DecompressGradient16x16 lines 28-130, ..., DecompressGradient4x4 lines 1208-1330 — all seven share it), restated in
Python: accepted tiles are visited in bitmap order (swizzle block by swizzle block, row-major inside a block); for each
the four corners TL, TR, BL, BR are looked up in a colour map on the 4-pixel lattice, and a corner colour is read from the
pass's rgbStream exactly when its lattice point has none yet (mapRGBMask persists over the seven passes).

`walk` replays that on a set of per-pass (bitmap, rgbStream) results and returns the lattice colours it ends up with.
Properties checked by the tests (size independent, no oracle needed): every stream is consumed to the last byte, every
coloured point carries CompressF(Round6(source pixel), 250) (EC.cpp:3183-3194, 4115-4132) of the clamped source pixel, and
the coloured points are exactly those the encoder marked in mappedRGB."""
import numpy as np

PASS_ORDER = [(4, 4), (4, 3), (3, 4), (3, 3), (3, 2), (2, 3), (2, 2)]
_SWZ = {(4, 4): (64, 64), (4, 3): (64, 64), (3, 4): (64, 64), (3, 3): (64, 64), (3, 2): (64, 32), (2, 3): (32, 64), (2, 2): (32, 32)}


def walk(w, h, passes):
    """passes: list of dicts with 'bitmap' and 'rgb' (uint8 arrays) in Convert()'s pass order.  Returns (has, col)."""
    lat_w, lat_h = w // 4 + 1, h // 4 + 1
    has = np.zeros((lat_h, lat_w), bool)
    col = np.zeros((lat_h, lat_w, 3), np.uint8)
    for (shx, shy), res in zip(PASS_ORDER, passes):
        tw, th = 1 << shx, 1 << shy
        bw, bh = _SWZ[(shx, shy)]
        bits = (bw // tw) * (bh // th)
        nbx = (w + bw - 1) // bw
        pos = np.nonzero(np.unpackbits(np.asarray(res["bitmap"], np.uint8), bitorder="little"))[0]
        blk, within = pos // bits, pos % bits
        xs = (blk % nbx) * bw + (within % (bw // tw)) * tw
        ys = (blk // nbx) * bh + (within // (bw // tw)) * th
        rgb = np.asarray(res["rgb"], np.uint8)
        rd = 0
        for x, y in zip(xs.tolist(), ys.tolist()):
            assert x + tw <= w and y + th <= h, "accepted tile outside the image"
            for lx, ly in ((x, y), (x + tw, y), (x, y + th), (x + tw, y + th)):       # LT, RT, LB, RB
                i, j = lx >> 2, ly >> 2
                if not has[j, i]:
                    has[j, i] = True
                    assert rd + 3 <= rgb.size, "rgbStream too short for the bitmap"
                    col[j, i] = rgb[rd:rd + 3]
                    rd += 3
        assert rd == rgb.size, f"pass {(shx, shy)}: {rgb.size - rd} bytes of rgbStream are never read by the decoder"
    return has, col


def expected_colours(planes):
    """CompressF(Round6(clamped source pixel), 250) at every lattice point."""
    c, h, w = planes.shape
    ys = np.minimum(np.arange(h // 4 + 1) * 4, h - 1)
    xs = np.minimum(np.arange(w // 4 + 1) * 4, w - 1)
    v = planes[:3][:, ys][:, :, xs].astype(np.int64)
    r6 = ((v >> 2) << 2) | (v >> 6)
    return np.moveaxis((r6 * 250 + 127) // 255, 0, -1).astype(np.uint8)


def check(planes, passes, mapped_rgb=None):
    c, h, w = planes.shape
    has, col = walk(w, h, passes)
    want = expected_colours(planes)
    assert np.array_equal(col[has], want[has]), "a decoded corner colour is not the source pixel's"
    if mapped_rgb is not None:
        m = np.asarray(mapped_rgb).reshape(h + 1, w + 1)[::4, ::4] != 0
        assert np.array_equal(m, has), "decoder's coloured lattice points differ from the encoder's mappedRGB"
    return int(has.sum())
