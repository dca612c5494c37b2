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


# ---- range stage: Decompress1D (decoder/YAIK_3DTile.cpp:24-240) ---------------------------------------------------
def walk_range1d(w, h, cell_claimed, idx, typ, range_compression=15):
    """Replays the decoder's consumption of one plane's DynamicTileCompressor streams.  cell_claimed: bool [h/4][w/4]
    (== smoothMap != 0 sampled per 4x4 cell).  Tiles in row-major order; a tile with an unclaimed quadrant reads
    {color0, base, delta} and then, band by band, row by row, the index bytes of its unclaimed quadrants (the decoder's
    switch on patternQuad & 3).  Returns the decoded plane (-1 where nothing is decoded), the per-pixel delta of the tile
    and the number of coded tiles."""
    idx = np.asarray(idx, np.uint8); typ = np.asarray(typ, np.uint8)
    out = np.full((h, w), -1, np.int32)
    dmap = np.zeros((h, w), np.int32)
    inv = (1 << 24) // range_compression
    ri = rt = 0
    tiles = 0
    for ty in range(h // 8):
        for tx in range(w // 8):
            q = cell_claimed[2 * ty:2 * ty + 2, 2 * tx:2 * tx + 2]
            if q.all():
                continue
            assert rt + 3 <= typ.size, "type stream too short"
            color0, base, delta = (int(v) for v in typ[rt:rt + 3]); rt += 3
            delta2 = ((delta * inv) >> 8) + 1
            tiles += 1
            dmap[8 * ty:8 * ty + 8, 8 * tx:8 * tx + 8] = delta
            for band in range(2):
                left, right = not q[band, 0], not q[band, 1]
                if not (left or right):
                    continue
                x0, x1 = (0 if left else 4), (8 if right else 4)
                for r in range(4):
                    n = x1 - x0
                    L = idx[ri:ri + n].astype(np.int32); ri += n
                    assert L.size == n, "index stream too short"
                    out[8 * ty + 4 * band + r, 8 * tx + x0:8 * tx + x1] = np.where(L != 0, base + (((L - 1) * delta2) >> 16), color0)
    assert ri == idx.size, f"{idx.size - ri} index bytes are never read by the decoder"
    assert rt == typ.size, f"{typ.size - rt} type bytes are never read by the decoder"
    return out, dmap, tiles


def check_range1d(plane, cell_claimed, idx, typ):
    """Consumption (every index and type byte is read, exactly the unclaimed cells are decoded) plus a closeness bound of
    the decoded samples: within one quantisation step (delta / 15) plus the rounding of Model1 / the decoder's fixed point.
    Tiles with delta <= 1 are left out of the bound (the reference's index -1 quirk, SURVEY.md hazard 9)."""
    h, w = plane.shape
    dec, dmap, tiles = walk_range1d(w, h, cell_claimed, idx, typ)
    coded = dec >= 0
    unclaimed_px = ~np.repeat(np.repeat(np.asarray(cell_claimed, bool), 4, 0), 4, 1)[:h, :w]
    assert np.array_equal(coded, unclaimed_px), "decoded pixels are not exactly the unclaimed cells"
    sel = coded & (dmap >= 2)
    err = np.abs(dec - plane.astype(np.int32))
    bound = (dmap + 14) // 15 + 2
    assert (err[sel] <= bound[sel]).all(), ("range-stage round trip too far off", int(err[sel].max()))
    return tiles
