"""Seeded synthetic illustration-like images shared by the parity tests and bench.py.

64x64 cells drawn from {flat fill, bilinear gradient, hard "anime" edge between two flats,
low-amplitude noise}; RGBA variants add alpha holes (half of them snapped to the 16-pixel tile grid
the alpha-rejection stage works on, the rest unsnapped, some with a soft 8-pixel ramp).

The generator is counter based (splitmix64 of (seed, stream, index)), so it does not depend on
numpy's RNG implementation and any sub-rectangle can be regenerated independently.
Planes are returned as int32 [C][H][W], the reference's ``Plane`` representation
(encoder/framework.h:74-127: ``int* pixels``, row-major ``x + y*w``).
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0x59414B00          # 'YAK\0' + config index
CELL = 64

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(z: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    z = z.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        z += np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _rand(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    """uint64 hash of (seed, stream, idx)."""
    base = _mix(np.asarray([(seed * 0x100000001B3 + stream * 0x9E3779B1) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))
    with np.errstate(over="ignore"):
        return _mix(np.asarray(idx, dtype=np.uint64) * np.uint64(0xD6E8FEB86659FD93) + base[0])


def make_image(w: int, h: int, channels: int = 3, seed: int = SEED_BASE, *,
               mix=(0.30, 0.40, 0.15, 0.15), holes: int | None = None) -> np.ndarray:
    """Return int32 planes [channels][h][w], values 0..255.

    mix = probabilities of (flat, gradient, edge, noise) cells.
    """
    assert channels in (3, 4)
    cw = (w + CELL - 1) // CELL
    ch = (h + CELL - 1) // CELL
    ncell = cw * ch
    cidx = np.arange(ncell, dtype=np.uint64)
    # cell kind
    u = (_rand(seed, 1, cidx) >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    edges = np.cumsum(mix) / np.sum(mix)
    kind = np.searchsorted(edges, u, side="right").clip(0, 3).reshape(ch, cw)

    # lattice of colours at cell corners, shared by neighbouring cells so adjacent gradients are continuous
    lidx = np.arange((cw + 1) * (ch + 1), dtype=np.uint64)
    lat = np.stack([(_rand(seed, 10 + c, lidx) >> np.uint64(33)).astype(np.int64) % 256 for c in range(3)])
    lat = lat.reshape(3, ch + 1, cw + 1)
    # second colour per cell (edge cells), noise amplitude
    col2 = np.stack([(_rand(seed, 20 + c, cidx) >> np.uint64(33)).astype(np.int64) % 256 for c in range(3)]).reshape(3, ch, cw)
    eparam = _rand(seed, 30, cidx).reshape(ch, cw)

    H = ch * CELL
    W = cw * CELL
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    cy, cx = yy // CELL, xx // CELL
    dy, dx = yy % CELL, xx % CELL
    k = kind[cy, cx]
    out = np.empty((3, H, W), dtype=np.int64)
    ep = eparam[cy, cx]
    # edge geometry: diagonal line dx*a + dy*b > c, optionally stepped to 4 pixels
    ea = ((ep >> np.uint64(8)) % np.uint64(5)).astype(np.int64) - 2       # -2..2
    eb = ((ep >> np.uint64(16)) % np.uint64(3)).astype(np.int64) + 1      # 1..3
    ec = ((ep >> np.uint64(24)) % np.uint64(96)).astype(np.int64)
    stepped = ((ep >> np.uint64(40)) & np.uint64(1)).astype(bool)
    sdx = np.where(stepped, (dx // 4) * 4, dx)
    sdy = np.where(stepped, (dy // 4) * 4, dy)
    side = (sdx * ea + sdy * eb) > ec
    pix = (yy.astype(np.uint64) * np.uint64(W) + xx.astype(np.uint64))
    namp = ((ep >> np.uint64(48)) % np.uint64(3)).astype(np.int64) * 4 + 6   # 6, 10, 14
    for c in range(3):
        tl = lat[c][cy, cx]; tr = lat[c][cy, cx + 1]
        bl = lat[c][cy + 1, cx]; br = lat[c][cy + 1, cx + 1]
        top = tl * (CELL - dx) + tr * dx
        bot = bl * (CELL - dx) + br * dx
        grad = (top * (CELL - dy) + bot * dy + CELL * CELL // 2) // (CELL * CELL)
        flat = tl
        edge = np.where(side, col2[c][cy, cx], tl)
        nz = (_rand(seed, 40 + c, pix.ravel()) >> np.uint64(35)).astype(np.int64).reshape(H, W)
        noise = np.clip(tl + nz % (2 * namp + 1) - namp, 0, 255)
        out[c] = np.select([k == 0, k == 1, k == 2], [flat, grad, edge], noise)
    out = out[:, :h, :w]

    if channels == 3:
        return np.ascontiguousarray(out.astype(np.int32))

    # alpha: opaque with rectangular holes
    alpha = np.full((h, w), 255, dtype=np.int64)
    nh = holes if holes is not None else max(2, (w * h) // (256 * 256))
    hidx = np.arange(nh, dtype=np.uint64)
    r0 = _rand(seed, 50, hidx); r1 = _rand(seed, 51, hidx); r2 = _rand(seed, 52, hidx)
    ya, xa = np.arange(h)[:, None], np.arange(w)[None, :]
    for i in range(nh):
        hw = 16 + int(r0[i] % np.uint64(max(17, w // 4)))
        hh = 16 + int((r0[i] >> np.uint64(20)) % np.uint64(max(17, h // 4)))
        x0 = int(r1[i] % np.uint64(max(1, w - hw)))
        y0 = int((r1[i] >> np.uint64(24)) % np.uint64(max(1, h - hh)))
        if i % 2 == 0:      # snapped to the 16x16 alpha tile grid
            x0 &= ~15; y0 &= ~15; hw = (hw + 15) & ~15; hh = (hh + 15) & ~15
        x1, y1 = min(w, x0 + hw), min(h, y0 + hh)
        if int(r2[i] & np.uint64(3)) == 0:     # soft 8-pixel ramp around the hole
            d = np.maximum(np.maximum(x0 - xa, xa - (x1 - 1)), np.maximum(y0 - ya, ya - (y1 - 1)))
            ramp = np.clip(d * 32, 0, 255)
            alpha = np.minimum(alpha, np.where(d <= 0, 0, ramp))
        else:
            alpha[y0:y1, x0:x1] = 0
    # never let the holes touch all four borders (SURVEY.md hazard 8): keep the top-left tile opaque
    alpha[0:16, 0:16] = 255
    res = np.empty((4, h, w), dtype=np.int32)
    res[:3] = out
    res[3] = alpha
    return res


def make_strip_image(w: int, h: int, seed: int = SEED_BASE + 3) -> np.ndarray:
    """RGB image for the tile-row-strip configuration (BASELINE.json configs[3])."""
    return make_image(w, h, 3, seed)


def mip_chain(top: int = 4096, seed: int = SEED_BASE + 4, channels: int = 4):
    """Box-filtered RGBA mip chain top..4 (BASELINE.json configs[4]): list of int32 [C][s][s]."""
    img = make_image(top, top, channels, seed).astype(np.int64)
    chain = [img.astype(np.int32)]
    s = top
    while s > 4:
        img = (img[:, 0::2, 0::2] + img[:, 1::2, 0::2] + img[:, 0::2, 1::2] + img[:, 1::2, 1::2] + 2) // 4
        s //= 2
        lvl = img.astype(np.int32)
        if channels == 4:       # keep alpha binary-ish so whole tiles stay rejected
            lvl[3] = np.where(lvl[3] < 128, 0, lvl[3])
        chain.append(np.ascontiguousarray(lvl))
    return chain


def to_ykin(planes: np.ndarray) -> bytes:
    """Serialise for oracle/_ref/yaik_ref: 'YKIN', int32 w,h,nplanes, then u8 samples plane-major
    ('YKI4' + int32 samples when the planes leave the byte range)."""
    c, h, w = planes.shape
    if int(planes.min()) < 0 or int(planes.max()) > 255:
        return b"YKI4" + np.asarray([w, h, c], dtype="<i4").tobytes() + planes.astype("<i4").tobytes()
    return b"YKIN" + np.asarray([w, h, c], dtype="<i4").tobytes() + planes.astype(np.uint8).tobytes()
