"""Seeded parity cases shared by the oracle-vs-reference tests, the golden-vector generator and the GPU
parity tests.  Each entry returns (planes int32 [C][H][W], stages) — stages name what the reference
harness can run on that size (SURVEY.md hazards 7 and 11: alpha needs pow2 squares >= 16, R2 needs
multiples of 8, R1 needs >= 32)."""
import numpy as np

from yaik_b200.synth import make_image, SEED_BASE, _rand


def _noise(w, h, c, seed, lo=0, hi=255):
    n = w * h * c
    v = (_rand(seed, 99, np.arange(n, dtype=np.uint64)) >> np.uint64(33)).astype(np.int64) % (hi - lo + 1) + lo
    return v.reshape(c, h, w).astype(np.int32)


def _smooth_noisy(w, h, seed, amp):
    """Global bilinear ramp + small noise: many tiles sit near the +-3 tolerance, so the 6 variants, the
    top-left-only eligibility and overlapping 16x8 / 8x16 (8x4 / 4x8) tiles are all exercised."""
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    out = np.empty((3, h, w), np.int64)
    for c in range(3):
        k = (_rand(seed, c, np.arange(4, dtype=np.uint64)) >> np.uint64(33)).astype(np.int64) % 256
        top = k[0] * (w - xx) + k[1] * xx
        bot = k[2] * (w - xx) + k[3] * xx
        out[c] = (top * (h - yy) + bot * yy) // (w * h)
    out += _noise(w, h, 3, seed + 7, 0, 2 * amp) - amp
    return np.clip(out, 0, 255).astype(np.int32)


def _patchy(w, h, seed, cell=8, maxamp=3):
    """Ramp whose noise amplitude changes per `cell` block (0..maxamp): acceptance at every tile size."""
    base = _smooth_noisy(w, h, seed, 0).astype(np.int64)
    by, bx = np.meshgrid(np.arange(h) // cell, np.arange(w) // cell, indexing="ij")
    amp = (_rand(seed, 5, (by * 4096 + bx).astype(np.uint64).ravel()) >> np.uint64(33)).astype(np.int64).reshape(h, w) % (maxamp + 1)
    amp = np.where((_rand(seed, 6, ((by // 4) * 4096 + bx // 4).astype(np.uint64).ravel()) >> np.uint64(60)).reshape(h, w) < 9, 0, amp)
    nz = _noise(w, h, 3, seed + 3, 0, 2 * maxamp).astype(np.int64) - maxamp
    nz = np.clip(nz, -amp, amp)
    return np.clip(base + nz, 0, 255).astype(np.int32)


def _with_alpha(rgb, alpha):
    c, h, w = rgb.shape
    out = np.empty((4, h, w), np.int32)
    out[:3] = rgb[:3]
    out[3] = alpha
    return out


def _alpha_island(w, h, x0, y0, x1, y1, holes=()):
    a = np.zeros((h, w), np.int32)
    a[y0:y1, x0:x1] = 255
    for (hx0, hy0, hx1, hy1) in holes:
        a[hy0:hy1, hx0:hx1] = 0
    return a


ALL = ("alpha", "grad", "r2", "r1")


def _stride_quirk128():
    """Quarter-size chroma reads the full-size mask with the REDUCED width as row stride (Plane.cpp:516): the second-row
    samples of a reduced pixel at (xx, yy) lie 64 pixels to the right of (2xx, yy).  Here the alpha tile under (2xx, yy) is
    rejected while the tile 64 pixels to the right is kept by the alpha stage but claimed by a 16x16 gradient tile, and
    the block still has coded pixels (its coding rule looks at rows 2yy..): the min/max rule must see mipmapMask == 0
    there (FittingQuadSmooth zeroes it, EC.cpp:4035), not just "alpha kept"."""
    rgb = _noise(128, 128, 3, 38, 40, 200)
    rgb[:, 0:33, 64:97] = 90                       # tile (4, 0) and its corner samples: accepted at 16x16
    a = np.full((128, 128), 255, np.int32)
    a[0:16, 0:16] = 0                              # tile (0, 0) rejected
    a[:, 112:128] = 0                              # keeps the kept-tile box smaller than the image (no reset)
    return _with_alpha(rgb, a)

SMALL_CASES = {
    # synthetic illustration-like content
    "synth128_rgb": lambda: (make_image(128, 128, 3, SEED_BASE + 11), ("grad", "r2", "r1")),
    "synth256_rgba": lambda: (make_image(256, 256, 4, SEED_BASE + 12), ALL),
    "synth256_rgb_3bit": lambda: (make_image(256, 256, 3, SEED_BASE + 13), ("grad", "r1_3bit")),
    # near-tolerance ramps: overlaps between non-nesting tile shapes, all six variants
    "ramp64_a2": lambda: (_smooth_noisy(64, 64, 21, 2), ("grad", "r2", "r1")),
    "ramp128_a1": lambda: (_smooth_noisy(128, 128, 22, 1), ("grad", "r2", "r1")),
    "patchy128": lambda: (_patchy(128, 128, 23), ("grad", "r2", "r1")),
    "patchy_192x136": lambda: (_patchy(192, 136, 24, 4, 4), ("grad", "r2", "r1")),   # partial swizzle blocks
    "patchy_72x40": lambda: (_patchy(72, 40, 25, 4, 3), ("grad", "r2", "r1")),
    # extremes
    "flat64": lambda: (np.full((3, 64, 64), 77, np.int32), ("grad", "r2", "r1")),
    "flat_255": lambda: (np.full((3, 64, 64), 255, np.int32), ("grad", "r2", "r1")),
    "noise64": lambda: (_noise(64, 64, 3, 31), ("grad", "r2", "r1")),
    "noise_lowamp96": lambda: (_noise(96, 96, 3, 32, 100, 108), ("grad", "r2", "r1")),
    "noise_delta1": lambda: (_noise(64, 64, 3, 33, 0, 3), ("grad", "r2", "r1")),      # R2 delta==1 -> index -1 quirk
    "noise_hi": lambda: (_noise(64, 64, 3, 34, 250, 255), ("grad", "r2", "r1")),      # R2 clamp 254, R1 base clamp 224
    # R1 alone on signed planes (int32 upload): the +128 shift of blocks with a negative minimum (EC.cpp:764-768).  The
    # reference indexes fullTables[min][max] with the shifted values, so a block needs min >= -128 and max + 128 <= 255
    # when negative, max <= 255 otherwise; anything wider reads outside that array in the reference (undefined there).
    "r1_signed96": lambda: (_noise(96, 96, 3, 36, -120, 100), ("r1",)),
    "r1_signed_edge64": lambda: (_noise(64, 64, 3, 37, -128, 127), ("r1",)),
    # mip tail (SURVEY.md hazard 11)
    "mip32_rgba": lambda: (make_image(32, 32, 4, SEED_BASE + 14), ALL),
    "mip16_rgba": lambda: (make_image(16, 16, 4, SEED_BASE + 15, holes=0), ("alpha", "grad", "r2")),
    "mip8_rgb": lambda: (_smooth_noisy(8, 8, 26, 2), ("grad", "r2")),
    "mip4_rgb": lambda: (_smooth_noisy(4, 4, 27, 1), ("grad",)),
    # alpha: bbox smaller than the image (MIPM chunk written), holes inside, R1 constraint quirks
    "alpha_island128": lambda: (_with_alpha(make_image(128, 128, 3, SEED_BASE + 16),
                                            _alpha_island(128, 128, 16, 32, 100, 90, [(48, 48, 80, 64)])), ALL),
    "alpha_island256": lambda: (_with_alpha(_patchy(256, 256, 28),
                                            _alpha_island(256, 256, 64, 16, 250, 200, [(96, 32, 160, 96), (170, 100, 171, 101)])), ALL),
    "alpha_full_reset64": lambda: (_with_alpha(_noise(64, 64, 3, 35, 90, 99),
                                               _alpha_island(64, 64, 0, 0, 64, 64, [(16, 16, 48, 32)])), ALL),
    "alpha_stride_quirk128": lambda: (_stride_quirk128(), ALL),
    "alpha_corner_only": lambda: (_with_alpha(_patchy(64, 64, 29, 4, 3), _alpha_island(64, 64, 40, 40, 64, 64)), ALL),
}


# Chroma front-end: image, stages run before, (halfCoW, halfCoH, halfCgW, halfCgH), (EDownSample Co, Cg)
CHROMA_CASES = [
    ("synth128_rgb", ("grad",), (1, 0, 1, 0), (2, 2)),            # the CLI's configuration (ImageEncoder.cpp:175-181)
    ("synth256_rgba", ("alpha", "grad"), (1, 1, 1, 1), (2, 2)),    # quarter-size chroma, alpha holes
    ("alpha_island128", ("alpha", "grad"), (1, 0, 0, 1), (2, 2)),  # one axis each, bound box smaller than the image
    ("patchy128", ("grad",), (0, 0, 1, 1), (2, 0)),                # full-size Co, nearest Cg
    ("alpha_island256", ("alpha", "grad"), (1, 1, 1, 1), (3, 4)),  # max / min box
    ("noise64", (), (1, 0, 1, 0), (2, 2)),                         # no gradient stage before (all pixels coded)
    ("alpha_corner_only", ("alpha", "grad"), (1, 1, 1, 0), (1, 2)),
    ("alpha_stride_quirk128", ("alpha", "grad"), (1, 1, 1, 1), (2, 2)),   # reduced-stride mask samples on claimed cells
]


def random_chroma_case(i):
    """Random 16-aligned shape, content, reduction and down-sampling mode; RGBA cases get alpha holes (power-of-two square)."""
    r = _rand(888, i, np.arange(12, dtype=np.uint64))
    rgba = int(r[0] % np.uint64(3)) == 0
    if rgba:
        w = h = 64 << int(r[1] % np.uint64(3))         # 64, 128, 256: the alpha stage's parity domain; reduced planes stay >= 32 (R1's domain)
        planes = make_image(w, h, 4, SEED_BASE + 700 + i)
        pre = ("alpha", "grad")
    else:
        w = 16 * (4 + int(r[1] % np.uint64(7)))        # 64 .. 160 (a reduced plane below 32 crashes the reference)
        h = 16 * (4 + int(r[2] % np.uint64(5)))        # 64 .. 128
        kind = int(r[3] % np.uint64(3))
        planes = (_patchy(w, h, 4000 + i, 4, 2) if kind == 0 else make_image(w, h, 3, SEED_BASE + 800 + i) if kind == 1
                  else _noise(w, h, 3, 5000 + i, 60, 200))
        pre = ("grad",) if int(r[4] % np.uint64(4)) else ()
    cfg = tuple(int(r[5 + k] % np.uint64(2)) for k in range(4))
    modes = []
    for k in range(2):
        both = cfg[2 * k] and cfg[2 * k + 1]
        modes.append(int(r[9 + k] % np.uint64(5)) if both else (0, 2)[int(r[9 + k] % np.uint64(2))])   # one axis: nearest / average only
    return planes, pre, cfg, tuple(modes)
