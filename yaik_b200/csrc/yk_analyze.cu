// yaik_b200 — yk_k_analyze: the persistent, TMA-staged analysis kernel (sm_100a).
//
// What the reference computes sequentially, tile after tile in stream order (KLab/YAIK, encoder/EncoderContext.cpp =
// "EC.cpp"), is restated in order-free form so that a 64x64 region (the largest swizzle block,
// include/YAIK_private.h:212-276) can be analysed independently of every other region:
//
//   MipPrefilter/quadRecursion (EC.cpp:1257-1427, 357-430)   per 16x16 tile "all alpha == 0", by ballot
//   FittingQuadSmooth accept decisions (EC.cpp:3810-3998)    all 7 passes of the region; tiles of every shape nest in an
//                                                            aligned 16x16 macro tile and eligibility only looks at the
//                                                            tile's own top-left cell, so one warp runs the whole cascade
//                                                            of a macro tile without any block barrier
//   DynamicTileCompressor (EC.cpp:8398-8522)                 the four 8x8 tiles of the macro tile are coded by the same warp
//                                                            right after its cascade (their cells are final by then), from
//                                                            the pixels already in shared memory, into a fixed per-tile
//                                                            place; yk_k_emit moves them to their stream offsets
//
// Structure: one CTA per SM, persistent.  4 producer warps: one thread issues TMA box loads (cp.async.bulk.tensor, one
// per plane) of the next region's int32 samples into a raw staging buffer, completion on an mbarrier; the producer warps
// then pack the samples to bytes into one of three byte tiles (65x65x3, 14 KB), test the alpha plane, and publish the
// tile.  20 consumer warps take (region, macro tile) items from a block-local queue spanning the published tiles, so
// load latency, conversion and the very uneven cost of macro tiles overlap; the warp that finishes the last macro tile
// of a region writes the region's results and recycles the byte tile.  No block-wide barrier after start-up.
//
// No tensor cores: the work is integer min/max reductions over bytes, bounded by HBM and the integer pipes.
#include "yk_device.h"

#define YKA_NB 3                    // byte tiles (regions on the consumer side)
#define YKA_NR 2                    // raw int32 staging buffers (TMA destinations)
#define YKA_CONS_WARPS 20
#define YKA_PROD_WARPS 4
#define YKA_PROD_THREADS (YKA_PROD_WARPS * 32)
#define YKA_THREADS ((YKA_CONS_WARPS + YKA_PROD_WARPS) * 32)
#define YKA_RAW_PLANE_INTS 4448     // 65 rows x 68 ints = 4420, rounded so every plane starts 128-byte aligned
#define YKA_RAW_STAGE_INTS (3 * YKA_RAW_PLANE_INTS + 64 * 64)
#define YKA_COLOR_TX (3u * YK_RAW_ROWS * YK_RAW_PITCH * 4u)
#define YKA_ALPHA_TX (64u * 64u * 4u)
// a consumer only waits on the "ready" sequence number of a byte tile, so items in flight may span at most YKA_NB regions
static_assert(YKA_CONS_WARPS <= 16 * (YKA_NB - 1), "work items in flight must not wrap the byte-tile ring");
static_assert(YKA_NR >= 2, "the ticket of a raw buffer is rewritten one iteration before it is read again");

struct YkaRegion {                  // per byte tile: state of the region being analysed in it
    uint32_t cell[16];              // claimed 4x4 cells, one 16-bit row per entry (cells outside the image count as claimed)
    uint32_t bits[YK_NPASS][8];     // accept bits of the region in swizzled order (EC.cpp:4026)
    int      stat[YK_NPASS][YK_ST_STRIDE];
    uint32_t touch[17 * 17];        // touch words of the region's lattice points
    uint32_t alpha;                 // bit ty*4+tx: 16x16 tile has a non-zero alpha sample
    int      done;                  // macro tiles finished
    int      item;                  // slot * nRegions + region
    int      readySeq;              // n + 1 once the n-th region of this CTA has been staged here
    int      freeSeq;               // n + 1 once it has been finalised
};

struct YkaShared {
    uint32_t pretestTab[41];
    int      queueHead;             // next (region sequence number * 16 + macro tile) to hand out
    int      endSeq;                // first region sequence number that does not exist
    int      rawItem[YKA_NR];       // work item whose samples are (being) loaded into raw buffer i, -1 = none
    unsigned long long mbar[YKA_NR];
    YkaRegion reg[YKA_NB];
};

#define YKA_SMEM_RAW   (YKA_NR * YKA_RAW_STAGE_INTS * 4)
#define YKA_SMEM_PIX   (YKA_NB * 3 * YK_PIXTILE)
#define YKA_SMEM_HIST  (YKA_CONS_WARPS * 3 * 256)
#define YKA_SMEM_BYTES (YKA_SMEM_RAW + YKA_SMEM_PIX + YKA_SMEM_HIST + (int)sizeof(YkaShared) + 128)

// ------------------------------------------------------------------------------------------------------------------
// mbarrier / TMA / named-barrier primitives (inline PTX), with stand-ins for the CPU logic emulation (tests/emu)
#ifdef YK_EMULATE
static std::barrier<> yka_emu_prod_bar(YKA_PROD_THREADS);
static inline void yka_mbar_init(unsigned long long* b, int) { __atomic_store_n(b, 0ull, __ATOMIC_SEQ_CST); }
static inline void yka_mbar_expect_tx(unsigned long long*, unsigned) {}
static inline void yka_mbar_wait(unsigned long long* b, unsigned parity) {     // emulated phase counter: completed loads
    while (((__atomic_load_n(b, __ATOMIC_SEQ_CST) >> 8) & 1ull) == parity) std::this_thread::yield();
}
static inline void yka_tma_box(void* dst, const YkTmap* tm, int x, int y, int bw, int bh, unsigned long long* b, int last) {
    const int32_t* base = (const int32_t*)tm->opaque[0]; const int w = (int)tm->opaque[1], h = (int)tm->opaque[2];
    int32_t* d = (int32_t*)dst;
    for (int r = 0; r < bh; r++) for (int c = 0; c < bw; c++)
        d[r * bw + c] = (y + r < h && x + c < w) ? base[(size_t)(y + r) * w + x + c] : 0;
    if (last) __atomic_fetch_add(b, 256ull, __ATOMIC_SEQ_CST);                  // all boxes of the stage have landed: flip the phase
}
static inline void yka_prod_sync() { yka_emu_prod_bar.arrive_and_wait(); }
static inline void yka_fence_async() {}
static inline void yka_tmap_acquire(const YkTmap*) {}
#else
static __device__ __forceinline__ uint32_t yka_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void yka_mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(yka_s32(b)), "r"(count) : "memory");
}
static __device__ __forceinline__ void yka_mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(yka_s32(b)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void yka_mbar_wait(unsigned long long* b, unsigned parity) {
    const uint32_t a = yka_s32(b);
    unsigned ok = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
static __device__ __forceinline__ void yka_tma_box(void* dst, const YkTmap* tm, int x, int y, int, int, unsigned long long* b, int) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(yka_s32(dst)), "l"((unsigned long long)tm), "r"(x), "r"(y), "r"(yka_s32(b)) : "memory");
}
static __device__ __forceinline__ void yka_prod_sync() { asm volatile("bar.sync 1, %0;" :: "n"(YKA_PROD_THREADS) : "memory"); }
static __device__ __forceinline__ void yka_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// the descriptors live in global memory and are rewritten by host copies between launches
static __device__ __forceinline__ void yka_tmap_acquire(const YkTmap* tm) {
    asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" :: "l"((unsigned long long)tm) : "memory");
}
#endif

// ------------------------------------------------------------------------------------------------------------------
static __device__ __forceinline__ unsigned yka_pack4(int4 v) {      // low bytes of four samples -> one word (3 PRMT)
    return __byte_perm(__byte_perm((unsigned)v.x, (unsigned)v.y, 0x0040), __byte_perm((unsigned)v.z, (unsigned)v.w, 0x0040), 0x5410);
}

// table entry of tile ti (0..40) of a macro tile: offX | offY << 4 | shx << 8 | shy << 11 | cell << 14
static __device__ __forceinline__ uint32_t yka_pretest_entry(int ti) {
    const int pid = (ti >= 1) + (ti >= 3) + (ti >= 5) + (ti >= 9) + (ti >= 17) + (ti >= 25);
    const int t = ti - (int)((0x19110905030100ull >> (8 * pid)) & 255ull);
    const int shx = (0x2233344 >> (4 * pid)) & 15, shy = (0x2323434 >> (4 * pid)) & 15;
    const int tx = t & ((16 >> shx) - 1), ty = t >> (4 - shx);
    const int offX = tx << shx, offY = ty << shy;
    return (uint32_t)(offX | (offY << 4) | (shx << 8) | (shy << 11) | (((offY >> 2) * 4 + (offX >> 2)) << 14));
}

struct YkaView {                    // what a consumer warp knows about the region it works on
    const uint8_t* pix;             // byte tile, channel c at pix + c * YK_PIXTILE
    YkaRegion* R;
    int X0, Y0, w, h, yOrg, rej;
};

// Producer warps: raw int32 boxes -> byte tile (clamped the way Plane::GetPixelValue clamps, framework.h:116-121),
// claimed cells of the region, alpha-zero test of its 16x16 tiles.
static __device__ void yka_convert(const YkSlotDev& S, const int32_t* __restrict__ raw, uint8_t* __restrict__ pix, YkaRegion& R,
                                   int bx, int X0, int Y0, int ptid, bool doAlpha, unsigned& bad) {
    const int w = S.w, h = S.h;
    if (ptid < 16) {
        const int cy = (Y0 >> 2) + ptid;
        uint32_t v = 0xFFFFu;
        if (cy * 4 < h) {
            v = S.cellMask[(size_t)cy * S.nbx + bx];
            const int cellsIn = (w - X0) >> 2;
            if (cellsIn < 16) v |= (0xFFFFu << cellsIn) & 0xFFFFu;
        }
        R.cell[ptid] = v;
    }
    if (X0 + YK_RAW_PITCH <= w && Y0 + YK_RAW_ROWS <= h) {
        // interior region: every sample of the 65x65 tile is inside the box
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int4* __restrict__ src = reinterpret_cast<const int4*>(raw + c * YKA_RAW_PLANE_INTS);
            uint8_t* __restrict__ d = pix + c * YK_PIXTILE;
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int idx = it * YKA_PROD_THREADS + ptid, ly = idx >> 4, q = idx & 15;
                const int4 v = src[ly * (YK_RAW_PITCH / 4) + q];
                bad |= (unsigned)(v.x | v.y | v.z | v.w);
                *reinterpret_cast<unsigned*>(d + ly * YK_RS + 4 * q) = yka_pack4(v);
            }
            if (ptid < 16) {
                const int4 v = src[64 * (YK_RAW_PITCH / 4) + ptid];
                bad |= (unsigned)(v.x | v.y | v.z | v.w);
                *reinterpret_cast<unsigned*>(d + 64 * YK_RS + 4 * ptid) = yka_pack4(v);
            } else if (ptid < 16 + 65) {
                const int ly = ptid - 16;
                const int s = raw[c * YKA_RAW_PLANE_INTS + ly * YK_RAW_PITCH + 64];
                bad |= (unsigned)s;
                d[ly * YK_RS + 64] = (uint8_t)s;
            }
        }
    } else {
        // region at the right / bottom edge (or a partial one): clamp inside the image; in strip mode the row under the
        // strip is the real image row (S.rowBelow), not a clamp
        const int wmax = min(64, w - 1 - X0), hmax = min(64, h - 1 - Y0);
        for (int idx = ptid; idx < 65 * 65; idx += YKA_PROD_THREADS) {
            const int ly = idx / 65, lx = idx - ly * 65;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                int s;
                if (ly > hmax && S.rowBelow[c]) s = __ldg(S.rowBelow[c] + min(X0 + lx, w - 1));
                else s = raw[c * YKA_RAW_PLANE_INTS + min(ly, hmax) * YK_RAW_PITCH + min(lx, wmax)];
                bad |= (unsigned)s;
                pix[c * YK_PIXTILE + ly * YK_RS + lx] = (uint8_t)s;
            }
        }
    }
    if (doAlpha) {
        // all(alpha == 0) per 16x16 tile (EC.cpp:357-430 restated per tile); samples outside the image arrive as zeros
        const int4* __restrict__ a = reinterpret_cast<const int4*>(raw + 3 * YKA_RAW_PLANE_INTS);
        unsigned am = 0;
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int idx = it * YKA_PROD_THREADS + ptid, ly = idx >> 4;
            const int4 v = a[idx];
            unsigned b = __ballot_sync(YK_FULL, (v.x | v.y | v.z | v.w) != 0);
            b |= b >> 16;                       // a warp covers two rows of the same tile row
            const int ty = ly >> 4;
#pragma unroll
            for (int tx = 0; tx < 4; tx++) if (b & (0xFu << (4 * tx))) am |= 1u << (ty * 4 + tx);
        }
        if ((ptid & 31) == 0 && am) atomicOr(&R.alpha, am);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Cheap rejection of all 41 tiles (1 + 2 + 2 + 4 + 8 + 8 + 16 over the seven shapes) of a 16x16 macro tile: lane = tile,
// one channel, the quad at the tile centre, raw corners.  A pixel whose raw-family U is outside [loWide, hiWide) cannot
// be accepted by any of the six variants (a family's corners differ from the raw ones by -3..+4), so a cleared bit is a
// proven rejection; a set bit only means "run the real test".  Bit (start(pid) + t) belongs to tile t of pass id pid.
static __device__ __forceinline__ unsigned long long yka_pretest(const YkaView& V, const uint32_t* sTab, int mlx, int mly, unsigned claimed) {
    const int lane = threadIdx.x & 31, R = V.rej;
    unsigned long long P = 0;
#pragma unroll
    for (int round = 0; round < 2; round++) {
        const int ti = lane + 32 * round;
        bool possible = false;
        if (ti < 41) {
            const uint32_t e = sTab[ti];
            const int shx = (e >> 8) & 7, shy = (e >> 11) & 7, sh = shx + shy;
            const int lx0 = mlx + (e & 15), ly0 = mly + ((e >> 4) & 15), TW = 1 << shx, TH = 1 << shy, N = 1 << sh;
            if (!((claimed >> (e >> 14)) & 1u) && V.X0 + lx0 + TW <= V.w && V.Y0 + ly0 + TH <= V.h) {
                const uint8_t* p = V.pix + ly0 * YK_RS + lx0;
                const int tl = p[0], tr = p[TW], bl = p[TH * YK_RS], br = p[TH * YK_RS + TW];
                const int dx0 = (TW >> 1) & ~3, dy = TH >> 1;
                const unsigned word = *reinterpret_cast<const unsigned*>(p + dy * YK_RS + dx0);
                const int B = (tr - tl) << shy, C = (bl - tl) << shx, D = tl - tr - bl + br;
                const int step = B + D * dy;
                const int s0 = ((tl + R) << sh) + B * dx0 + dy * (C + D * dx0);
                const int u0 = s0 - (int)((word & 255u) << sh), u1 = s0 + step - (int)(((word >> 8) & 255u) << sh);
                const int u2 = s0 + 2 * step - (int)(((word >> 16) & 255u) << sh), u3 = s0 + 3 * step - (int)((word >> 24) << sh);
                const int umin = __vimin3_s32(min(u0, u1), u2, u3), umax = __vimax3_s32(max(u0, u1), u2, u3);
                possible = !(umin < -(4 * N + N / 2 - 1) || umax >= (2 * R + 4) * N);
            }
        }
        P |= (unsigned long long)__ballot_sync(YK_FULL, possible) << (32 * round);
    }
    return P;
}

// four consecutive pixels of one channel: U = S + R*N - cur*N folded into a running min/max.
// |cur - S/N| <= R  <=>  0 <= U < (2R+1)N;   |cur - (S+N/2-1)/N| <= R  <=>  -(N/2-1) <= U < (2R+1)N-(N/2-1)
// (S = bilinear numerator with integer weights; identical to ((bT*tF+bB*bF)[+2^19-1])>>20 of EC.cpp:3937-3965).
static __device__ __forceinline__ void yka_quad(unsigned word, int s, int step, int negN, int& umin, int& umax) {
    const int u0 = (int)(word & 255u) * negN + s;
    const int u1 = (int)((word >> 8) & 255u) * negN + (s + step);
    const int u2 = (int)((word >> 16) & 255u) * negN + (s + 2 * step);
    const int u3 = (int)(word >> 24) * negN + (s + 3 * step);
    umin = __vimin3_s32(umin, u0, u1); umin = __vimin3_s32(umin, u2, u3);
    umax = __vimax3_s32(umax, u0, u1); umax = __vimax3_s32(umax, u2, u3);
}

// The accept test of FittingQuadSmooth (EC.cpp:3810-3998) for the tile this lane belongs to.  The lanes in `gmask` share
// the tile; each holds nq (1 or 2) quads of one pixel row of it in `wd` (quad 0 at dx0, quad 1 at dx0 + 4, row dy).
// Returns, uniformly over the group, whether any of the six variants (3 corner families x rounded / truncated) keeps
// every pixel of every channel within the reject factor.
static __device__ __forceinline__ bool yka_tile_test(const uint8_t* __restrict__ corner, int shx, int shy, int dx0, int dy, bool two,
                                                     const unsigned (&wd)[3][2], unsigned gmask, bool active, int R) {
    const int sh = shx + shy, N = 1 << sh, TW = 1 << shx, THp = YK_RS << shy;
    const int hiT = (2 * R + 1) * N;                    // |cur - S/N| <= R            <=>  0 <= U < hiT
    const int loR = -(N / 2 - 1);                       // |cur - (S+N/2-1)/N| <= R    <=>  loR <= U < hiT + loR
    const int loWide = -(4 * N + N / 2 - 1), hiWide = hiT + 3 * N;
    int cr[3][4];                                       // TL TR BL BR, clamped at the image edge by the staging (EC.cpp:3845-3868)
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const uint8_t* p = corner + c * YK_PIXTILE;
        cr[c][0] = p[0]; cr[c][1] = p[TW]; cr[c][2] = p[THp]; cr[c][3] = p[THp + TW];
    }
    bool resolved = !active, accepted = false;
#pragma unroll 1
    for (int fam = 0; fam < 3; fam++) {
        int umin = INT_MAX, umax = INT_MIN;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            int tl = cr[c][0], tr = cr[c][1], bl = cr[c][2], br = cr[c][3];
            if (fam == 1) { tl = yk_round6(tl); tr = yk_round6(tr); bl = yk_round6(bl); br = yk_round6(br); }
            else if (fam == 2) { tl = yk_round6p(tl); tr = yk_round6p(tr); bl = yk_round6p(bl); br = yk_round6p(br); }
            const int B = (tr - tl) << shy, C = (bl - tl) << shx, D = tl - tr - bl + br;
            const int step = B + D * dy;
            const int s = ((tl + R) << sh) + B * dx0 + dy * (C + D * dx0);
            yka_quad(wd[c][0], s, step, -N, umin, umax);
            if (two) yka_quad(wd[c][1], s + 4 * step, step, -N, umin, umax);
            if (fam == 0 && c == 0) {
                // one channel with the raw corners proves most non-gradient tiles hopeless for every variant
                const unsigned bH = __ballot_sync(YK_FULL, (umin < loWide) || (umax >= hiWide));
                if (bH & gmask) resolved = true;
                if (!__any_sync(YK_FULL, !resolved)) return false;
            }
        }
        const bool dT = (umin < 0) || (umax >= hiT);
        const bool dR = (umin < loR) || (umax >= hiT + loR);
        const unsigned bT = __ballot_sync(YK_FULL, dT), bR = __ballot_sync(YK_FULL, dR);
        const bool famDead = ((bT & gmask) != 0u) && ((bR & gmask) != 0u);
        if (fam == 0) {
            const unsigned bH = __ballot_sync(YK_FULL, (umin < loWide) || (umax >= hiWide));
            if (bH & gmask) resolved = true;            // no family can accept this tile
        }
        if (!resolved && !famDead) { accepted = true; resolved = true; }         // EC.cpp:3998: any surviving variant accepts
        if (!__any_sync(YK_FULL, !resolved)) break;
    }
    return accepted;
}

// Side effects of an accepted tile (EC.cpp:3998-4132) that are local to the region: bitmap bit, TileDone / bounding box,
// the four lattice points it touches with its role at each.  (lx0, ly0) = tile origin inside the region.
static __device__ __forceinline__ void yka_commit(const YkaView& V, const YkGeomS& g, int pid, int rp, int lx0, int ly0) {
    YkaRegion& R = *V.R;
    const int TW = 1 << g.shx, TH = 1 << g.shy;
    const int sub = (ly0 >> g.lbh) * (64 >> g.lbw) + (lx0 >> g.lbw);
    const int li = sub * g.bits + (((ly0 & ((1 << g.lbh) - 1)) >> g.shy) << (g.lbw - g.shx)) + ((lx0 & ((1 << g.lbw) - 1)) >> g.shx);
    atomicOr(&R.bits[pid][li >> 5], 1u << (li & 31));                          // EC.cpp:4026
    atomicAdd(&R.stat[pid][YK_ST_TILEDONE], 1);                                // EC.cpp:4039-4044 (mins stored as extent - value)
    atomicMax(&R.stat[pid][YK_ST_MINX], V.w - (V.X0 + lx0));
    atomicMax(&R.stat[pid][YK_ST_MINY], INT_MAX / 2 - (V.yOrg + V.Y0 + ly0));
    atomicMax(&R.stat[pid][YK_ST_MAXX], V.X0 + lx0 + TW);
    atomicMax(&R.stat[pid][YK_ST_MAXY], V.yOrg + V.Y0 + ly0 + TH);
    const int i0 = lx0 >> 2, j0 = ly0 >> 2;                                    // mappedRGB claim, EC.cpp:4001-4021
    atomicOr(&R.touch[j0 * 17 + i0], 1u << (4 * rp + 0));
    atomicOr(&R.touch[j0 * 17 + i0 + (TW >> 2)], 1u << (4 * rp + 1));
    atomicOr(&R.touch[(j0 + (TH >> 2)) * 17 + i0], 1u << (4 * rp + 2));
    atomicOr(&R.touch[(j0 + (TH >> 2)) * 17 + i0 + (TW >> 2)], 1u << (4 * rp + 3));
}

// One FittingQuadSmooth pass over one 16x16 macro tile, by one warp.  Lane (row = lane >> 1, half = lane & 1) owns the
// eight pixels (8*half .. 8*half+7, row) of the macro tile in every pass — they are loaded once into `pw` — so a tile
// of 8 or 16 pixels width is shared by the lanes of its rows, and a 4-pixel-wide pass is run as two sub-passes (left and
// right quad of every lane).  `claimed` (16 bits, bit = 4*cellY + cellX) is warp-uniform and returned updated.
static __device__ __forceinline__ unsigned yka_macro_pass(const YkaView& V, int pid, int rp, int mlx, int mly, unsigned claimed, unsigned poss,
                                                          const uint2 (&pw)[3]) {
    const YkGeomS g = yk_geom_s(pid);
    const int lane = threadIdx.x & 31, row = lane >> 1, half = lane & 1;
    const int shx = g.shx, shy = g.shy, TH = 1 << shy;
    const int ty = row >> shy, lyT = ty << shy, dy = row - lyT;
    const unsigned rowMask = (shy == 4) ? YK_FULL : (((1u << (2 * TH)) - 1u) << (2 * lyT));
    const unsigned gmask = (shx == 4) ? rowMask : (rowMask & (0x55555555u << half));
    const bool leader = lane == __ffs((int)gmask) - 1;
    const int nSub = (shx == 2) ? 2 : 1;
    unsigned newCells = 0;
    for (int sub = 0; sub < nSub; sub++) {
        const int tx = (shx == 2) ? (2 * half + sub) : ((8 * half) >> shx);
        const int lxT = tx << shx, dx0 = (shx == 2) ? 0 : (8 * half - lxT);
        const int t = ty * (16 >> shx) + tx;
        const int cellX = lxT >> 2, cellY = lyT >> 2;
        const bool active = ((poss >> t) & 1u) && !((claimed >> (cellY * 4 + cellX)) & 1u);      // EC.cpp:3818, 3826, 3871-3875
        if (!__any_sync(YK_FULL, active)) continue;
        unsigned wd[3][2];
#pragma unroll
        for (int c = 0; c < 3; c++) { wd[c][0] = (shx == 2 && sub) ? pw[c].y : pw[c].x; wd[c][1] = pw[c].y; }
        const uint8_t* corner = V.pix + (mly + lyT) * YK_RS + mlx + lxT;
        const bool acc = yka_tile_test(corner, shx, shy, dx0, dy, shx != 2, wd, gmask, active, V.rej);
        unsigned mine = 0;
        if (acc && leader) {
            yka_commit(V, g, pid, rp, mlx + lxT, mly + lyT);
            // EC.cpp:4029-4037: the tile's cells become claimed
            const unsigned cols = ((1u << (1 << (shx - 2))) - 1u) << cellX;
            const unsigned rowsPat = (0x1111u & ((1u << (4 << (shy - 2))) - 1u)) << (4 * cellY);
            mine = cols * rowsPat;
        }
        newCells |= __reduce_or_sync(YK_FULL, mine);
    }
    return claimed | newCells;
}

// DynamicTileCompressor (EC.cpp:8398-8522) for the 8x8 tile at (lx8, ly8) of the region; q = its quadrants to code
// (bit0 TL, 1 TR, 2 BL, 3 BR: top-left map pixel 0, EC.cpp:8420-8430 == 4x4 cell unclaimed).  Lane = two pixels; the
// three planes side by side.  Output goes to the tile's fixed place in r2Raw / r2RawType.
static __device__ __forceinline__ void yka_range_tile(const YkSlotDev& S, const YkaView& V, uint8_t* __restrict__ hist, const uint32_t* __restrict__ magicTab,
                                                      int lx8, int ly8, unsigned q) {
    const int lane = threadIdx.x & 31;
    const int r = lane >> 2, c0 = (lane & 3) * 2;           // pixel row / first column of this lane inside the tile
    const int band = r >> 2, right = c0 >> 2;
    const bool valid = (q >> (band * 2 + right)) & 1u;
    const unsigned qb = (q >> (band * 2)) & 3u;             // coded quadrants of this band: bit0 left, bit1 right
    const int lengthX = (qb == 3u) ? 8 : 4, x2 = (qb == 2u) ? 4 : 0;
    const int pos = (band ? 16 * __popc(q & 3u) : 0) + (r & 3) * lengthX + (c0 - x2);
    int vx[3], vy[3];
    unsigned mxA[3], mxB[3];
#pragma unroll
    for (int p = 0; p < 3; p++) {
        const unsigned short two = *reinterpret_cast<const unsigned short*>(V.pix + p * YK_PIXTILE + (ly8 + r) * YK_RS + lx8 + c0);
        vx[p] = two & 255; vy[p] = two >> 8;                 // CompressF(v,255) == v (EC.cpp:8442)
        // FindAndRemoveMostUsedColor (EC.cpp:8335-8356): counts per present value from two match rounds
        mxA[p] = __match_any_sync(YK_FULL, valid ? vx[p] : 256 + lane);
        mxB[p] = __match_any_sync(YK_FULL, valid ? vy[p] : 512 + lane);
    }
#pragma unroll
    for (int p = 0; p < 3; p++)
        if (valid && lane == __ffs((int)mxA[p]) - 1) hist[p * 256 + vx[p]] = (uint8_t)__popc(mxA[p]);
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 3; p++)
        if (valid && lane == __ffs((int)mxB[p]) - 1) hist[p * 256 + vy[p]] = (uint8_t)(hist[p * 256 + vy[p]] + __popc(mxB[p]));
    __syncwarp();
    unsigned key[3];
#pragma unroll
    for (int p = 0; p < 3; p++) {
        // highest index among the maximal counts (`>=`, EC.cpp:8340); only present values can win
        key[p] = valid ? max(((unsigned)hist[p * 256 + vx[p]] << 8) | (unsigned)vx[p], ((unsigned)hist[p * 256 + vy[p]] << 8) | (unsigned)vy[p]) : 0u;
        key[p] = __reduce_max_sync(YK_FULL, key[p]);
    }
    __syncwarp();
    if (valid) {
#pragma unroll
        for (int p = 0; p < 3; p++) { hist[p * 256 + vx[p]] = 0; hist[p * 256 + vy[p]] = 0; }
    }
    const int tilesW = V.w >> 3;
    const size_t tile = (size_t)((V.Y0 + ly8) >> 3) * tilesW + ((V.X0 + lx8) >> 3);
#pragma unroll
    for (int p = 0; p < 3; p++) {
        const int color0 = min(max((int)(key[p] & 255u), 1), 254);
        // Model1 (EC.cpp:8358-8381) over what is left of the histogram
        const bool remx = valid && (vx[p] < color0 - 1 || vx[p] > color0 + 1);
        const bool remy = valid && (vy[p] < color0 - 1 || vy[p] > color0 + 1);
        const int mn = __reduce_min_sync(YK_FULL, min(remx ? vx[p] : 999, remy ? vy[p] : 999));
        const int mx = __reduce_max_sync(YK_FULL, max(remx ? vx[p] : -1, remy ? vy[p] : -1));
        int minCol = 0, delta = 0;
        if (mn != 999) { minCol = mn; delta = mx - mn; }
        if (valid) {
            // GetValueModel1 (EC.cpp:8383-8391): C division of a numerator in -1..3951 by delta in 1..255.
            // floor(n/d) == (n * ceil(2^20/d)) >> 20 for 0 <= n < 4112, d <= 255; n == -1 only happens for delta == 1.
            int bxv = 0, byv = 0;
            if (delta) {
                const unsigned magic = magicTab[delta];
                const int rnd = (delta >> 1) - 1;
                if (remx) { const int n = (vx[p] - minCol) * 15 + rnd; bxv = 1 + (n < 0 ? n : (int)(((unsigned)n * magic) >> 20)); }
                if (remy) { const int n = (vy[p] - minCol) * 15 + rnd; byv = 1 + (n < 0 ? n : (int)(((unsigned)n * magic) >> 20)); }
            } else { bxv = remx ? 1 : 0; byv = remy ? 1 : 0; }
            *reinterpret_cast<uint16_t*>(S.r2Raw[p] + tile * 64 + pos) = (uint16_t)((bxv & 255) | ((byv & 255) << 8));
        }
        if (lane == p) S.r2RawType[p][tile] = (uint32_t)color0 | ((uint32_t)minCol << 8) | ((uint32_t)delta << 16);     // EC.cpp:8503-8505
    }
    __syncwarp();       // the histogram entries are clean again before the next tile fills them
}

// One (region, macro tile) work item: the cascade of Convert()'s passes (EC.cpp:9057-9093), then the range stage.
static __device__ void yka_macro_tile(const YkSlotDev& S, const YkaView& V, const YkRun& run, const uint32_t* sTab, uint8_t* hist,
                                      const uint32_t* magicTab, int m) {
    const int lane = threadIdx.x & 31;
    YkaRegion& R = *V.R;
    const int mx = m & 3, my = m >> 2, mlx = 16 * mx, mly = 16 * my;
    unsigned claimed = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) claimed |= ((R.cell[my * 4 + r] >> (4 * mx)) & 15u) << (4 * r);
    const unsigned claimed0 = claimed;
    if (run.nPasses > 0 && claimed != 0xFFFFu) {
        uint2 pw[3];
#pragma unroll
        for (int c = 0; c < 3; c++)
            pw[c] = *reinterpret_cast<const uint2*>(V.pix + c * YK_PIXTILE + (mly + (lane >> 1)) * YK_RS + mlx + 8 * (lane & 1));
        // the 16x16 pass usually runs first and straight away (most macro tiles of illustration-like content end there);
        // the other shapes are pre-tested together, once, the first time one of them comes up
        const bool in16 = (V.X0 + mlx + 16 <= V.w) && (V.Y0 + mly + 16 <= V.h);
        unsigned long long P = in16 ? 1ull : 0ull;
        bool pretested = false;
        for (int rp = 0; rp < run.nPasses && claimed != 0xFFFFu; rp++) {
            const int pid = run.passId[rp];
            if (pid != 0 && !pretested) { P = yka_pretest(V, sTab, mlx, mly, claimed); pretested = true; }
            const unsigned poss = (unsigned)(P >> yk_geom_s_tab[pid].start) & ((1u << (256 >> (yk_geom_s_tab[pid].shx + yk_geom_s_tab[pid].shy))) - 1u);
            if (poss) claimed = yka_macro_pass(V, pid, rp, mlx, mly, claimed, poss, pw);
        }
        if (lane < 4 && claimed != claimed0) atomicOr(&R.cell[my * 4 + lane], ((claimed >> (4 * lane)) & 15u) << (4 * mx));
    }
    if (run.doR2 && claimed != 0xFFFFu) {
#pragma unroll 1
        for (int t8 = 0; t8 < 4; t8++) {
            const int qx = t8 & 1, qy = t8 >> 1;
            const unsigned c4 = claimed >> (8 * qy + 2 * qx);
            const unsigned q = (~((c4 & 3u) | (((c4 >> 4) & 3u) << 2))) & 15u;
            if (q) yka_range_tile(S, V, hist, magicTab, mlx + 8 * qx, mly + 8 * qy, q);
        }
    }
}

// Results of a finished region, by one warp: accept bitmaps, claimed cells, corner colours and touch words of its
// lattice points, alpha tiles, per-pass counters.  Leaves the region state zeroed for the next region staged here.
static __device__ void yka_finalize(const YkSlotDev& S, const YkaView& V, const YkRun& run, int bx, int by, bool doAlpha) {
    const int lane = threadIdx.x & 31;
    YkaRegion& R = *V.R;
    const int w = V.w, h = V.h, X0 = V.X0, Y0 = V.Y0, nbx = S.nbx;
    // accept bitmaps in the reference's swizzled layout: 16-bit units of each sub-block
    for (int i = lane; i < run.nPasses * 16; i += 32) {
        const int pid = run.passId[i >> 4], u = i & 15;
        const YkGeomS g = yk_geom_s(pid);
        const int nsub = (64 >> g.lbw) * (64 >> g.lbh);
        if (u * 16 < nsub * g.bits) {
            const int sub = (u * 16) / g.bits, within = (u * 16) % g.bits;
            const int sx = X0 + ((sub % (64 >> g.lbw)) << g.lbw), sy = Y0 + ((sub / (64 >> g.lbw)) << g.lbh);
            if (sx < w && sy < h) {
                const int nSwzX = (w + (1 << g.lbw) - 1) >> g.lbw;
                const int gb = (sy >> g.lbh) * nSwzX + (sx >> g.lbw);
                const uint32_t v = (R.bits[pid][(u * 16) >> 5] >> ((u * 16) & 31)) & 0xFFFFu;
                reinterpret_cast<uint16_t*>(S.bitmap[pid])[((size_t)gb * g.bits + within) >> 4] = (uint16_t)v;
            }
        }
    }
    if (lane < 16 && run.nPasses > 0) {
        const int cy = (Y0 >> 2) + lane;
        if (cy * 4 < h) S.cellMask[(size_t)cy * nbx + bx] = (uint16_t)R.cell[lane];
    }
    // corner colours at every 4-pixel lattice point of the region (what an accepted tile would emit, EC.cpp:4115-4132),
    // and the touch words of the lattice points (interior points are exclusive to the region, border points are shared)
    if (run.nPasses > 0) {
        const int iMax = (bx == nbx - 1) ? 17 : 16, jMax = (by == S.nby - 1) ? 17 : 16;
        for (int idx = lane; idx < 17 * 17; idx += 32) {
            const int jj = idx / 17, i = idx - jj * 17;
            const int gx = (X0 >> 2) + i, gy = (Y0 >> 2) + jj;
            if (gx < S.latW && gy < S.latH) {
                if (i < iMax && jj < jMax) {
                    uint8_t* d = S.latRGB + ((size_t)gy * S.latW + gx) * 3;
#pragma unroll
                    for (int c = 0; c < 3; c++) d[c] = (uint8_t)yk_compress250(yk_round6(V.pix[c * YK_PIXTILE + (4 * jj) * YK_RS + 4 * i]));
                }
                const uint32_t tv = R.touch[idx];
                if (tv) atomicOr(&S.touchMap[(size_t)gy * S.latW + gx], tv);
            }
            R.touch[idx] = 0;
        }
    }
    if (doAlpha) {
        const int tx = lane & 3, ty = (lane >> 2) & 3;
        const int px = X0 + 16 * tx, py = Y0 + 16 * ty;
        const bool in = lane < 16 && px < w && py < h;
        const bool kept = in && ((R.alpha >> lane) & 1u);
        if (in) S.alphaKept[(size_t)(py >> 4) * ((w + 15) >> 4) + (px >> 4)] = kept ? 1 : 0;
        // bounding box of kept tiles (EC.cpp:416-422), mins stored as extent - value so the header can be memset to 0
        const int big = INT_MAX / 2;
        const int mnx = __reduce_max_sync(YK_FULL, kept ? w - px : 0);
        const int mny = __reduce_max_sync(YK_FULL, kept ? big - (V.yOrg + py) : 0);
        const int mxx = __reduce_max_sync(YK_FULL, kept ? min(px + 16, w) : 0);
        const int mxy = __reduce_max_sync(YK_FULL, kept ? V.yOrg + min(py + 16, h) : 0);
        const int cnt = __popc(__ballot_sync(YK_FULL, kept));
        if (lane == 0 && cnt) {
            atomicMax(&S.hdr[YK_HD_ALPHA_MINX], mnx); atomicMax(&S.hdr[YK_HD_ALPHA_MINY], mny);
            atomicMax(&S.hdr[YK_HD_ALPHA_MAXX], mxx); atomicMax(&S.hdr[YK_HD_ALPHA_MAXY], mxy);
            atomicAdd(&S.hdr[YK_HD_ALPHA_KEPT], cnt);
        }
    }
    if (lane < YK_NPASS) {
        const int pid = lane;
        if (R.stat[pid][YK_ST_TILEDONE] > 0) {
            int* d = S.hdr + YK_HD_PASS0 + pid * YK_ST_STRIDE;
            atomicAdd(&d[YK_ST_TILEDONE], R.stat[pid][YK_ST_TILEDONE]);
            atomicMax(&d[YK_ST_MINX], R.stat[pid][YK_ST_MINX]); atomicMax(&d[YK_ST_MINY], R.stat[pid][YK_ST_MINY]);
            atomicMax(&d[YK_ST_MAXX], R.stat[pid][YK_ST_MAXX]); atomicMax(&d[YK_ST_MAXY], R.stat[pid][YK_ST_MAXY]);
        }
    }
    __syncwarp();
    for (int i = lane; i < YK_NPASS * 8; i += 32) (&R.bits[0][0])[i] = 0;
    for (int i = lane; i < YK_NPASS * YK_ST_STRIDE; i += 32) (&R.stat[0][0])[i] = 0;
    if (lane == 0) { R.alpha = 0; R.done = 0; }
}

__global__ void __launch_bounds__(YKA_THREADS, 1)
yk_k_analyze(const YkSlotDev* __restrict__ slots, int slot0, int nSlots, int nRegions, YkRun run) {
#ifdef YK_EMULATE
    static unsigned char smemRaw[YKA_SMEM_BYTES + 128];
#else
    extern __shared__ unsigned char smemRaw[];
#endif
    __shared__ uint32_t sMagic[256];                         // ceil(2^20 / d): exact floor(n / d) for n < 4112, d <= 255
    unsigned char* base = (unsigned char*)(((uintptr_t)smemRaw + 127) & ~(uintptr_t)127);
    int32_t* raw = (int32_t*)base;
    uint8_t* pixAll = base + YKA_SMEM_RAW;
    uint8_t* histAll = pixAll + YKA_SMEM_PIX;
    YkaShared& sh = *reinterpret_cast<YkaShared*>(histAll + YKA_SMEM_HIST);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = nSlots * nRegions;
    int* ticket = slots[slot0].hdr + YK_HD_TICKET_ANALYZE;

    // ---- start-up (the only block-wide barrier)
    for (int i = tid; i < (int)(sizeof(YkaShared) / 4); i += YKA_THREADS) reinterpret_cast<uint32_t*>(&sh)[i] = 0;
    for (int i = tid; i < YKA_SMEM_HIST / 4; i += YKA_THREADS) reinterpret_cast<uint32_t*>(histAll)[i] = 0;
    if (tid < 256) sMagic[tid] = tid ? ((1u << 20) + (unsigned)tid - 1u) / (unsigned)tid : 0u;
    __syncthreads();
    if (tid < 41) sh.pretestTab[tid] = yka_pretest_entry(tid);
    if (tid == 64) {
        sh.endSeq = INT_MAX;
        for (int i = 0; i < YKA_NR; i++) { yka_mbar_init(&sh.mbar[i], 1); sh.rawItem[i] = -1; }
#ifndef YK_EMULATE
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    __syncthreads();

    if (warp >= YKA_CONS_WARPS) {
        // ================================================= producers =================================================
        const int ptid = tid - YKA_CONS_WARPS * 32;
        int lastSlot = -1;
        auto issue = [&](int i, int item) {                 // one thread: TMA box loads of `item` into raw buffer i
            const int slot = slot0 + item / nRegions, region = item % nRegions;
            const YkSlotDev& S = slots[slot];
            if (slot != lastSlot) { for (int c = 0; c < S.nPlanes; c++) yka_tmap_acquire(&S.tmap[c]); lastSlot = slot; }
            const int bx = region % S.nbx, by = region / S.nbx;
            const bool alpha = run.doAlpha && S.nPlanes == 4;
            int32_t* dst = raw + i * YKA_RAW_STAGE_INTS;
            yka_fence_async();
            yka_mbar_expect_tx(&sh.mbar[i], YKA_COLOR_TX + (alpha ? YKA_ALPHA_TX : 0u));
            for (int c = 0; c < 3; c++)
                yka_tma_box(dst + c * YKA_RAW_PLANE_INTS, &S.tmap[c], bx * 64, by * 64, YK_RAW_PITCH, YK_RAW_ROWS, &sh.mbar[i], !alpha && c == 2);
            if (alpha) yka_tma_box(dst + 3 * YKA_RAW_PLANE_INTS, &S.tmap[3], bx * 64, by * 64, 64, 64, &sh.mbar[i], 1);
        };
        if (ptid == 0) {
            for (int i = 0; i < YKA_NR; i++) {
                const int t = atomicAdd(ticket, 1);
                sh.rawItem[i] = t < total ? t : -1;
                if (t < total) issue(i, t);
            }
        }
        yka_prod_sync();
        for (int n = 0;; n++) {
            const int i = n % YKA_NR, j = n % YKA_NB;
            const int item = yk_ldvi(&sh.rawItem[i]);
            if (item < 0) { if (ptid == 0) yk_stvi(&sh.endSeq, n); break; }
            YkaRegion& R = sh.reg[j];
            if (n >= YKA_NB) { while (yk_ldvi(&R.freeSeq) != n - YKA_NB + 1) yk_spin(); }
            __threadfence_block();
            yka_mbar_wait(&sh.mbar[i], (unsigned)((n / YKA_NR) & 1));
            const int slot = slot0 + item / nRegions, region = item % nRegions;
            const YkSlotDev& S = slots[slot];
            const int bx = region % S.nbx, by = region / S.nbx;
            unsigned bad = 0;
            yka_convert(S, raw + i * YKA_RAW_STAGE_INTS, pixAll + j * 3 * YK_PIXTILE, R, bx, bx * 64, by * 64, ptid, run.doAlpha && S.nPlanes == 4, bad);
            if (bad & ~255u) atomicOr(&S.hdr[YK_HD_ERR], 1);
            __threadfence_block();
            yka_prod_sync();                                // tile complete, raw buffer i free
            if (ptid == 0) {
                R.item = item;
                __threadfence_block();
                yk_stvi(&R.readySeq, n + 1);
                const int t = atomicAdd(ticket, 1);
                yk_stvi(&sh.rawItem[i], t < total ? t : -1);
                if (t < total) issue(i, t);
            }
        }
        return;
    }

    // ===================================================== consumers =====================================================
    uint8_t* hist = histAll + warp * 3 * 256;
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(&sh.queueHead, 1);
        q = __shfl_sync(YK_FULL, q, 0);
        const int n = q >> 4, m = q & 15, j = n % YKA_NB;
        YkaRegion& R = sh.reg[j];
        bool alive = true;
        while (yk_ldvi(&R.readySeq) != n + 1) {
            if (yk_ldvi(&sh.endSeq) <= n) { alive = false; break; }
            yk_spin();
        }
        alive = __all_sync(YK_FULL, alive);
        if (!alive) break;
        __threadfence_block();
        const int item = R.item;
        const int slot = slot0 + item / nRegions, region = item % nRegions;
        const YkSlotDev& S = slots[slot];
        const int bx = region % S.nbx, by = region / S.nbx;
        YkaView V;
        V.pix = pixAll + j * 3 * YK_PIXTILE; V.R = &R; V.X0 = bx * 64; V.Y0 = by * 64; V.w = S.w; V.h = S.h; V.yOrg = S.y0; V.rej = run.rejectFactor;
        yka_macro_tile(S, V, run, sh.pretestTab, hist, sMagic, m);
        __threadfence_block();
        __syncwarp();
        int d = 0;
        if (lane == 0) d = atomicAdd(&R.done, 1);
        d = __shfl_sync(YK_FULL, d, 0);
        if (d == 15) {
            __threadfence_block();
            yka_finalize(S, V, run, bx, by, run.doAlpha && S.nPlanes == 4);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) yk_stvi(&R.freeSeq, n + 1);
        }
    }
}

// A new launch on a state that already holds claims: every touched lattice point becomes "claimed before" (bit 31).
__global__ void __launch_bounds__(256)
yk_k_fold_touch(const YkSlotDev* __restrict__ slots, int slot0, int nWords) {
    const YkSlotDev& S = slots[slot0 + blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nWords) { uint32_t v = S.touchMap[i]; if (v && v != 0x80000000u) S.touchMap[i] = 0x80000000u; }
}

// ------------------------------------------------------------------------------------------------------------------
int yk_analyze_setup(int* numSMs) {
#ifdef YK_EMULATE
    if (numSMs) *numSMs = 2;
    return 0;
#else
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    if (numSMs) *numSMs = n;
    return (int)cudaFuncSetAttribute(yk_k_analyze, cudaFuncAttributeMaxDynamicSharedMemorySize, YKA_SMEM_BYTES);
#endif
}
void yk_launch_analyze(const YkSlotDev* slotsDev, int slot0, int nSlots, int nRegions, int gridCtas, const YkRun& run, cudaStream_t st) {
    const int total = nSlots * nRegions;
    const int grid = gridCtas < total ? gridCtas : total;
    YK_LAUNCH(yk_k_analyze, dim3(grid), dim3(YKA_THREADS), YKA_SMEM_BYTES, st, slotsDev, slot0, nSlots, nRegions, run);
}
void yk_launch_fold_touch(const YkSlotDev* slotsDev, int slot0, int nSlots, int nWords, cudaStream_t st) {
    YK_LAUNCH(yk_k_fold_touch, dim3((nWords + 255) / 256, nSlots), dim3(256), 0, st, slotsDev, slot0, nWords);
}
