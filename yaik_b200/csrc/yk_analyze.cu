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
// Structure: one CTA per SM, persistent.  One producer warp takes regions by ticket and issues TMA box loads
// (cp.async.bulk.tensor, one per plane) of one macro-tile row of a region (17 x 68 int32 samples per colour plane, 16 x 64
// alpha) into a ring of raw staging buffers, completion on an mbarrier per buffer.  23 consumer warps take
// (region, macro tile) items from a block-local queue: a warp packs its macro tile's 17x17x3 samples to bytes into a
// warp-private tile (clamped the way Plane::GetPixelValue clamps), tests its alpha tile, releases the raw buffer, then
// runs the cascade and the range stage out of the private tile.  Load latency and the very uneven cost of macro tiles
// overlap; the warp that finishes the last macro tile of a region writes the region's results.  No block-wide barrier
// after start-up; every wait is an mbarrier try_wait (the hardware suspends the warp).
//
// No tensor cores: the work is integer min/max reductions over bytes, bounded by HBM and the integer pipes.
#include "yk_device.h"

#define YKA_NR 8                    // raw int32 staging buffers (TMA destinations), one macro-tile row of a region each
#define YKA_NBS 16                  // region states in flight
#define YKA_CONS_WARPS 23
#define YKA_THREADS ((YKA_CONS_WARPS + 1) * 32)
#define YKA_RAW_PLANE_INTS 1184     // 17 rows x 68 ints = 1156, rounded so every plane starts 128-byte aligned
#define YKA_RAW_STAGE_INTS (3 * YKA_RAW_PLANE_INTS + 16 * 64)
#define YKA_COLOR_TX (3u * YK_RAW_ROWS * YK_RAW_PITCH * 4u)
#define YKA_ALPHA_TX (16u * 64u * 4u)
#define YKP_RS 24                   // row pitch in bytes of a warp-private 17x17 byte tile
#define YKP_CH (17 * YKP_RS)        // bytes of one channel of it
#define YKP_TILE 1232               // 3 channels, rounded to a multiple of 16
// a consumer waits on the phase parity of a raw buffer's barrier: the items in flight (one per consumer warp, consecutive
// in the queue, four per raw buffer) must span fewer raw buffers than the ring holds
static_assert((YKA_CONS_WARPS + 2) / 4 + 2 <= YKA_NR, "work items in flight must not wrap the raw ring");
static_assert(YKA_CONS_WARPS <= 16 * (YKA_NBS - 1), "work items in flight must not wrap the region-state ring");

struct YkaRegion {                  // state of one region being analysed
    uint32_t cell[16];              // claimed 4x4 cells, one 16-bit row per entry (cells outside the image count as claimed)
    uint32_t bits[YK_NPASS][8];     // accept bits of the region in swizzled order (EC.cpp:4026)
    int      stat[YK_NPASS][YK_ST_STRIDE];
    uint32_t touch[17 * 17];        // touch words of the region's lattice points
    uint32_t alpha;                 // bit ty*4+tx: 16x16 tile has a non-zero alpha sample
    int      done;                  // macro tiles finished
    int      slot;                  // absolute slot index of the image
    int      bx, by, X0, Y0, w, h, yOrg, latW, latH, lastX, lastY, doAlpha;    // of the region / its image
    const int32_t* rowBelow[3];
    uint8_t*  latRGB;
    uint8_t*  r2Raw[3];
    uint32_t* r2RawType[3];
    int*      hdr;
};

struct YkaShared {
    uint32_t pretestTab[41];
    int      queueHead;             // next (region sequence number * 16 + macro tile) to hand out
    int      endSeq;                // first region sequence number that does not exist
    YkRun    run;
    unsigned long long rawFull[YKA_NR];     // raw buffer i: its TMA boxes have landed (and the region state is initialised)
    unsigned long long rawFree[YKA_NR];     // raw buffer i: the four macro tiles of the row have been packed out of it
    unsigned long long freed[YKA_NBS];      // region state j: its region has been finalised
    YkaRegion reg[YKA_NBS];
};

#define YKA_SMEM_RAW   (YKA_NR * YKA_RAW_STAGE_INTS * 4)
#define YKA_SMEM_PIX   ((YKA_CONS_WARPS * YKP_TILE + 127) / 128 * 128)
#define YKA_SMEM_HIST  (YKA_CONS_WARPS * 3 * 256)
#define YKA_SMEM_BYTES (YKA_SMEM_RAW + YKA_SMEM_PIX + YKA_SMEM_HIST + 1024 + (int)sizeof(YkaShared))
static_assert(YKA_SMEM_PIX % 128 == 0 && (YKA_CONS_WARPS * 3 * 256) % 128 == 0, "shared-memory carving keeps 128-byte alignment");


// ------------------------------------------------------------------------------------------------------------------
// mbarrier / TMA / shared-flag primitives (inline PTX), with stand-ins for the CPU logic emulation (tests/emu)
#ifdef YK_EMULATE
static inline void yka_mbar_init(unsigned long long* b, int count) { __atomic_store_n(b, (unsigned long long)count << 32, __ATOMIC_SEQ_CST); }
static inline void yka_mbar_expect_tx(unsigned long long*, unsigned) {}
static inline bool yka_mbar_try_wait(unsigned long long* b, unsigned parity) {  // emulated barrier: count << 32 | pending << 16 | completed phases
    if ((__atomic_load_n(b, __ATOMIC_SEQ_CST) & 1ull) != parity) return true;
    std::this_thread::yield();
    return false;
}
static inline void yka_mbar_wait(unsigned long long* b, unsigned parity) { while (!yka_mbar_try_wait(b, parity)) {} }
static inline void yka_mbar_arrive(unsigned long long* b) {
    for (;;) {
        unsigned long long v = __atomic_load_n(b, __ATOMIC_SEQ_CST);
        const unsigned long long count = v >> 32, pending = ((v >> 16) & 0xFFFFull) + 1ull, phases = v & 0xFFFFull;
        const unsigned long long nv = pending == count ? ((count << 32) | ((phases + 1ull) & 0xFFFFull)) : ((count << 32) | (pending << 16) | phases);
        if (__atomic_compare_exchange_n(b, &v, nv, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return;
    }
}
static inline void yka_tma_box(void* dst, const YkTmap* tm, int x, int y, int bw, int bh, unsigned long long* b, int last) {
    const int32_t* base = (const int32_t*)tm->opaque[0]; const int w = (int)tm->opaque[1], h = (int)tm->opaque[2];
    int32_t* d = (int32_t*)dst;
    for (int r = 0; r < bh; r++) for (int c = 0; c < bw; c++)
        d[r * bw + c] = (y + r < h && x + c < w) ? base[(size_t)(y + r) * w + x + c] : 0;
    if (last) yka_mbar_arrive(b);                                               // all boxes of the stage have landed: the phase completes
}
static inline void yka_fence_async() {}
static inline void yka_tmap_acquire(const YkTmap*) {}
static inline int  yka_flag_ld(const int* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static inline void yka_flag_st(int* p, int v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
#else
static __device__ __forceinline__ uint32_t yka_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void yka_mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(yka_s32(b)), "r"(count) : "memory");
}
static __device__ __forceinline__ void yka_mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(yka_s32(b)), "r"(bytes) : "memory");
}
// potentially blocking: the warp is suspended by the hardware until the phase completes or a time limit passes
static __device__ __forceinline__ bool yka_mbar_try_wait(unsigned long long* b, unsigned parity) {
    unsigned ok = 0;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(yka_s32(b)), "r"(parity) : "memory");
    return ok != 0;
}
static __device__ __forceinline__ void yka_mbar_wait(unsigned long long* b, unsigned parity) { while (!yka_mbar_try_wait(b, parity)) {} }
static __device__ __forceinline__ void yka_mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(yka_s32(b)) : "memory");
}
static __device__ __forceinline__ void yka_tma_box(void* dst, const YkTmap* tm, int x, int y, int, int, unsigned long long* b, int) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(yka_s32(dst)), "l"((unsigned long long)tm), "r"(x), "r"(y), "r"(yka_s32(b)) : "memory");
}
static __device__ __forceinline__ void yka_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// the descriptors live in global memory and are rewritten by host copies between launches
static __device__ __forceinline__ void yka_tmap_acquire(const YkTmap* tm) {
    asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" :: "l"((unsigned long long)tm) : "memory");
}
// flags in shared memory (LDS / STS, not generic system-scope accesses)
static __device__ __forceinline__ int yka_flag_ld(const int* p) {
    int v; asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(yka_s32(p)) : "memory"); return v;
}
static __device__ __forceinline__ void yka_flag_st(int* p, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" :: "r"(yka_s32(p)), "r"(v) : "memory");
}
#endif

// ------------------------------------------------------------------------------------------------------------------
static __device__ __forceinline__ unsigned yka_pack4(int4 v) {      // low bytes of four samples -> one word (3 PRMT)
    return __byte_perm(__byte_perm((unsigned)v.x, (unsigned)v.y, 0x0040), __byte_perm((unsigned)v.z, (unsigned)v.w, 0x0040), 0x5410);
}
static __device__ __forceinline__ int yka_byte(unsigned word, int k) { return (int)__byte_perm(word, 0u, 0x4440u | (unsigned)k); }

// table entry of tile ti (0..40) of a macro tile: offX | offY << 4 | shx << 8 | shy << 11 | cell << 14
static __device__ __forceinline__ uint32_t yka_pretest_entry(int ti) {
    const int pid = (ti >= 1) + (ti >= 3) + (ti >= 5) + (ti >= 9) + (ti >= 17) + (ti >= 25);
    const int t = ti - (int)((0x19110905030100ull >> (8 * pid)) & 255ull);
    const int shx = (0x2233344 >> (4 * pid)) & 15, shy = (0x2323434 >> (4 * pid)) & 15;
    const int tx = t & ((16 >> shx) - 1), ty = t >> (4 - shx);
    const int offX = tx << shx, offY = ty << shy;
    return (uint32_t)(offX | (offY << 4) | (shx << 8) | (shy << 11) | (((offY >> 2) * 4 + (offX >> 2)) << 14));
}

// Producer warp, once per region: claimed cells and the fields the consumers need from the slot descriptor.
static __device__ void yka_region_init(const YkSlotDev& S, YkaRegion& R, int slot, int bx, int by, bool doAlpha, int lane) {
    const int X0 = bx * 64, Y0 = by * 64;
    if (lane < 16) {
        const int cy = (Y0 >> 2) + lane;
        uint32_t v = 0xFFFFu;
        if (cy * 4 < S.h) {
            v = S.cellMask[(size_t)cy * S.nbx + bx];
            const int cellsIn = (S.w - X0) >> 2;
            if (cellsIn < 16) v |= (0xFFFFu << cellsIn) & 0xFFFFu;
        }
        R.cell[lane] = v;
    } else if (lane == 16) {
        R.slot = slot; R.bx = bx; R.by = by; R.X0 = X0; R.Y0 = Y0; R.w = S.w; R.h = S.h; R.yOrg = S.y0;
        R.latW = S.latW; R.latH = S.latH; R.lastX = bx == S.nbx - 1; R.lastY = by == S.nby - 1; R.doAlpha = doAlpha;
        R.latRGB = S.latRGB; R.hdr = S.hdr;
    } else if (lane >= 20 && lane < 23) {
        const int p = lane - 20;
        R.r2Raw[p] = S.r2Raw[p]; R.r2RawType[p] = S.r2RawType[p]; R.rowBelow[p] = S.rowBelow[p];
    }
}

// Consumer warp: the 17x17 samples of macro tile (mx, row k) of the region, three colour planes, from the raw int32
// rows staged by TMA into the warp-private byte tile, clamped the way Plane::GetPixelValue clamps (framework.h:116-121);
// then the alpha-zero test of the 16x16 tile (EC.cpp:357-430 restated per tile).
static __device__ __forceinline__ void yka_pack_macro_tile(const int32_t* __restrict__ raw, uint8_t* __restrict__ priv, YkaRegion& R, int mx, int k) {
    const int lane = threadIdx.x & 31;
    const int w = R.w, h = R.h;
    const int Xm = R.X0 + 16 * mx, Yk = R.Y0 + 16 * k;
    unsigned bad = 0;
    if (Xm + 20 <= w && Yk + YK_RAW_ROWS <= h) {
        // interior: 17 rows x 5 int4 (columns 0..19 of the macro tile, 17 needed) per plane, all inside the image
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int4* __restrict__ src = reinterpret_cast<const int4*>(raw + c * YKA_RAW_PLANE_INTS) + 4 * mx;
            uint8_t* __restrict__ d = priv + c * YKP_CH;
#pragma unroll
            for (int it = 0; it < 3; it++) {
                const int idx = it * 32 + lane;                 // 85 = 17 rows x 5
                if (idx < 85) {
                    const int lr = idx / 5, q = idx - lr * 5;
                    const int4 v = src[lr * (YK_RAW_PITCH / 4) + q];
                    bad |= (unsigned)(v.x | v.y | v.z | v.w);
                    *reinterpret_cast<unsigned*>(d + lr * YKP_RS + 4 * q) = yka_pack4(v);
                }
            }
        }
    } else if (Yk < h && Xm < w) {
        // at the right / bottom edge: clamp inside the image; in strip mode the row under the strip is the real image
        // row (rowBelow), not a clamp
        const int wmax = w - 1 - R.X0, hmax = min(16, h - 1 - Yk);
        for (int idx = lane; idx < 17 * 17; idx += 32) {
            const int lr = idx / 17, lx = idx - lr * 17;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                int s;
                if (lr > hmax && R.rowBelow[c]) s = __ldg(R.rowBelow[c] + min(Xm + lx, w - 1));
                else s = raw[c * YKA_RAW_PLANE_INTS + min(lr, hmax) * YK_RAW_PITCH + min(16 * mx + lx, wmax)];
                bad |= (unsigned)s;
                priv[c * YKP_CH + lr * YKP_RS + lx] = (uint8_t)s;
            }
        }
    }
    if (bad & ~255u) atomicOr(&R.hdr[YK_HD_ERR], 1);
    if (R.doAlpha) {
        // samples outside the image arrive as zeros
        const int4* __restrict__ a = reinterpret_cast<const int4*>(raw + 3 * YKA_RAW_PLANE_INTS) + 4 * mx;
        const int4 v0 = a[(lane >> 2) * 16 + (lane & 3)], v1 = a[((lane >> 2) + 8) * 16 + (lane & 3)];
        const bool nz = (v0.x | v0.y | v0.z | v0.w | v1.x | v1.y | v1.z | v1.w) != 0;
        if (__any_sync(YK_FULL, nz) && lane == 0) atomicOr(&R.alpha, 1u << (k * 4 + mx));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Cheap rejection of all 41 tiles (1 + 2 + 2 + 4 + 8 + 8 + 16 over the seven shapes) of a 16x16 macro tile: lane = tile,
// one channel, the quad at the tile centre, raw corners.  A pixel whose raw-family U is outside [loWide, hiWide) cannot
// be accepted by any of the six variants (a family's corners differ from the raw ones by -3..+4), so a cleared bit is a
// proven rejection; a set bit only means "run the real test".  Bit (start(pid) + t) belongs to tile t of pass id pid.
static __device__ __forceinline__ unsigned long long yka_pretest(const uint8_t* __restrict__ priv, const uint32_t* sTab, int wIn, int hIn,
                                                                 unsigned claimed, int R) {
    const int lane = threadIdx.x & 31;
    unsigned long long P = 0;
#pragma unroll
    for (int round = 0; round < 2; round++) {
        const int ti = lane + 32 * round;
        bool possible = false;
        if (ti < 41) {
            const uint32_t e = sTab[ti];
            const int shx = (e >> 8) & 7, shy = (e >> 11) & 7, sh = shx + shy;
            const int lx0 = e & 15, ly0 = (e >> 4) & 15, TW = 1 << shx, TH = 1 << shy, N = 1 << sh;
            if (!((claimed >> (e >> 14)) & 1u) && lx0 + TW <= wIn && ly0 + TH <= hIn) {
                const uint8_t* p = priv + ly0 * YKP_RS + lx0;
                const int tl = p[0], tr = p[TW], bl = p[TH * YKP_RS], br = p[TH * YKP_RS + TW];
                const int dx0 = (TW >> 1) & ~3, dy = TH >> 1;
                const unsigned word = *reinterpret_cast<const unsigned*>(p + dy * YKP_RS + dx0);
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
    const int s1 = s + step, s2 = s1 + step, s3 = s2 + step;
    const int u0 = yka_byte(word, 0) * negN + s;
    const int u1 = yka_byte(word, 1) * negN + s1;
    const int u2 = yka_byte(word, 2) * negN + s2;
    const int u3 = yka_byte(word, 3) * negN + s3;
    umin = __vimin3_s32(umin, u0, u1); umin = __vimin3_s32(umin, u2, u3);
    umax = __vimax3_s32(umax, u0, u1); umax = __vimax3_s32(umax, u2, u3);
}

struct YkaTile {                    // per lane: the tile it belongs to in the current (sub-)pass
    int shx, shy, sh, negN, dx0, dy, R;
    bool two;
};

// one channel of one corner family: this lane's one or two quads against the bilinear fit of (tl, tr, bl, br)
static __device__ __forceinline__ void yka_channel(const YkaTile& T, unsigned w0, unsigned w1, int tl, int tr, int bl, int br, int& umin, int& umax) {
    const int B = (tr - tl) << T.shy, C = (bl - tl) << T.shx, D = tl - tr - bl + br;
    const int step = B + D * T.dy;
    const int s = ((tl + T.R) << T.sh) + B * T.dx0 + T.dy * (C + D * T.dx0);
    yka_quad(w0, s, step, T.negN, umin, umax);
    if (T.two) yka_quad(w1, s + 4 * step, step, T.negN, umin, umax);
}

// The accept test of FittingQuadSmooth (EC.cpp:3810-3998) for the tile this lane belongs to.  The lanes in `gmask` share
// the tile; each holds one or two quads of one pixel row of it in `wd` (quad 0 at dx0, quad 1 at dx0 + 4, row dy).
// Returns, uniformly over the group, whether any of the six variants (3 corner families x rounded / truncated) keeps
// every pixel of every channel within the reject factor.
static __device__ __forceinline__ bool yka_tile_test(const uint8_t* __restrict__ corner, const YkaTile& T, const unsigned (&wd)[3][2],
                                                     unsigned gmask, bool active) {
    const int N = 1 << T.sh, TW = 1 << T.shx, THp = YKP_RS << T.shy;
    const int hiT = (2 * T.R + 1) * N;                  // |cur - S/N| <= R            <=>  0 <= U < hiT
    const int loR = -(N / 2 - 1);                       // |cur - (S+N/2-1)/N| <= R    <=>  loR <= U < hiT + loR
    const int loWide = -(4 * N + N / 2 - 1), hiWide = hiT + 3 * N;
    int cr[3][4];                                       // TL TR BL BR, clamped at the image edge by the packing (EC.cpp:3845-3868)
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const uint8_t* p = corner + c * YKP_CH;
        cr[c][0] = p[0]; cr[c][1] = p[TW]; cr[c][2] = p[THp]; cr[c][3] = p[THp + TW];
    }
    bool resolved = !active;
    // ---- raw corners
    int umin = INT_MAX, umax = INT_MIN;
    yka_channel(T, wd[0][0], wd[0][1], cr[0][0], cr[0][1], cr[0][2], cr[0][3], umin, umax);
    {   // one channel with the raw corners proves most non-gradient tiles hopeless for every variant
        const unsigned bH = __ballot_sync(YK_FULL, (umin < loWide) || (umax >= hiWide));
        if (bH & gmask) resolved = true;
        if (!__any_sync(YK_FULL, !resolved)) return false;
    }
    yka_channel(T, wd[1][0], wd[1][1], cr[1][0], cr[1][1], cr[1][2], cr[1][3], umin, umax);
    yka_channel(T, wd[2][0], wd[2][1], cr[2][0], cr[2][1], cr[2][2], cr[2][3], umin, umax);
    const unsigned bT = __ballot_sync(YK_FULL, (umin < 0) || (umax >= hiT));
    const unsigned bR = __ballot_sync(YK_FULL, (umin < loR) || (umax >= hiT + loR));
    const unsigned bH = __ballot_sync(YK_FULL, (umin < loWide) || (umax >= hiWide));
    bool accepted = !resolved && (((bT & gmask) == 0u) || ((bR & gmask) == 0u));     // EC.cpp:3998: any surviving variant accepts
    if (accepted || (bH & gmask)) resolved = true;                                   // no family can accept a tile outside the wide bounds
    if (!__any_sync(YK_FULL, !resolved)) return accepted;
    // ---- Round6 corners
    {
        int umin1 = INT_MAX, umax1 = INT_MIN;
#pragma unroll
        for (int c = 0; c < 3; c++)
            yka_channel(T, wd[c][0], wd[c][1], yk_round6(cr[c][0]), yk_round6(cr[c][1]), yk_round6(cr[c][2]), yk_round6(cr[c][3]), umin1, umax1);
        const unsigned bT1 = __ballot_sync(YK_FULL, (umin1 < 0) || (umax1 >= hiT));
        const unsigned bR1 = __ballot_sync(YK_FULL, (umin1 < loR) || (umax1 >= hiT + loR));
        if (!resolved && (((bT1 & gmask) == 0u) || ((bR1 & gmask) == 0u))) { accepted = true; resolved = true; }
        if (!__any_sync(YK_FULL, !resolved)) return accepted;
    }
    // ---- Round6P corners
    {
        int umin2 = INT_MAX, umax2 = INT_MIN;
#pragma unroll
        for (int c = 0; c < 3; c++)
            yka_channel(T, wd[c][0], wd[c][1], yk_round6p(cr[c][0]), yk_round6p(cr[c][1]), yk_round6p(cr[c][2]), yk_round6p(cr[c][3]), umin2, umax2);
        const unsigned bT2 = __ballot_sync(YK_FULL, (umin2 < 0) || (umax2 >= hiT));
        const unsigned bR2 = __ballot_sync(YK_FULL, (umin2 < loR) || (umax2 >= hiT + loR));
        if (!resolved && (((bT2 & gmask) == 0u) || ((bR2 & gmask) == 0u))) accepted = true;
    }
    return accepted;
}

// Side effects of an accepted tile (EC.cpp:3998-4132) that are local to the region: bitmap bit, TileDone / bounding box,
// the four lattice points it touches with its role at each.  (lx0, ly0) = tile origin inside the region.
static __device__ __forceinline__ void yka_commit(YkaRegion& R, const YkGeomS& g, int pid, int rp, int lx0, int ly0) {
    const int TW = 1 << g.shx, TH = 1 << g.shy;
    const int sub = (ly0 >> g.lbh) * (64 >> g.lbw) + (lx0 >> g.lbw);
    const int li = sub * g.bits + (((ly0 & ((1 << g.lbh) - 1)) >> g.shy) << (g.lbw - g.shx)) + ((lx0 & ((1 << g.lbw) - 1)) >> g.shx);
    atomicOr(&R.bits[pid][li >> 5], 1u << (li & 31));                          // EC.cpp:4026
    atomicAdd(&R.stat[pid][YK_ST_TILEDONE], 1);                                // EC.cpp:4039-4044 (mins stored as extent - value)
    atomicMax(&R.stat[pid][YK_ST_MINX], R.w - (R.X0 + lx0));
    atomicMax(&R.stat[pid][YK_ST_MINY], INT_MAX / 2 - (R.yOrg + R.Y0 + ly0));
    atomicMax(&R.stat[pid][YK_ST_MAXX], R.X0 + lx0 + TW);
    atomicMax(&R.stat[pid][YK_ST_MAXY], R.yOrg + R.Y0 + ly0 + TH);
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
static __device__ __forceinline__ unsigned yka_macro_pass(const uint8_t* __restrict__ priv, YkaRegion& Rg, int pid, int rp, int mlx, int mly,
                                                          unsigned claimed, unsigned poss, const uint2 (&pw)[3], int rej) {
    const YkGeomS g = yk_geom_s(pid);
    const int lane = threadIdx.x & 31, row = lane >> 1, half = lane & 1;
    const int shx = g.shx, shy = g.shy, TH = 1 << shy;
    const int ty = row >> shy, lyT = ty << shy;
    const unsigned rowMask = (shy == 4) ? YK_FULL : (((1u << (2 * TH)) - 1u) << (2 * lyT));
    const unsigned gmask = (shx == 4) ? rowMask : (rowMask & (0x55555555u << half));
    const bool leader = lane == __ffs((int)gmask) - 1;
    const int nSub = (shx == 2) ? 2 : 1;
    YkaTile T;
    T.shx = shx; T.shy = shy; T.sh = shx + shy; T.negN = -(1 << T.sh); T.dy = row - lyT; T.R = rej; T.two = shx != 2;
    unsigned newCells = 0;
    for (int sub = 0; sub < nSub; sub++) {
        const int tx = (shx == 2) ? (2 * half + sub) : ((8 * half) >> shx);
        const int lxT = tx << shx;
        T.dx0 = (shx == 2) ? 0 : (8 * half - lxT);
        const int t = ty * (16 >> shx) + tx;
        const int cellX = lxT >> 2, cellY = lyT >> 2;
        const bool active = ((poss >> t) & 1u) && !((claimed >> (cellY * 4 + cellX)) & 1u);      // EC.cpp:3818, 3826, 3871-3875
        if (!__any_sync(YK_FULL, active)) continue;
        unsigned wd[3][2];
#pragma unroll
        for (int c = 0; c < 3; c++) { wd[c][0] = (shx == 2 && sub) ? pw[c].y : pw[c].x; wd[c][1] = pw[c].y; }
        const bool acc = yka_tile_test(priv + lyT * YKP_RS + lxT, T, wd, gmask, active);
        unsigned mine = 0;
        if (acc && leader) {
            yka_commit(Rg, g, pid, rp, mlx + lxT, mly + lyT);
            // EC.cpp:4029-4037: the tile's cells become claimed
            const unsigned cols = ((1u << (1 << (shx - 2))) - 1u) << cellX;
            const unsigned rowsPat = (0x1111u & ((1u << (4 << (shy - 2))) - 1u)) << (4 * cellY);
            mine = cols * rowsPat;
        }
        newCells |= __reduce_or_sync(YK_FULL, mine);
    }
    return claimed | newCells;
}

// DynamicTileCompressor (EC.cpp:8398-8522) for the 8x8 tile at (lx8, ly8) of the macro tile; q = its quadrants to code
// (bit0 TL, 1 TR, 2 BL, 3 BR: top-left map pixel 0, EC.cpp:8420-8430 == 4x4 cell unclaimed).  Lane = two pixels; the
// three planes side by side.  Output goes to the tile's fixed place in r2Raw / r2RawType.
static __device__ __forceinline__ void yka_range_tile(const uint8_t* __restrict__ priv, const YkaRegion& Rg, uint8_t* __restrict__ hist,
                                                      const uint32_t* __restrict__ magicTab, size_t tile, int lx8, int ly8, unsigned q) {
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
        const unsigned short two = *reinterpret_cast<const unsigned short*>(priv + p * YKP_CH + (ly8 + r) * YKP_RS + lx8 + c0);
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
            *reinterpret_cast<uint16_t*>(Rg.r2Raw[p] + tile * 64 + pos) = (uint16_t)((bxv & 255) | ((byv & 255) << 8));
        }
        if (lane == p) Rg.r2RawType[p][tile] = (uint32_t)color0 | ((uint32_t)minCol << 8) | ((uint32_t)delta << 16);     // EC.cpp:8503-8505
    }
    __syncwarp();       // the histogram entries are clean again before the next tile fills them
}

// One (region, macro tile) work item after its pixels have been packed: the cascade of Convert()'s passes
// (EC.cpp:9057-9093), the corner colours of its lattice points, then the range stage.
static __device__ void yka_macro_tile(const uint8_t* __restrict__ priv, YkaRegion& R, const YkRun& run, const uint32_t* sTab, uint8_t* hist,
                                      const uint32_t* magicTab, int m) {
    const int lane = threadIdx.x & 31;
    const int mx = m & 3, my = m >> 2, mlx = 16 * mx, mly = 16 * my;
    unsigned claimed = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) claimed |= ((R.cell[my * 4 + r] >> (4 * mx)) & 15u) << (4 * r);
    const unsigned claimed0 = claimed;
    const int nPasses = run.nPasses;
    const int wIn = R.w - R.X0 - mlx, hIn = R.h - R.Y0 - mly;       // image extent seen from the macro tile's origin
    if (nPasses > 0 && claimed != 0xFFFFu) {
        uint2 pw[3];
#pragma unroll
        for (int c = 0; c < 3; c++)
            pw[c] = *reinterpret_cast<const uint2*>(priv + c * YKP_CH + (lane >> 1) * YKP_RS + 8 * (lane & 1));
        // the 16x16 pass usually runs first and straight away (most macro tiles of illustration-like content end there);
        // the other shapes are pre-tested together, once, the first time one of them comes up
        unsigned long long P = (wIn >= 16 && hIn >= 16) ? 1ull : 0ull;
        bool pretested = false;
        const int rej = run.rejectFactor;
        for (int rp = 0; rp < nPasses && claimed != 0xFFFFu; rp++) {
            const int pid = run.passId[rp];
            if (pid != 0 && !pretested) { P = yka_pretest(priv, sTab, wIn, hIn, claimed, rej); pretested = true; }
            const YkGeomS g = yk_geom_s(pid);
            const unsigned poss = (unsigned)(P >> g.start) & ((1u << (256 >> (g.shx + g.shy))) - 1u);
            if (poss) claimed = yka_macro_pass(priv, R, pid, rp, mlx, mly, claimed, poss, pw, rej);
        }
        if (lane < 4 && claimed != claimed0) atomicOr(&R.cell[my * 4 + lane], ((claimed >> (4 * lane)) & 15u) << (4 * mx));
    }
    if (nPasses > 0 && lane < 25) {
        // corner colours at the 4-pixel lattice points of the macro tile (what an accepted tile would emit, EC.cpp:4115-4132);
        // the points on its right / bottom edge belong to the next macro tile unless the image ends there
        const int jj = lane / 5, i = lane - jj * 5;
        const int gx = ((R.X0 + mlx) >> 2) + i, gy = ((R.Y0 + mly) >> 2) + jj;
        if (wIn > 0 && hIn > 0 && (i < 4 || wIn <= 16) && (jj < 4 || hIn <= 16) && gx < R.latW && gy < R.latH) {
            uint8_t* d = R.latRGB + ((size_t)gy * R.latW + gx) * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) d[c] = (uint8_t)yk_compress250(yk_round6(priv[c * YKP_CH + (4 * jj) * YKP_RS + 4 * i]));
        }
    }
    if (run.doR2 && claimed != 0xFFFFu) {
        const int tilesW = R.w >> 3;
#pragma unroll 1
        for (int t8 = 0; t8 < 4; t8++) {
            const int qx = t8 & 1, qy = t8 >> 1;
            const unsigned c4 = claimed >> (8 * qy + 2 * qx);
            const unsigned q = (~((c4 & 3u) | (((c4 >> 4) & 3u) << 2))) & 15u;
            if (q) {
                const size_t tile = (size_t)((R.Y0 + mly + 8 * qy) >> 3) * tilesW + ((R.X0 + mlx + 8 * qx) >> 3);
                yka_range_tile(priv, R, hist, magicTab, tile, 8 * qx, 8 * qy, q);
            }
        }
    }
}

// Results of a finished region, by one warp: accept bitmaps, claimed cells, touch words of its lattice points, alpha
// tiles, per-pass counters.  Leaves the region state zeroed for the next region that uses it.
static __device__ void yka_finalize(const YkSlotDev& S, YkaRegion& R, const YkRun& run) {
    const int lane = threadIdx.x & 31;
    const int w = R.w, h = R.h, X0 = R.X0, Y0 = R.Y0, nbx = S.nbx, bx = R.bx;
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
    // touch words of the lattice points (interior points are exclusive to the region, border points are shared)
    if (run.nPasses > 0) {
        const int latW = R.latW, latH = R.latH;
        uint32_t* __restrict__ touchMap = S.touchMap;
        for (int idx = lane; idx < 17 * 17; idx += 32) {
            const uint32_t tv = R.touch[idx];
            if (tv) {
                const int jj = idx / 17, i = idx - jj * 17;
                const int gx = (X0 >> 2) + i, gy = (Y0 >> 2) + jj;
                if (gx < latW && gy < latH) atomicOr(&touchMap[(size_t)gy * latW + gx], tv);
                R.touch[idx] = 0;
            }
        }
    }
    if (R.doAlpha) {
        const int tx = lane & 3, ty = (lane >> 2) & 3;
        const int px = X0 + 16 * tx, py = Y0 + 16 * ty;
        const bool in = lane < 16 && px < w && py < h;
        const bool kept = in && ((R.alpha >> lane) & 1u);
        if (in) S.alphaKept[(size_t)(py >> 4) * ((w + 15) >> 4) + (px >> 4)] = kept ? 1 : 0;
        // bounding box of kept tiles (EC.cpp:416-422), mins stored as extent - value so the header can be memset to 0
        const int big = INT_MAX / 2;
        const int mnx = __reduce_max_sync(YK_FULL, kept ? w - px : 0);
        const int mny = __reduce_max_sync(YK_FULL, kept ? big - (R.yOrg + py) : 0);
        const int mxx = __reduce_max_sync(YK_FULL, kept ? min(px + 16, w) : 0);
        const int mxy = __reduce_max_sync(YK_FULL, kept ? R.yOrg + min(py + 16, h) : 0);
        const int cnt = __popc(__ballot_sync(YK_FULL, kept));
        if (lane == 0 && cnt) {
            atomicMax(&R.hdr[YK_HD_ALPHA_MINX], mnx); atomicMax(&R.hdr[YK_HD_ALPHA_MINY], mny);
            atomicMax(&R.hdr[YK_HD_ALPHA_MAXX], mxx); atomicMax(&R.hdr[YK_HD_ALPHA_MAXY], mxy);
            atomicAdd(&R.hdr[YK_HD_ALPHA_KEPT], cnt);
        }
    }
    if (lane < YK_NPASS) {
        const int pid = lane;
        if (R.stat[pid][YK_ST_TILEDONE] > 0) {
            int* d = R.hdr + YK_HD_PASS0 + pid * YK_ST_STRIDE;
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

#ifdef YK_TIMING
__device__ unsigned long long yk_timing[16];
#define YKT_DECL long long t__ = clock64(), t2__
#define YKT(i) (t2__ = clock64(), atomicAdd(&yk_timing[i], (unsigned long long)(t2__ - t__)), t__ = t2__)
extern "C" void yk_debug_timing(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, yk_timing, sizeof(yk_timing));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(yk_timing, z, sizeof(z)); }
}
#else
#define YKT_DECL
#define YKT(i)
#endif

__global__ void __launch_bounds__(YKA_THREADS, 1)
yk_k_analyze(const YkSlotDev* __restrict__ slots, int slot0, int nSlots, int nRegions, YkRun runArg) {
    // one dynamic shared-memory block, carved by constant offsets from the array itself so that every access stays in
    // the shared address space (LDS / STS / ATOMS, no generic loads)
#ifdef YK_EMULATE
    static __align__(128) unsigned char smem[YKA_SMEM_BYTES];
#else
    extern __shared__ __align__(128) unsigned char smem[];
#endif
    int32_t* raw = reinterpret_cast<int32_t*>(smem);
    uint8_t* privAll = smem + YKA_SMEM_RAW;
    uint8_t* histAll = smem + YKA_SMEM_RAW + YKA_SMEM_PIX;
    uint32_t* sMagic = reinterpret_cast<uint32_t*>(smem + YKA_SMEM_RAW + YKA_SMEM_PIX + YKA_SMEM_HIST);    // ceil(2^20 / d): exact floor(n / d) for n < 4112, d <= 255
    YkaShared& sh = *reinterpret_cast<YkaShared*>(smem + YKA_SMEM_RAW + YKA_SMEM_PIX + YKA_SMEM_HIST + 1024);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = nSlots * nRegions;

    // ---- start-up (the only block-wide barriers)
    for (int i = tid; i < (int)(sizeof(YkaShared) / 4); i += YKA_THREADS) reinterpret_cast<uint32_t*>(&sh)[i] = 0;
    for (int i = tid; i < YKA_SMEM_HIST / 4; i += YKA_THREADS) reinterpret_cast<uint32_t*>(histAll)[i] = 0;
    if (tid < 256) sMagic[tid] = tid ? ((1u << 20) + (unsigned)tid - 1u) / (unsigned)tid : 0u;
    __syncthreads();
    if (tid < 41) sh.pretestTab[tid] = yka_pretest_entry(tid);
    if (tid == 64) {
        sh.run = runArg;
        sh.endSeq = INT_MAX;
        for (int i = 0; i < YKA_NR; i++) { yka_mbar_init(&sh.rawFull[i], 1); yka_mbar_init(&sh.rawFree[i], 4); }
        for (int j = 0; j < YKA_NBS; j++) yka_mbar_init(&sh.freed[j], 1);
#ifndef YK_EMULATE
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    __syncthreads();
    const YkRun& run = sh.run;

    if (warp == YKA_CONS_WARPS) {
        // ================================================== producer ==================================================
        // unit u = macro-tile row (u & 3) of this CTA's (u >> 2)-th region; raw buffer u % YKA_NR, region state (u >> 2) % YKA_NBS
        int* ticket = slots[slot0].hdr + YK_HD_TICKET_ANALYZE;
        const bool wantAlpha = run.doAlpha != 0;
        int lastSlot = -1, nextItem = -1;
        // of the current region (lane 0 issues; every lane keeps them so the region state can be initialised together)
        const YkSlotDev* S = nullptr;
        int bx = 0, by = 0, alpha = 0;
        if (lane == 0) { const int t = atomicAdd(ticket, 1); nextItem = t < total ? t : -1; }
        nextItem = __shfl_sync(YK_FULL, nextItem, 0);
        YKT_DECL;
        for (int u = 0;; u++) {
            const int i = u % YKA_NR, n = u >> 2, k = u & 3, j = n % YKA_NBS;
            if (k == 0) {
                // next region: its ticket was fetched one region ahead
                const int item = nextItem;
                if (item < 0) { if (lane == 0) yka_flag_st(&sh.endSeq, n); break; }
                if (lane == 0) { const int t = atomicAdd(ticket, 1); nextItem = t < total ? t : -1; }
                const int slot = slot0 + item / nRegions, region = item % nRegions;
                S = &slots[slot];
                if (slot != lastSlot) {
                    if (lane < S->nPlanes) yka_tmap_acquire(&S->tmap[lane]);
                    lastSlot = slot;
                }
                const int nbx = S->nbx;
                bx = region % nbx; by = region / nbx;
                alpha = wantAlpha && S->nPlanes == 4;
                if (n >= YKA_NBS) yka_mbar_wait(&sh.freed[j], (unsigned)((n / YKA_NBS - 1) & 1));
                yka_region_init(*S, sh.reg[j], slot, bx, by, alpha, lane);
                __threadfence_block();
                __syncwarp();
                if (lane == 0) YKT(0);
            }
            if (u >= YKA_NR) yka_mbar_wait(&sh.rawFree[i], (unsigned)((u / YKA_NR - 1) & 1));
            if (lane == 0) {
                YKT(1);
                int32_t* dst = raw + i * YKA_RAW_STAGE_INTS;
                yka_fence_async();
                yka_mbar_expect_tx(&sh.rawFull[i], YKA_COLOR_TX + (alpha ? YKA_ALPHA_TX : 0u));
                for (int c = 0; c < 3; c++)
                    yka_tma_box(dst + c * YKA_RAW_PLANE_INTS, &S->tmap[c], bx * 64, by * 64 + 16 * k, YK_RAW_PITCH, YK_RAW_ROWS, &sh.rawFull[i], !alpha && c == 2);
                if (alpha) yka_tma_box(dst + 3 * YKA_RAW_PLANE_INTS, &S->tmap[3], bx * 64, by * 64 + 16 * k, 64, 16, &sh.rawFull[i], 1);
                YKT(2);
            }
            if (k == 3) nextItem = __shfl_sync(YK_FULL, nextItem, 0);
        }
        return;
    }

    // ===================================================== consumers =====================================================
    uint8_t* hist = histAll + warp * 3 * 256;
    uint8_t* priv = privAll + warp * YKP_TILE;
    YKT_DECL;
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(&sh.queueHead, 1);
        q = __shfl_sync(YK_FULL, q, 0);
        const int n = q >> 4, m = q & 15, j = n % YKA_NBS;
        const int u = 4 * n + (m >> 2), i = u % YKA_NR;
        YkaRegion& R = sh.reg[j];
        bool alive = true;
        // the macro-tile row of this item has landed (the wait suspends the warp; it wakes up now and then to see
        // whether the CTA has run out of regions)
        while (!yka_mbar_try_wait(&sh.rawFull[i], (unsigned)((u / YKA_NR) & 1))) {
            if (yka_flag_ld(&sh.endSeq) <= n) { alive = false; break; }
        }
        alive = __all_sync(YK_FULL, alive);
        if (tid == 0) YKT(8);
        if (!alive) break;
        yka_pack_macro_tile(raw + i * YKA_RAW_STAGE_INTS, priv, R, m & 3, m >> 2);
        __syncwarp();
        if (lane == 0) yka_mbar_arrive(&sh.rawFree[i]);      // this warp is done with the raw rows
        if (tid == 0) YKT(9);
        yka_macro_tile(priv, R, run, sh.pretestTab, hist, sMagic, m);
        __threadfence_block();
        __syncwarp();
        if (tid == 0) YKT(10);
        int d = 0;
        if (lane == 0) d = atomicAdd(&R.done, 1);
        d = __shfl_sync(YK_FULL, d, 0);
        if (d == 15) {
            __threadfence_block();
            yka_finalize(slots[R.slot], R, run);
            __threadfence_block();
            __syncwarp();
            if (lane == 0) yka_mbar_arrive(&sh.freed[j]);
            if (tid == 0) YKT(11);
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
