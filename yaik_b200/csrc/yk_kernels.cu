// yaik_b200 — hand-written sm_100a kernels of the YAIK encoder-analysis stage.
//
// What the reference computes sequentially, tile after tile in stream order (KLab/YAIK,
// encoder/EncoderContext.cpp = "EC.cpp"), is restated here in order-free form so that one CTA can analyse one
// 64x64 region (the largest swizzle block, include/YAIK_private.h:212-276) independently:
//
//   yk_k_analyze      MipPrefilter/quadRecursion (EC.cpp:1257-1427, 357-430): per 16x16 tile "all alpha == 0"
//                     by ballot; then the accept decision of all FittingQuadSmooth passes (EC.cpp:3810-3998)
//                     of the region, pixels staged once in shared memory as packed bytes.  Accept decisions of
//                     pass k depend only on earlier passes of the same region (tiles nest in 64x64).
//                     One warp runs the whole 7-pass cascade of a 16x16 macro tile without block barriers.
//   yk_k_emit         corner ownership + rgbStream (EC.cpp:4001-4021, 4115-4132): a lattice point is emitted by the
//                     first pass that touches it, by the accepted tile with the smallest stream position; stream
//                     offsets by decoupled look-back over swizzle blocks taken in stream order.
//   yk_k_range1d      DynamicTileCompressor (EC.cpp:8398-8522), one warp per 8-tile segment, offsets by look-back.
//   yk_k_state        expands the compact masks into the reference's int32 state planes (compat download).
//   yk_k_r1_*         DynamicTileEncode (EC.cpp:4365-4503, 747-1212), LUT search at 3/4 bits per pixel.
//
// No tensor cores: the work is integer min/max reductions over bytes, bounded by HBM and the integer pipes.
#include "yk_internal.h"
#include <limits.h>

#define YK_RS 72                 // shared-memory row pitch in bytes of the staged 65x65 byte tile (18 words: conflict-free rows)
#define YK_FULL 0xffffffffu

static __device__ __forceinline__ int yk_round6(int v) { int r = v >> 2; return (r << 2) | (r >> 4); }                 // EC.cpp:3183-3189
static __device__ __forceinline__ int yk_round6p(int v) { v = min(v + 1, 255); int r = v >> 2; return (r << 2) | (r >> 4); }  // EC.cpp:3202-3207
static __device__ __forceinline__ int yk_compress250(int v) { return (v * 250 + 127) / 255; }                          // CompressF(v, colorCompressionQuad), EC.cpp:3191-3194

struct YkGeomC { int shx, shy, bw, bh, bits; };
__constant__ YkGeomC yk_geom_tab[YK_NPASS] = YK_PASS_TABLE;
static __device__ __forceinline__ YkGeomC yk_geom(int pid) { return yk_geom_tab[pid]; }

// stream position (== bitmap bit index) of the tile at global tile coords (gtx, gty), EC.cpp:3801-3828, 4227-4234
static __device__ __forceinline__ int yk_tile_pos(const YkGeomC& g, int w, int gtx, int gty) {
    int x = gtx << g.shx, y = gty << g.shy;
    int nSwzX = (w + g.bw - 1) / g.bw;
    int gb = (y / g.bh) * nSwzX + (x / g.bw);
    return gb * g.bits + (((y % g.bh) >> g.shy) * (g.bw >> g.shx)) + ((x % g.bw) >> g.shx);
}

// ------------------------------------------------------------------------------------------------------------------
// pixel staging: 65x65 samples (the region plus the right/bottom corner row) of the three colour planes, clamped the
// way Plane::GetPixelValue clamps (framework.h:116-121), packed to bytes.
static __device__ __forceinline__ int yk_src(const YkSlotDev& S, int c, int x, int y) {
    x = min(x, S.w - 1);
    if (y >= S.h) {
        if (S.rowBelow[c]) return __ldg(S.rowBelow[c] + x);      // strip mode: the real row below
        y = S.h - 1;
    }
    return __ldg(S.plane[c] + (size_t)y * S.w + x);
}

static __device__ __forceinline__ unsigned yk_pack4(int4 v) {      // low bytes of four samples -> one word (3 PRMT)
    return __byte_perm(__byte_perm((unsigned)v.x, (unsigned)v.y, 0x0040), __byte_perm((unsigned)v.z, (unsigned)v.w, 0x0040), 0x5410);
}

static __device__ void yk_stage_pixels(const YkSlotDev& S, int X0, int Y0, uint8_t (*pix)[65 * YK_RS], unsigned& bad) {
    const int tid = threadIdx.x;
    const int w = S.w, h = S.h;
    for (int c = 0; c < 3; c++) {
        const int32_t* __restrict__ P = S.plane[c];
        int4 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int ly = (tid >> 4) + 16 * k, lx = (tid & 15) * 4;
            int y = Y0 + ly, x = X0 + lx;
            if (y < h && x + 3 < w) {
                v[k] = __ldg(reinterpret_cast<const int4*>(P + (size_t)y * w + x));
            } else {
                v[k].x = yk_src(S, c, x, y); v[k].y = yk_src(S, c, x + 1, y);
                v[k].z = yk_src(S, c, x + 2, y); v[k].w = yk_src(S, c, x + 3, y);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int ly = (tid >> 4) + 16 * k, lx = (tid & 15) * 4;
            bad |= (unsigned)(v[k].x | v[k].y | v[k].z | v[k].w);
            *reinterpret_cast<unsigned*>(&pix[c][ly * YK_RS + lx]) = yk_pack4(v[k]);
        }
        if (tid < 65) {                 // column 64
            int s = yk_src(S, c, X0 + 64, Y0 + tid);
            bad |= (unsigned)s;
            pix[c][tid * YK_RS + 64] = (uint8_t)s;
        } else if (tid < 65 + 64) {     // row 64
            int lx = tid - 65;
            int s = yk_src(S, c, X0 + lx, Y0 + 64);
            bad |= (unsigned)s;
            pix[c][64 * YK_RS + lx] = (uint8_t)s;
        }
    }
}

// one 4-pixel quad of one channel: U = S + R*N - cur*N for the four pixels, folded into a running min/max.
// |cur - S/N| <= R  <=>  0 <= U < (2R+1)N;   |cur - (S+N/2-1)/N| <= R  <=>  -(N/2-1) <= U < (2R+1)N-(N/2-1)
// (S = bilinear numerator with integer weights; identical to ((bT*tF+bB*bF)[+2^19-1])>>20 of EC.cpp:3937-3965).
template <int N>
static __device__ __forceinline__ void yk_quad(const uint8_t* __restrict__ pixc, int off, int dx0, int dy,
                                               int A3, int B, int C, int D, int& umin, int& umax) {
    unsigned word = *reinterpret_cast<const unsigned*>(pixc + off);
    int step = B + D * dy;
    int s = A3 + B * dx0 + dy * (C + D * dx0);
    int u0 = s - (int)(word & 255u) * N;
    int u1 = s + step - (int)((word >> 8) & 255u) * N;
    int u2 = s + 2 * step - (int)((word >> 16) & 255u) * N;
    int u3 = s + 3 * step - (int)(word >> 24) * N;
    umin = __vimin3_s32(umin, u0, u1); umin = __vimin3_s32(umin, u2, u3);
    umax = __vimax3_s32(umax, u0, u1); umax = __vimax3_s32(umax, u2, u3);
}

template <int FAM> static __device__ __forceinline__ int yk_family(int v) {
    return FAM == 0 ? v : (FAM == 1 ? yk_round6(v) : yk_round6p(v));
}

// Cheap rejection of all 41 tiles (1 + 2 + 2 + 4 + 8 + 8 + 16 over the seven shapes) of a 16x16 macro tile before the
// cascade: lane = tile, one channel, the quad at the tile centre, raw corners.  A pixel whose raw-family U is outside
// [loWide, hiWide) cannot be accepted by any of the six variants, so a cleared bit is a proven rejection; a set bit only
// means "run the real test".  Bit (start(pid) + t) of the result belongs to tile t of pass id pid.
// table entry of tile ti (0..40): offX | offY << 4 | shx << 8 | shy << 11 | cell << 14   (offsets in pixels inside the macro tile)
static __device__ __forceinline__ uint32_t yk_pretest_entry(int ti) {
    const int pid = (ti >= 1) + (ti >= 3) + (ti >= 5) + (ti >= 9) + (ti >= 17) + (ti >= 25);
    const int t = ti - (int)((0x19110905030100ull >> (8 * pid)) & 255ull);
    const int shx = (0x2233344 >> (4 * pid)) & 15, shy = (0x2323434 >> (4 * pid)) & 15;
    const int tx = t & ((16 >> shx) - 1), ty = t >> (4 - shx);
    const int offX = tx << shx, offY = ty << shy;
    return (uint32_t)(offX | (offY << 4) | (shx << 8) | (shy << 11) | (((offY >> 2) * 4 + (offX >> 2)) << 14));
}

static __device__ __forceinline__ unsigned long long yk_pretest(const uint8_t (*pix)[65 * YK_RS], const uint32_t* sTab, int mlx, int mly,
                                                                int X0, int Y0, int w, int h, int R, unsigned claimed) {
    const int lane = threadIdx.x & 31;
    unsigned long long P = 0;
#pragma unroll
    for (int round = 0; round < 2; round++) {
        const int ti = lane + 32 * round;
        bool possible = false;
        if (ti < 41) {
            const uint32_t e = sTab[ti];
            const int shx = (e >> 8) & 7, shy = (e >> 11) & 7, sh = shx + shy;
            const int lx0 = mlx + (e & 15), ly0 = mly + ((e >> 4) & 15), TW = 1 << shx, TH = 1 << shy, N = 1 << sh;
            if (!((claimed >> (e >> 14)) & 1u) && X0 + lx0 + TW <= w && Y0 + ly0 + TH <= h) {
                const uint8_t* p = pix[0] + ly0 * YK_RS + lx0;
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

// One FittingQuadSmooth pass over one 16x16 macro tile, by one warp.  Every tile shape of the cascade nests inside an
// aligned 16x16 macro tile, and eligibility (EC.cpp:3871-3875) only looks at cells of the same macro tile, so the whole
// 7-pass cascade of a macro tile is independent of every other macro tile: no block barrier between passes.
// The 32 lanes split into 32/NTM groups of G lanes, one group per tile of this shape; a lane evaluates 8 pixels
// (two 4-pixel quads) per family and the group votes after every quad.  `claimed` (16 bits, bit = 4*cellY + cellX) is
// warp-uniform and returned updated.
template <int SHX, int SHY, int BW, int BH>
static __device__ __forceinline__ unsigned yk_macro_pass(const uint8_t (*pix)[65 * YK_RS], uint32_t* sBits, int* sStat, uint32_t* sTouch,
                                                         int rp, int mlx, int mly, int X0, int Y0, int w, int h, int yOrg, int R, unsigned claimed, unsigned poss) {
    constexpr int TW = 1 << SHX, TH = 1 << SHY, N = TW * TH;
    constexpr int NXM = 16 / TW, NTM = NXM * (16 / TH), G = 32 / NTM, QR = TW / 4, BITS = (BW / TW) * (BH / TH);
    const int lane = threadIdx.x & 31;
    const int t = lane / G, j = lane % G;
    const int tx = t % NXM, ty = t / NXM;
    const int lx0 = mlx + tx * TW, ly0 = mly + ty * TH;
    const int cell = (ty * (TH / 4)) * 4 + tx * (TW / 4);
    // eligible (EC.cpp:3818, 3826, 3871-3875; the pre-test already checked the image bounds) and not yet proven hopeless
    const bool active = ((poss >> t) & 1u) && !((claimed >> cell) & 1u);
    if (!__any_sync(YK_FULL, active)) return claimed;

    const int hiT = (2 * R + 1) * N;                    // |cur - S/N| <= R            <=>  0 <= U < hiT
    const int loR = -(N / 2 - 1);                       // |cur - (S+N/2-1)/N| <= R    <=>  loR <= U < hiT + loR
    // a family's corners differ from the raw ones by -3..+4 (Round6: -3..+3, Round6P: -2..+4), so no variant can accept a
    // pixel whose raw-family U is outside [loWide, hiWide): one failed quad of family 0 can reject the tile for good
    const int loWide = -(4 * N + N / 2 - 1), hiWide = hiT + 3 * N;
    const unsigned gmask = (G == 32) ? YK_FULL : (((1u << (G & 31)) - 1u) << (t * G));

    // this lane's two quads
    const int q0 = j, q1 = j + G;
    const int dxA = 4 * (q0 % QR), dyA = q0 / QR, dxB = 4 * (q1 % QR), dyB = q1 / QR;
    const int offA = (ly0 + dyA) * YK_RS + lx0 + dxA, offB = (ly0 + dyB) * YK_RS + lx0 + dxB;

    int cr[3][4];                                       // TL TR BL BR, clamped at the image edge by the staging (EC.cpp:3845-3868)
    bool resolved = !active;
    int umin = INT_MAX, umax = INT_MIN;
    {
        // cheap rejection first: one quad of one channel with the raw corners proves most non-gradient tiles hopeless
        const uint8_t* p = pix[0];
        cr[0][0] = p[ly0 * YK_RS + lx0]; cr[0][1] = p[ly0 * YK_RS + lx0 + TW];
        cr[0][2] = p[(ly0 + TH) * YK_RS + lx0]; cr[0][3] = p[(ly0 + TH) * YK_RS + lx0 + TW];
        if (active)
            yk_quad<N>(p, offA, dxA, dyA, cr[0][0] * N + R * N, TH * (cr[0][1] - cr[0][0]), TW * (cr[0][2] - cr[0][0]),
                       cr[0][0] - cr[0][1] - cr[0][2] + cr[0][3], umin, umax);
        const unsigned bH = __ballot_sync(YK_FULL, (umin < loWide) || (umax >= hiWide));
        if (bH & gmask) resolved = true;
        if (!__any_sync(YK_FULL, !resolved)) return claimed;
    }
#pragma unroll
    for (int c = 1; c < 3; c++) {
        const uint8_t* p = pix[c];
        cr[c][0] = p[ly0 * YK_RS + lx0]; cr[c][1] = p[ly0 * YK_RS + lx0 + TW];
        cr[c][2] = p[(ly0 + TH) * YK_RS + lx0]; cr[c][3] = p[(ly0 + TH) * YK_RS + lx0 + TW];
    }
    bool accepted = false;
#pragma unroll
    for (int fam = 0; fam < 3; fam++) {
        if (fam > 0) { umin = INT_MAX; umax = INT_MIN; }
        if (!resolved) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                int tl, tr, bl, br;
                if (fam == 0) { tl = cr[c][0]; tr = cr[c][1]; bl = cr[c][2]; br = cr[c][3]; }
                else if (fam == 1) { tl = yk_round6(cr[c][0]); tr = yk_round6(cr[c][1]); bl = yk_round6(cr[c][2]); br = yk_round6(cr[c][3]); }
                else { tl = yk_round6p(cr[c][0]); tr = yk_round6p(cr[c][1]); bl = yk_round6p(cr[c][2]); br = yk_round6p(cr[c][3]); }
                const int A3 = tl * N + R * N, B = TH * (tr - tl), C = TW * (bl - tl), D = tl - tr - bl + br;
                if (!(fam == 0 && c == 0)) yk_quad<N>(pix[c], offA, dxA, dyA, A3, B, C, D, umin, umax);     // family 0 / channel 0 / quad A is already in
                yk_quad<N>(pix[c], offB, dxB, dyB, A3, B, C, D, umin, umax);
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
        if (fam < 2 && !__any_sync(YK_FULL, !resolved)) break;
    }
    const unsigned accB = __ballot_sync(YK_FULL, accepted && j == 0);
    if (accB == 0u) return claimed;
    if (accepted && j == 0) {
        const int sub = (ly0 / BH) * (64 / BW) + (lx0 / BW);
        const int li = sub * BITS + ((ly0 % BH) / TH) * (BW / TW) + (lx0 % BW) / TW;
        atomicOr(&sBits[li >> 5], 1u << (li & 31));                              // EC.cpp:4026
        atomicAdd(&sStat[YK_ST_TILEDONE], 1);                                    // EC.cpp:4039-4044 (mins stored as extent - value)
        atomicMax(&sStat[YK_ST_MINX], w - (X0 + lx0));
        atomicMax(&sStat[YK_ST_MINY], INT_MAX / 2 - (yOrg + Y0 + ly0));
        atomicMax(&sStat[YK_ST_MAXX], X0 + lx0 + TW);
        atomicMax(&sStat[YK_ST_MAXY], yOrg + Y0 + ly0 + TH);
        // the four lattice points this tile touches, with its role at each (mappedRGB claim, EC.cpp:4001-4021)
        const int i0 = lx0 >> 2, j0 = ly0 >> 2;
        atomicOr(&sTouch[j0 * 17 + i0], 1u << (4 * rp + 0));
        atomicOr(&sTouch[j0 * 17 + i0 + TW / 4], 1u << (4 * rp + 1));
        atomicOr(&sTouch[(j0 + TH / 4) * 17 + i0], 1u << (4 * rp + 2));
        atomicOr(&sTouch[(j0 + TH / 4) * 17 + i0 + TW / 4], 1u << (4 * rp + 3));
    }
    // EC.cpp:4029-4037: mark the accepted tiles' cells (uniformly, from the ballot)
    unsigned b = accB;
    while (b) {
        const int l = __ffs((int)b) - 1; b &= b - 1u;
        const int t2 = l / G, tx2 = t2 % NXM, ty2 = t2 / NXM;
        const unsigned cols = ((1u << (TW / 4)) - 1u) << (tx2 * (TW / 4));
#pragma unroll
        for (int r = 0; r < TH / 4; r++) claimed |= cols << (4 * (ty2 * (TH / 4) + r));
    }
    return claimed;
}

__global__ void __launch_bounds__(YK_THREADS, 4)
yk_k_analyze(const YkSlotDev* __restrict__ slots, int slot0, YkRun run) {
    __shared__ __align__(16) uint8_t pix[3][65 * YK_RS];
    __shared__ uint32_t sCell[16];
    __shared__ uint32_t sBits[YK_NPASS][8];
    __shared__ int sStat[YK_NPASS][YK_ST_STRIDE];
    __shared__ uint32_t sTouch[17 * 17];
    __shared__ uint32_t sAlpha;
    __shared__ int sNext;
    __shared__ uint32_t sTab[41];

    const YkSlotDev& S = slots[slot0 + blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31;
    const int w = S.w, h = S.h, nbx = S.nbx;
    const int bx = blockIdx.x % nbx, by = blockIdx.x / nbx;
    const int X0 = bx * YK_REGION, Y0 = by * YK_REGION;

    if (tid < 16) {
        // claimed 4x4 cells of the region; cells outside the image count as claimed
        int cy = (Y0 >> 2) + tid;
        uint32_t v = 0xFFFFu;
        if (cy * 4 < h) {
            v = S.cellMask[(size_t)cy * nbx + bx];
            int cellsIn = (w - X0) >> 2;
            if (cellsIn < 16) v |= (0xFFFFu << cellsIn) & 0xFFFFu;
        }
        sCell[tid] = v;
    }
    for (int i = tid; i < YK_NPASS * 8; i += YK_THREADS) (&sBits[0][0])[i] = 0;
    for (int i = tid; i < YK_NPASS * YK_ST_STRIDE; i += YK_THREADS) (&sStat[0][0])[i] = 0;
    for (int i = tid; i < 17 * 17; i += YK_THREADS) sTouch[i] = 0;
    if (tid == 0) { sAlpha = 0; sNext = 0; }
    if (tid >= 64 && tid < 64 + 41) sTab[tid - 64] = yk_pretest_entry(tid - 64);

    // ---- alpha plane first (its loads stay in flight while the colour planes are staged)
    const bool doAlpha = run.doAlpha && S.nPlanes == 4;
    int anz[4] = { 0, 0, 0, 0 };
    if (doAlpha) {
        const int32_t* __restrict__ P = S.plane[3];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int ly = (tid >> 4) + 16 * k, lx = (tid & 15) * 4;
            int y = Y0 + ly, x = X0 + lx;
            if (y < h && x + 3 < w) {
                int4 v = __ldg(reinterpret_cast<const int4*>(P + (size_t)y * w + x));
                anz[k] = v.x | v.y | v.z | v.w;
            } else if (y < h) {
                for (int i = 0; i < 4; i++) if (x + i < w) anz[k] |= __ldg(P + (size_t)y * w + x + i);
            }
        }
    }
    unsigned bad = 0;
    yk_stage_pixels(S, X0, Y0, pix, bad);
    __syncthreads();        // also orders the sAlpha/sBits/... initialisation

    // ---- alpha-zero tile rejection: all(alpha == 0) per 16x16 tile (EC.cpp:357-430 restated per tile) by ballot
    if (doAlpha) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned b = __ballot_sync(YK_FULL, anz[k] != 0);
            if (lane == 0 && b) {
                unsigned m = 0;
#pragma unroll
                for (int tx = 0; tx < 4; tx++) if (b & (0x000F000Fu << (4 * tx))) m |= 1u << (k * 4 + tx);
                atomicOr(&sAlpha, m);
            }
        }
    }
    if (bad & ~255u) atomicOr(&S.hdr[YK_HD_ERR], 1);

    // ---- the cascade: each warp owns two 16x16 macro tiles and runs all passes on them without block barriers
    const int R = run.rejectFactor;
    for (;;) {
        int m = 0;
        if (lane == 0) m = atomicAdd(&sNext, 1);        // macro tiles are handed out dynamically: their cost varies a lot
        m = __shfl_sync(YK_FULL, m, 0);
        if (m >= 16) break;
        const int mx = m & 3, my = m >> 2;
        unsigned claimed = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) claimed |= ((sCell[my * 4 + r] >> (4 * mx)) & 15u) << (4 * r);
        const unsigned claimed0 = claimed;
        // the 16x16 pass runs straight away (most macro tiles of illustration-like content end there); the other
        // shapes are pre-tested together, once, the first time one of them comes up
        const bool in16 = (X0 + 16 * mx + 16 <= w) && (Y0 + 16 * my + 16 <= h);
        unsigned long long P = in16 ? 1ull : 0ull;
        bool pretested = false;
        for (int rp = 0; rp < run.nPasses && claimed != 0xFFFFu; rp++) {
            const int pid = run.passId[rp];
            if (pid != 0 && !pretested) { P = yk_pretest(pix, sTab, 16 * mx, 16 * my, X0, Y0, w, h, R, claimed); pretested = true; }
            const unsigned poss = (unsigned)(P >> ((0x19110905030100ull >> (8 * pid)) & 255ull)) & 0xFFFFu;     // tiles of this shape
            switch (pid) {      // Convert()'s order, EC.cpp:9057-9093
            case 0: if (poss & 0x1u)    claimed = yk_macro_pass<4, 4, 64, 64>(pix, sBits[0], sStat[0], sTouch, rp, 16 * mx, 16 * my, X0, Y0, w, h, S.y0, R, claimed, poss); break;
            case 1: if (poss & 0x3u)    claimed = yk_macro_pass<4, 3, 64, 64>(pix, sBits[1], sStat[1], sTouch, rp, 16 * mx, 16 * my, X0, Y0, w, h, S.y0, R, claimed, poss); break;
            case 2: if (poss & 0x3u)    claimed = yk_macro_pass<3, 4, 64, 64>(pix, sBits[2], sStat[2], sTouch, rp, 16 * mx, 16 * my, X0, Y0, w, h, S.y0, R, claimed, poss); break;
            case 3: if (poss & 0xFu)    claimed = yk_macro_pass<3, 3, 64, 64>(pix, sBits[3], sStat[3], sTouch, rp, 16 * mx, 16 * my, X0, Y0, w, h, S.y0, R, claimed, poss); break;
            case 4: if (poss & 0xFFu)   claimed = yk_macro_pass<3, 2, 64, 32>(pix, sBits[4], sStat[4], sTouch, rp, 16 * mx, 16 * my, X0, Y0, w, h, S.y0, R, claimed, poss); break;
            case 5: if (poss & 0xFFu)   claimed = yk_macro_pass<2, 3, 32, 64>(pix, sBits[5], sStat[5], sTouch, rp, 16 * mx, 16 * my, X0, Y0, w, h, S.y0, R, claimed, poss); break;
            default: if (poss & 0xFFFFu) claimed = yk_macro_pass<2, 2, 32, 32>(pix, sBits[6], sStat[6], sTouch, rp, 16 * mx, 16 * my, X0, Y0, w, h, S.y0, R, claimed, poss); break;
            }
        }
        if (lane < 4 && claimed != claimed0) atomicOr(&sCell[my * 4 + lane], ((claimed >> (4 * lane)) & 15u) << (4 * mx));
    }
    __syncthreads();

    // ---- results of the region
    // accept bitmaps in the reference's swizzled layout: 16-bit units of each sub-block
    for (int i = tid; i < run.nPasses * 16; i += YK_THREADS) {
        const int pid = run.passId[i >> 4], u = i & 15;
        const YkGeomC g = yk_geom(pid);
        const int nsub = (64 / g.bw) * (64 / g.bh);
        if (u * 16 < nsub * g.bits) {
            const int sub = (u * 16) / g.bits, within = (u * 16) % g.bits;
            const int sx = X0 + (sub % (64 / g.bw)) * g.bw, sy = Y0 + (sub / (64 / g.bw)) * g.bh;
            if (sx < w && sy < h) {
                const int nSwzX = (w + g.bw - 1) / g.bw;
                const int gb = (sy / g.bh) * nSwzX + sx / g.bw;
                const uint32_t v = (sBits[pid][(u * 16) >> 5] >> ((u * 16) & 31)) & 0xFFFFu;
                reinterpret_cast<uint16_t*>(S.bitmap[pid])[((size_t)gb * g.bits + within) >> 4] = (uint16_t)v;
            }
        }
    }
    if (tid < 16) {
        int cy = (Y0 >> 2) + tid;
        if (cy * 4 < h) S.cellMask[(size_t)cy * nbx + bx] = (uint16_t)sCell[tid];
    }
    // corner colours at every 4-pixel lattice point of the region (what an accepted tile would emit, EC.cpp:4115-4132),
    // and the touch words of the lattice points (interior points are exclusive to the region, border points are shared)
    {
        const int iMax = (bx == nbx - 1) ? 17 : 16, jMax = (by == S.nby - 1) ? 17 : 16;
        for (int idx = tid; idx < 17 * 17; idx += YK_THREADS) {
            const int i = idx % 17, jj = idx / 17;
            const int gx = (X0 >> 2) + i, gy = (Y0 >> 2) + jj;
            if (gx < S.latW && gy < S.latH) {
                if (i < iMax && jj < jMax) {
                    uint8_t* d = S.latRGB + ((size_t)gy * S.latW + gx) * 3;
#pragma unroll
                    for (int c = 0; c < 3; c++) d[c] = (uint8_t)yk_compress250(yk_round6(pix[c][(4 * jj) * YK_RS + 4 * i]));
                }
                const uint32_t tv = sTouch[idx];
                if (tv) atomicOr(&S.touchMap[(size_t)gy * S.latW + gx], tv);
            }
        }
    }
    if (doAlpha && tid < 32) {
        const int tx = tid & 3, ty = (tid >> 2) & 3;
        const int px = X0 + 16 * tx, py = Y0 + 16 * ty;
        const bool in = tid < 16 && px < w && py < h;
        const bool kept = in && ((sAlpha >> tid) & 1u);
        if (in) S.alphaKept[(size_t)(py >> 4) * ((w + 15) >> 4) + (px >> 4)] = kept ? 1 : 0;
        // bounding box of kept tiles (EC.cpp:416-422), mins stored as extent - value so the header can be memset to 0
        const int big = INT_MAX / 2;
        int mnx = __reduce_max_sync(YK_FULL, kept ? w - px : 0);
        int mny = __reduce_max_sync(YK_FULL, kept ? big - (S.y0 + py) : 0);
        int mxx = __reduce_max_sync(YK_FULL, kept ? min(px + 16, w) : 0);
        int mxy = __reduce_max_sync(YK_FULL, kept ? S.y0 + min(py + 16, h) : 0);
        int cnt = __popc(__ballot_sync(YK_FULL, kept));
        if (tid == 0 && cnt) {
            atomicMax(&S.hdr[YK_HD_ALPHA_MINX], mnx); atomicMax(&S.hdr[YK_HD_ALPHA_MINY], mny);
            atomicMax(&S.hdr[YK_HD_ALPHA_MAXX], mxx); atomicMax(&S.hdr[YK_HD_ALPHA_MAXY], mxy);
            atomicAdd(&S.hdr[YK_HD_ALPHA_KEPT], cnt);
        }
    }
    if (tid >= 64 && tid < 64 + YK_NPASS) {
        const int pid = tid - 64;
        if (sStat[pid][YK_ST_TILEDONE] > 0) {
            int* d = S.hdr + YK_HD_PASS0 + pid * YK_ST_STRIDE;
            atomicAdd(&d[YK_ST_TILEDONE], sStat[pid][YK_ST_TILEDONE]);
            atomicMax(&d[YK_ST_MINX], sStat[pid][YK_ST_MINX]); atomicMax(&d[YK_ST_MINY], sStat[pid][YK_ST_MINY]);
            atomicMax(&d[YK_ST_MAXX], sStat[pid][YK_ST_MAXX]); atomicMax(&d[YK_ST_MAXY], sStat[pid][YK_ST_MAXY]);
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
// volatile access + decoupled look-back.  Work units are handed out by an atomic ticket in stream order, so a unit only
// ever waits for units with smaller tickets, which are already running or finished.
#ifdef YK_EMULATE
static inline unsigned yk_ldv(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static inline void yk_stv(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long yk_ldv64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static inline void yk_stv64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
static inline void yk_spin() { std::this_thread::yield(); }
#else
static __device__ __forceinline__ unsigned yk_ldv(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }
static __device__ __forceinline__ void yk_stv(unsigned* p, unsigned v) { *reinterpret_cast<volatile unsigned*>(p) = v; }
static __device__ __forceinline__ unsigned long long yk_ldv64(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }
static __device__ __forceinline__ void yk_stv64(unsigned long long* p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long*>(p) = v; }
static __device__ __forceinline__ void yk_spin() { __nanosleep(20); }
#endif

// status word: value << 2 | flag (1 = this unit's own total, 2 = inclusive prefix).  Returns the exclusive prefix of unit u
// and publishes its inclusive prefix.  All 32 lanes call it with the same arguments.
static __device__ unsigned yk_lookback32(uint32_t* status, int u, unsigned total) {
    const int lane = threadIdx.x & 31;
    if (u == 0) { if (lane == 0) yk_stv(&status[0], (total << 2) | 2u); return 0u; }
    if (lane == 0) yk_stv(&status[u], (total << 2) | 1u);
    unsigned base = 0;
    int look = u - 1;
    while (look >= 0) {
        const int idx = look - lane;
        const unsigned st = idx >= 0 ? yk_ldv(&status[idx]) : 2u;       // before the first unit: inclusive prefix 0
        const unsigned ready = __ballot_sync(YK_FULL, (st & 3u) != 0u);
        const unsigned incl = __ballot_sync(YK_FULL, (st & 3u) == 2u);
        const unsigned need = incl ? ((2u << (__ffs((int)incl) - 1)) - 1u) : YK_FULL;     // lanes up to the nearest inclusive prefix
        if ((ready & need) != need) { yk_spin(); continue; }
        base += __reduce_add_sync(YK_FULL, ((need >> lane) & 1u) ? (st >> 2) : 0u);
        if (incl) break;
        look -= 32;
    }
    if (lane == 0) yk_stv(&status[u], ((base + total) << 2) | 2u);
    return base;
}

// same with two counters packed in 64 bits: hi << 32 | lo << 2 | flag
static __device__ unsigned long long yk_lookback64(unsigned long long* status, int u, unsigned hi, unsigned lo) {
    const int lane = threadIdx.x & 31;
    const unsigned long long mine = ((unsigned long long)hi << 32) | ((unsigned long long)lo << 2);
    if (u == 0) { if (lane == 0) yk_stv64(&status[0], mine | 2ull); return 0ull; }
    if (lane == 0) yk_stv64(&status[u], mine | 1ull);
    unsigned bhi = 0, blo = 0;
    int look = u - 1;
    while (look >= 0) {
        const int idx = look - lane;
        const unsigned long long st = idx >= 0 ? yk_ldv64(&status[idx]) : 2ull;
        const unsigned fl = (unsigned)(st & 3ull);
        const unsigned ready = __ballot_sync(YK_FULL, fl != 0u);
        const unsigned incl = __ballot_sync(YK_FULL, fl == 2u);
        const unsigned need = incl ? ((2u << (__ffs((int)incl) - 1)) - 1u) : YK_FULL;
        if ((ready & need) != need) { yk_spin(); continue; }
        const bool use = (need >> lane) & 1u;
        bhi += __reduce_add_sync(YK_FULL, use ? (unsigned)(st >> 32) : 0u);
        blo += __reduce_add_sync(YK_FULL, use ? (unsigned)((st & 0xFFFFFFFFull) >> 2) : 0u);
        if (incl) break;
        look -= 32;
    }
    if (lane == 0) yk_stv64(&status[u], (((unsigned long long)(bhi + hi)) << 32) | ((unsigned long long)(blo + lo) << 2) | 2ull);
    return ((unsigned long long)bhi << 32) | blo;
}

// ------------------------------------------------------------------------------------------------------------------
// Corner ownership and emission in one kernel.  The reference walks tiles in stream order and lets a tile emit a corner
// colour only if no earlier tile (of this or an earlier pass) touched that lattice point (mappedRGB, EC.cpp:4001-4021,
// 4115-4132).  Order-free: a lattice point is emitted in the first pass that touches it, by the accepted toucher with the
// smallest stream position — all of which the point's touch word says.  One warp per (pass, swizzle block); swizzle
// blocks are taken in stream order, the running rgb byte offset comes from a decoupled look-back.
struct YkGeomS { int shx, shy, lbw, lbh, bits; };
__constant__ YkGeomS yk_geom_s_tab[YK_NPASS] = { {4,4,6,6,16}, {4,3,6,6,32}, {3,4,6,6,32}, {3,3,6,6,64}, {3,2,6,5,64}, {2,3,5,6,64}, {2,2,5,5,64} };
static __device__ __forceinline__ YkGeomS yk_geom_s(int pid) { return yk_geom_s_tab[pid]; }
static __device__ __forceinline__ int yk_pos_s(const YkGeomS& g, int nSwzX, int gtx, int gty) {
    const int x = gtx << g.shx, y = gty << g.shy;
    return (((y >> g.lbh) * nSwzX + (x >> g.lbw)) * g.bits) + (((y & ((1 << g.lbh) - 1)) >> g.shy) << (g.lbw - g.shx)) + ((x & ((1 << g.lbw) - 1)) >> g.shx);
}

// emit mask (which of TL,TR,BL,BR the accepted tile at (gtx,gty), stream position myPos, emits) from the touch words
static __device__ __forceinline__ int yk_emit_mask(const YkSlotDev& S, const YkGeomS& g, int nSwzX, int rp, int gtx, int gty, int myPos) {
    int m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int LX = gtx + (k & 1), LY = gty + (k >> 1);                       // lattice point in tile units
        const uint32_t word = __ldg(&S.touchMap[(size_t)((LY << g.shy) >> 2) * S.latW + ((LX << g.shx) >> 2)]);
        if (word & 0x80000000u) continue;                                        // claimed by an earlier launch
        if (((__ffs((int)(word & 0x0FFFFFFFu)) - 1) >> 2) != rp) continue;       // an earlier pass of this launch got it
        const unsigned nib = (word >> (4 * rp)) & 15u;                           // roles present in this pass
        bool owner = true;
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++)
            if (k2 != k && ((nib >> k2) & 1u) && yk_pos_s(g, nSwzX, LX - (k2 & 1), LY - (k2 >> 1)) < myPos) owner = false;
        if (owner) m |= 1 << k;
    }
    return m;
}

// One warp per swizzle block (lane = up to 2 tiles), one 1024-thread CTA per group of 32 consecutive blocks of a pass.
// Groups are taken in stream order (ticket) and chained by one look-back per CTA, so the chain is nUnits/32 long.
// Tickets past the gradient groups scan the DynamicTileCompressor segments (r2Off), 1024 segments per CTA.
#define YK_EMIT_THREADS 1024
__global__ void __launch_bounds__(YK_EMIT_THREADS)
yk_k_emit(const YkSlotDev* __restrict__ slots, int slot0, YkRun run, int gradGroups, int r2Groups) {
    __shared__ int sTicket;
    __shared__ unsigned sA[33], sB[33];
    const YkSlotDev& S = slots[slot0 + blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = S.w, h = S.h;
    if (tid == 0) sTicket = atomicAdd(&S.hdr[YK_HD_TICKET_EMIT], 1);
    __syncthreads();
    const int ticket = sTicket;
    if (ticket >= gradGroups + r2Groups) return;
    if (ticket >= gradGroups) {
        // ---- offsets of DynamicTileCompressor's 8-tile segments in its row-major tile order (EC.cpp:8412-8413)
        const int grp = ticket - gradGroups, nbx = S.nbx, nSegs = (h >> 3) * nbx;
        const int seg = grp * YK_EMIT_THREADS + tid;
        unsigned chunks = 0, tiles = 0;
        if (seg < nSegs) {
            const int bx = seg % nbx, ty = seg / nbx;
            uint32_t r0 = S.cellMask[(size_t)(2 * ty) * nbx + bx], r1 = S.cellMask[(size_t)(2 * ty + 1) * nbx + bx];
            const int cellsIn = (w - bx * 64) >> 2;
            if (cellsIn < 16) { r0 |= (0xFFFFu << cellsIn) & 0xFFFFu; r1 |= (0xFFFFu << cellsIn) & 0xFFFFu; }
#pragma unroll
            for (int x = 0; x < 8; x++) {
                const int n = 4 - __popc(((r0 >> (2 * x)) & 3u) | (((r1 >> (2 * x)) & 3u) << 2));
                chunks += n; tiles += (n > 0);
            }
        }
        unsigned ic = chunks, it = tiles;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned t0 = __shfl_up_sync(YK_FULL, ic, d), t1 = __shfl_up_sync(YK_FULL, it, d);
            if (lane >= d) { ic += t0; it += t1; }
        }
        if (lane == 31) { sA[warp] = ic; sB[warp] = it; }
        __syncthreads();
        if (warp == 0) {
            unsigned a = sA[lane], b2 = sB[lane], ia = a, ib = b2;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned t0 = __shfl_up_sync(YK_FULL, ia, d), t1 = __shfl_up_sync(YK_FULL, ib, d);
                if (lane >= d) { ia += t0; ib += t1; }
            }
            const unsigned totC = __shfl_sync(YK_FULL, ia, 31), totT = __shfl_sync(YK_FULL, ib, 31);
            const unsigned long long base = yk_lookback64(S.r2Status, grp, totC, totT);
            const unsigned bc = (unsigned)(base >> 32), bt = (unsigned)(base & 0xFFFFFFFFull);
            sA[lane] = bc + ia - a; sB[lane] = bt + ib - b2;
            if (grp == r2Groups - 1 && lane == 0) { S.hdr[YK_HD_R2_CHUNKS] = (int)(bc + totC); S.hdr[YK_HD_R2_TILES] = (int)(bt + totT); }
        }
        __syncthreads();
        if (seg < nSegs) S.r2Off[seg] = make_uint2(sA[warp] + ic - chunks, sB[warp] + it - tiles);
        return;
    }
    // ticket -> (pass position, group of 32 swizzle blocks)
    int rp = 0, grp = ticket, nUnits = 0, nSwzX = 0, nGroups = 0;
    YkGeomS g = yk_geom_s(run.passId[0]);
    for (;;) {
        g = yk_geom_s(run.passId[rp]);
        nSwzX = (w + (1 << g.lbw) - 1) >> g.lbw;
        nUnits = nSwzX * ((h + (1 << g.lbh) - 1) >> g.lbh);
        nGroups = (nUnits + 31) >> 5;
        if (grp < nGroups) break;
        grp -= nGroups; rp++;
    }
    const int pid = run.passId[rp];
    const int u = grp * 32 + warp;
    const int tprShift = g.lbw - g.shx;                                     // log2(tiles per row of the swizzle block)
    unsigned long long acc = 0;
    int tx0 = 0, ty0 = 0;
    if (u < nUnits) {
        if (g.bits == 16) acc = __ldg(&reinterpret_cast<const uint16_t*>(S.bitmap[pid])[u]);
        else if (g.bits == 32) acc = __ldg(&reinterpret_cast<const uint32_t*>(S.bitmap[pid])[u]);
        else { uint2 v = __ldg(&reinterpret_cast<const uint2*>(S.bitmap[pid])[u]); acc = v.x | ((unsigned long long)v.y << 32); }
        tx0 = ((u % nSwzX) << g.lbw) >> g.shx; ty0 = ((u / nSwzX) << g.lbh) >> g.shy;
    }
    int mask[2] = { 0, 0 };
#pragma unroll
    for (int s2 = 0; s2 < 2; s2++) {
        const int li = lane + 32 * s2;
        if (li < g.bits && ((acc >> li) & 1ull))
            mask[s2] = yk_emit_mask(S, g, nSwzX, rp, tx0 + (li & ((1 << tprShift) - 1)), ty0 + (li >> tprShift), u * g.bits + li);
    }
    // exclusive prefix of the emitted bytes inside the block (tile order = lane order, second half after the first)
    const unsigned c0 = 3u * __popc(mask[0]), c1 = 3u * __popc(mask[1]);
    unsigned i0 = c0, i1 = c1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned t0 = __shfl_up_sync(YK_FULL, i0, d), t1 = __shfl_up_sync(YK_FULL, i1, d);
        if (lane >= d) { i0 += t0; i1 += t1; }
    }
    const unsigned tot0 = __shfl_sync(YK_FULL, i0, 31), tot1 = __shfl_sync(YK_FULL, i1, 31);
    if (lane == 0) sA[warp] = tot0 + tot1;
    __syncthreads();
    if (warp == 0) {
        unsigned a = sA[lane], ia = a;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned t0 = __shfl_up_sync(YK_FULL, ia, d); if (lane >= d) ia += t0; }
        const unsigned tot = __shfl_sync(YK_FULL, ia, 31);
        const unsigned base = yk_lookback32(S.emitStatus[pid], grp, tot);
        sA[lane] = base + ia - a;
        if (grp == nGroups - 1 && lane == 0) S.hdr[YK_HD_PASS0 + pid * YK_ST_STRIDE + YK_ST_RGBBYTES] = (int)(base + tot);
    }
    __syncthreads();
    const unsigned base = sA[warp];
    uint8_t* out = S.rgb[pid];
#pragma unroll
    for (int s2 = 0; s2 < 2; s2++) {
        if (mask[s2]) {
            const int li = lane + 32 * s2;
            const int gtx = tx0 + (li & ((1 << tprShift) - 1)), gty = ty0 + (li >> tprShift);
            unsigned off = base + (s2 ? tot0 + i1 - c1 : i0 - c0);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if ((mask[s2] >> k) & 1) {                                             // TL, TR, BL, BR (EC.cpp:4115-4132)
                    const int gx = ((gtx + (k & 1)) << g.shx) >> 2, gy = ((gty + (k >> 1)) << g.shy) >> 2;
                    const uint8_t* sp = S.latRGB + ((size_t)gy * S.latW + gx) * 3;
                    out[off] = sp[0]; out[off + 1] = sp[1]; out[off + 2] = sp[2];
                    off += 3;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// DynamicTileCompressor (EC.cpp:8398-8522).  One warp per 8-tile segment (one tile row of one 64-pixel column) in the
// reference's row-major tile order; stream offsets were scanned by yk_k_emit; per tile and plane one pass with lane = 2 pixels.
__global__ void __launch_bounds__(YK_THREADS)
yk_k_range1d(const YkSlotDev* __restrict__ slots, int slot0, int nSegs) {
    __shared__ uint32_t hist[YK_THREADS / 32][3][256];
    __shared__ uint32_t sMagic[256];                        // ceil(2^20 / d): exact floor(n / d) for n < 4112, d <= 255
    const YkSlotDev& S = slots[slot0 + blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = S.w, nbx = S.nbx;
    sMagic[tid] = tid ? ((1u << 20) + (unsigned)tid - 1u) / (unsigned)tid : 0u;
    __syncthreads();
    const int ticket = blockIdx.x * (YK_THREADS / 32) + warp;
    if (ticket >= nSegs) return;
    const int bx = ticket % nbx, ty = ticket / nbx;                 // tile row ty, 64-pixel column bx
    const int X0 = bx * 64, Y0 = ty * 8;
    uint32_t r0 = S.cellMask[(size_t)(2 * ty) * nbx + bx], r1 = S.cellMask[(size_t)(2 * ty + 1) * nbx + bx];
    {
        const int cellsIn = (w - X0) >> 2;
        if (cellsIn < 16) { r0 |= (0xFFFFu << cellsIn) & 0xFFFFu; r1 |= (0xFFFFu << cellsIn) & 0xFFFFu; }
    }
    if (((r0 & r1) & 0xFFFFu) == 0xFFFFu) return;           // every cell claimed: nothing to code in this segment
    for (int i = lane; i < 3 * 256; i += 32) (&hist[warp][0][0])[i] = 0;
    const uint2 off = __ldg(&S.r2Off[ticket]);              // scanned by yk_k_emit
    int chunkOff = (int)off.x, tileOff = (int)off.y;
    __syncwarp();

    const int r = lane >> 2, c0 = (lane & 3) * 2;           // pixel row / first column of this lane inside the tile
    const int band = r >> 2, right = c0 >> 2;
    const int32_t* __restrict__ P0 = S.plane[0] + (size_t)(Y0 + r) * w + X0 + c0;
    const int32_t* __restrict__ P1 = S.plane[1] + (size_t)(Y0 + r) * w + X0 + c0;
    const int32_t* __restrict__ P2 = S.plane[2] + (size_t)(Y0 + r) * w + X0 + c0;
    for (int tx = 0; tx < 8; tx++) {
        // quadrant needs coding iff its top-left map pixel is 0 (EC.cpp:8420-8430) == its 4x4 cell is unclaimed
        const unsigned q = (~(((r0 >> (2 * tx)) & 3u) | (((r1 >> (2 * tx)) & 3u) << 2))) & 15u;   // bit0 TL, 1 TR, 2 BL, 3 BR
        if (q == 0) continue;
        const bool valid = (q >> (band * 2 + right)) & 1u;
        const unsigned qb = (q >> (band * 2)) & 3u;                     // coded quadrants of this band: bit0 left, bit1 right
        const int lengthX = (qb == 3u) ? 8 : 4, x2 = (qb == 2u) ? 4 : 0;
        const int pos = (band ? 16 * __popc(q & 3u) : 0) + (r & 3) * lengthX + (c0 - x2);
        // the three planes of the tile are coded side by side so their latencies overlap
        int vx[3] = { 0, 0, 0 }, vy[3] = { 0, 0, 0 };
        if (valid) {
            const int2 a = __ldg(reinterpret_cast<const int2*>(P0 + 8 * tx));
            const int2 b = __ldg(reinterpret_cast<const int2*>(P1 + 8 * tx));
            const int2 c = __ldg(reinterpret_cast<const int2*>(P2 + 8 * tx));
            vx[0] = a.x & 255; vy[0] = a.y & 255; vx[1] = b.x & 255; vy[1] = b.y & 255; vx[2] = c.x & 255; vy[2] = c.y & 255;   // CompressF(v,255) == v (EC.cpp:8442)
#pragma unroll
            for (int p = 0; p < 3; p++) { atomicAdd(&hist[warp][p][vx[p]], 1u); atomicAdd(&hist[warp][p][vy[p]], 1u); }
        }
        __syncwarp();
        // FindAndRemoveMostUsedColor (EC.cpp:8335-8356): highest index among the maximal counts; only present values can win
        unsigned key[3] = { 0, 0, 0 };
        if (valid) {
#pragma unroll
            for (int p = 0; p < 3; p++) key[p] = max((hist[warp][p][vx[p]] << 8) | (unsigned)vx[p], (hist[warp][p][vy[p]] << 8) | (unsigned)vy[p]);
        }
#pragma unroll
        for (int p = 0; p < 3; p++) key[p] = __reduce_max_sync(YK_FULL, key[p]);
        __syncwarp();
        if (valid) {
#pragma unroll
            for (int p = 0; p < 3; p++) { hist[warp][p][vx[p]] = 0; hist[warp][p][vy[p]] = 0; }
        }
        int color0[3], mn[3], mx[3];
        bool remx[3], remy[3];
#pragma unroll
        for (int p = 0; p < 3; p++) {
            color0[p] = min(max((int)(key[p] & 255u), 1), 254);
            // Model1 (EC.cpp:8358-8381) over what is left of the histogram
            remx[p] = valid && (vx[p] < color0[p] - 1 || vx[p] > color0[p] + 1);
            remy[p] = valid && (vy[p] < color0[p] - 1 || vy[p] > color0[p] + 1);
            mn[p] = min(remx[p] ? vx[p] : 999, remy[p] ? vy[p] : 999);
            mx[p] = max(remx[p] ? vx[p] : -1, remy[p] ? vy[p] : -1);
        }
#pragma unroll
        for (int p = 0; p < 3; p++) { mn[p] = __reduce_min_sync(YK_FULL, mn[p]); mx[p] = __reduce_max_sync(YK_FULL, mx[p]); }
#pragma unroll
        for (int p = 0; p < 3; p++) {
            int minCol = 0, delta = 0;
            if (mn[p] != 999) { minCol = mn[p]; delta = mx[p] - mn[p]; }
            if (valid) {
                // GetValueModel1 (EC.cpp:8383-8391): C division of a numerator in -1..3951 by delta in 1..255.
                // floor(n/d) == (n * ceil(2^20/d)) >> 20 for 0 <= n < 4112, d <= 255; n == -1 only happens for delta == 1.
                int bxv = 0, byv = 0;
                if (delta) {
                    const unsigned magic = sMagic[delta];
                    const int rnd = (delta >> 1) - 1;
                    if (remx[p]) { const int n = (vx[p] - minCol) * 15 + rnd; bxv = 1 + (n < 0 ? n : (int)(((unsigned)n * magic) >> 20)); }
                    if (remy[p]) { const int n = (vy[p] - minCol) * 15 + rnd; byv = 1 + (n < 0 ? n : (int)(((unsigned)n * magic) >> 20)); }
                } else { bxv = remx[p] ? 1 : 0; byv = remy[p] ? 1 : 0; }
                uint8_t* d = S.r2Idx[p] + (size_t)chunkOff * 16 + pos;
                *reinterpret_cast<uint16_t*>(d) = (uint16_t)((bxv & 255) | ((byv & 255) << 8));
            }
            if (lane == p) {
                uint8_t* t = S.r2Type[p] + (size_t)tileOff * 3;            // EC.cpp:8503-8505
                t[0] = (uint8_t)color0[p]; t[1] = (uint8_t)minCol; t[2] = (uint8_t)delta;
            }
        }
        chunkOff += __popc(q); tileOff += 1;
    }
}

static __device__ __forceinline__ int yk_accept_bit(const YkSlotDev& S, int pid, const YkGeomC& g, int gtx, int gty) {
    if (gtx < 0 || gty < 0 || ((gtx + 1) << g.shx) > S.w || ((gty + 1) << g.shy) > S.h) return 0;
    int pos = yk_tile_pos(g, S.w, gtx, gty);
    return (S.bitmap[pid][pos >> 3] >> (pos & 7)) & 1;
}

// ------------------------------------------------------------------------------------------------------------------
// Compat download: expand the compact masks into the reference's int32 state planes (EncoderContext.h:300-323).
__global__ void __launch_bounds__(YK_THREADS)
yk_k_state(const YkSlotDev* __restrict__ slots, int slot, int32_t* smoothMap, int32_t* mipmapMask, int32_t* mappedRGB,
           int32_t* recon0, int32_t* recon1, int32_t* recon2) {
    __shared__ __align__(16) uint8_t pix[3][65 * YK_RS];
    __shared__ uint32_t sCell[16];
    const YkSlotDev& S = slots[slot];
    const int tid = threadIdx.x;
    const int w = S.w, h = S.h, nbx = S.nbx;
    const int bx = blockIdx.x % nbx, by = blockIdx.x / nbx;
    const int X0 = bx * YK_REGION, Y0 = by * YK_REGION;
    if (tid < 16) {
        int cy = (Y0 >> 2) + tid;
        sCell[tid] = (cy * 4 < h) ? S.cellMask[(size_t)cy * nbx + bx] : 0u;
    }
    unsigned bad = 0;
    if (recon0) yk_stage_pixels(S, X0, Y0, pix, bad);
    __syncthreads();
    const int tw16 = (w + 15) >> 4;
    for (int i = tid; i < 64 * 64; i += YK_THREADS) {
        const int lx = i & 63, ly = i >> 6, x = X0 + lx, y = Y0 + ly;
        if (x >= w || y >= h) continue;
        const bool claimed = (sCell[ly >> 2] >> (lx >> 2)) & 1u;
        const size_t o = (size_t)y * w + x;
        if (smoothMap) smoothMap[o] = claimed ? 255 : 0;
        if (mipmapMask) {
            int mv = 255;
            if (S.alphaValid && !S.alphaReset && !S.alphaKept[(size_t)(y >> 4) * tw16 + (x >> 4)]) mv = 0;    // EC.cpp:394-414
            if (claimed) mv = 0;                                                                               // EC.cpp:4035
            mipmapMask[o] = mv;
        }
        if (recon0) { recon0[o] = 0; recon1[o] = 0; recon2[o] = 0; }
    }
    if (mappedRGB) {
        const int xMax = (bx == nbx - 1) ? 65 : 64, yMax = (by == S.nby - 1) ? 65 : 64;
        for (int i = tid; i < 65 * 65; i += YK_THREADS) {
            const int lx = i % 65, ly = i / 65, x = X0 + lx, y = Y0 + ly;
            if (lx >= xMax || ly >= yMax || x > w || y > h) continue;
            int v = 0;
            if (!(x & 3) && !(y & 3)) v = S.touchMap[(size_t)(y >> 2) * S.latW + (x >> 2)] ? 255 : 0;
            mappedRGB[(size_t)y * (w + 1) + x] = v;
        }
    }
    if (!recon0) return;
    __syncthreads();
    // recon = Round6P family, rounded variant, of every accepted tile; later passes overwrite earlier ones
    // (EC.cpp:3969-3971, 4096-4104), passes in Convert()'s order
    int32_t* rec[3] = { recon0, recon1, recon2 };
    for (int pid = 0; pid < YK_NPASS; pid++) {
        const YkGeomC g = yk_geom(pid);
        const int tw = 1 << g.shx, th = 1 << g.shy, N = tw * th, NX = 64 / tw;
        for (int i = tid; i < 64 * 64; i += YK_THREADS) {
            const int lx = i & 63, ly = i >> 6;
            const int tx = lx / tw, ty = ly / th;
            if (!yk_accept_bit(S, pid, g, bx * NX + tx, by * (64 / th) + ty)) continue;
            const int lx0 = tx * tw, ly0 = ty * th, dx = lx - lx0, dy = ly - ly0;
            for (int c = 0; c < 3; c++) {
                const uint8_t* p = pix[c];
                int tl = yk_round6p(p[ly0 * YK_RS + lx0]), tr = yk_round6p(p[ly0 * YK_RS + lx0 + tw]);
                int bl = yk_round6p(p[(ly0 + th) * YK_RS + lx0]), br = yk_round6p(p[(ly0 + th) * YK_RS + lx0 + tw]);
                int s = (tl * (tw - dx) + tr * dx) * (th - dy) + (bl * (tw - dx) + br * dx) * dy;
                rec[c][(size_t)(Y0 + ly) * w + X0 + lx] = (s + N / 2 - 1) / N;
            }
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Range stage R1: DynamicTileEncode (EC.cpp:4365-4503) = LeftRightOrder walk (framework.h:228-256) over the 8-aligned
// bound box, GetMinMax_Y (Plane.cpp:489-587) and GetTileDynamic_Y (EC.cpp:747-1212) per 8x8 block.
// Block i of the walk is (cx + 8*(i % nbw), cy + 8*(i / nbw)); the reference's size quirk
// `w = (x+8 > constraint.w) ? x%8 : 8` (compares with the width, not the right edge) empties blocks with
// x + 8 > cw or y + 8 > ch.  valid pixel = mipmapMask != 0 && smoothMap == 0 (Plane.cpp:525-527).
static __device__ __forceinline__ unsigned yk_r1_cells(const YkSlotDev& S, int x, int y) {
    // unclaimed-and-kept 4x4 cells of the 8x8 block at (x, y): bit0 TL, bit1 TR, bit2 BL, bit3 BR
    if (S.alphaValid && !S.alphaReset && !S.alphaKept[(size_t)(y >> 4) * ((S.w + 15) >> 4) + (x >> 4)]) return 0u;
    const int cx = x >> 2, cy = y >> 2;
    const unsigned r0 = S.cellMask[(size_t)cy * S.nbx + (cx >> 4)], r1 = S.cellMask[(size_t)(cy + 1) * S.nbx + (cx >> 4)];
    return (~(((r0 >> (cx & 15)) & 3u) | (((r1 >> (cx & 15)) & 3u) << 2))) & 15u;
}

// block-wide exclusive scan of one int per thread (used by the R1 offset scan)
static __device__ int yk_block_exclusive(int v, int* sWarp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(YK_FULL, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int x = lane < nw ? sWarp[lane] : 0, xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(YK_FULL, xi, d); if (lane >= d) xi += t; }
        sWarp[lane] = xi - x;
        if (lane == 31) sWarp[32] = xi;
    }
    __syncthreads();
    total = sWarp[32];
    const int r = sWarp[warp] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256)
yk_k_r1_count(const YkSlotDev* __restrict__ slots, int slot, int cx, int cy, int cw, int ch, int nBlocks) {
    const YkSlotDev& S = slots[slot];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nBlocks) return;
    const int nbw = cw >> 3;
    const int x = cx + 8 * (i % nbw), y = cy + 8 * (i / nbw);
    int n = 0;
    if (x + 8 <= cw && y + 8 <= ch && x + 8 <= S.w && y + 8 <= S.h) n = 16 * __popc(yk_r1_cells(S, x, y));
    S.r1Cnt[i] = n;
}

__global__ void __launch_bounds__(1024)
yk_k_r1_scan(const YkSlotDev* __restrict__ slots, int slot, int nBlocks, int plane) {
    __shared__ int sWarp[33];
    const YkSlotDev& S = slots[slot];
    const int tid = threadIdx.x;
    const int per = (nBlocks + (int)blockDim.x - 1) / (int)blockDim.x;
    const int b = min(nBlocks, tid * per), e = min(nBlocks, b + per);
    int sn = 0, sd = 0;
    for (int i = b; i < e; i++) { int v = S.r1Cnt[i]; sn += v; sd += (v > 0); }
    int totN, totD;
    int rn = yk_block_exclusive(sn, sWarp, totN);
    int rd = yk_block_exclusive(sd, sWarp, totD);
    for (int i = b; i < e; i++) { int v = S.r1Cnt[i]; S.r1Cnt[i] = rn | (v ? (1 << 31) : 0); S.r1Def[i] = rd; rn += v; rd += (v > 0); }
    if (tid == 0) { S.hdr[YK_HD_R1_NIB0 + plane] = totN; S.hdr[YK_HD_R1_DEF0 + plane] = totD; }
}

// lut: [64 base6][176 range7][72] ints = six LUTs (16,16,16,8,8,8 entries) of DynamicTile::buildTable (EC.cpp:625-699)
__global__ void __launch_bounds__(YK_THREADS)
yk_k_r1_encode(const YkSlotDev* __restrict__ slots, int slot, int plane, int mode3, int cx, int cy, int cw, int ch,
               int nBlocks, const int* __restrict__ lut) {
    __shared__ float sTerm[YK_THREADS / 32][6][64];
    const YkSlotDev& S = slots[slot];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * (YK_THREADS / 32) + warp;
    const int nbw = cw >> 3;
    const bool inRange = i < nBlocks;
    const int x = cx + 8 * ((inRange ? i : 0) % nbw), y = cy + 8 * ((inRange ? i : 0) / nbw);
    unsigned cells = 0;
    int offWord = inRange ? S.r1Cnt[i] : 0;
    if (inRange && (offWord < 0) ) cells = yk_r1_cells(S, x, y);          // sign bit = block has valid pixels
    if (cells == 0) return;                                               // warp-uniform
    const int nibOff = offWord & 0x7FFFFFFF, defOff = S.r1Def[i];
    const int r = lane >> 2, c0 = (lane & 3) * 2;
    const bool valid = (cells >> ((r >> 2) * 2 + (c0 >> 2))) & 1u;
    int v0 = 0, v1 = 0;
    if (valid) {
        int2 p = __ldg(reinterpret_cast<const int2*>(S.plane[plane] + (size_t)(y + r) * S.w + x + c0));
        v0 = p.x; v1 = p.y;
    }
    int mn = __reduce_min_sync(YK_FULL, valid ? min(v0, v1) : INT_MAX);
    int mx = __reduce_max_sync(YK_FULL, valid ? max(v0, v1) : INT_MIN);
    int sgn = 0;
    if (mn < 0) { mn += 128; mx += 128; sgn = 128; }                       // EC.cpp:764-768
    mn = min(max(mn, 0), 255); mx = min(max(mx, mn), 255);
    // DynamicTile::buildTable index (EC.cpp:635-650)
    const int m = min(mn, 224);
    const int diff = max(mx - m, 16);
    const int b6 = (m * 63 + 112) / 224, BN = (b6 * 224) / 63;
    const int scale = 223 - BN;
    const int r7 = ((max(diff, 32) - 32) * 127 + scale - 1) / scale;
    const int* T = lut + ((size_t)b6 * 176 + min(r7, 175)) * 72;
    const int o0 = v0 + sgn, o1 = v1 + sgn;
    unsigned codes0 = 0, codes1 = 0;                                       // 4 bits per mode
    const int startMode = mode3 ? 3 : 0;
    for (int mode = startMode; mode < 6; mode++) {
        const int count = mode < 3 ? 16 : 8;
        const int* L = T + (mode < 3 ? 16 * mode : 48 + 8 * (mode - 3));
        int d0 = 99999, d1 = 99999, f0 = 0, f1 = 0;
        for (int n = 0; n < count; n++) {                                  // first strict minimum, EC.cpp:873-881
            const int e = __ldg(L + n);
            const int a0 = abs(e - o0), a1 = abs(e - o1);
            if (a0 < d0) { d0 = a0; f0 = n; }
            if (a1 < d1) { d1 = a1; f1 = n; }
        }
        codes0 |= (unsigned)f0 << (4 * mode); codes1 |= (unsigned)f1 << (4 * mode);
        // cumulated relative error term, float32 (EC.cpp:884-886); invalid pixels add +0.0f which leaves the sum unchanged
        sTerm[warp][mode][2 * lane] = (valid && o0 != 0) ? ((float)d0 / (float)o0) : 0.0f;
        sTerm[warp][mode][2 * lane + 1] = (valid && o1 != 0) ? ((float)d1 / (float)o1) : 0.0f;
    }
    __syncwarp();
    float err = 0.0f;
    if (lane >= startMode && lane < 6) {
        // the reference sums in row-major valid-pixel order; float addition is not associative, so one lane per mode
        for (int k = 0; k < 64; k++) err = __fadd_rn(err, sTerm[warp][lane][k]);
    }
    int bestMode = -1;
    float bestErr = 99999999.0f;
    for (int mode = startMode; mode < 6; mode++) {                          // `<=`: later modes win ties, EC.cpp:897-905
        const float e = __shfl_sync(YK_FULL, err, mode);
        if (e <= bestErr) { bestErr = e; bestMode = mode; }
    }
    const unsigned bv0 = __ballot_sync(YK_FULL, valid);
    if (valid) {
        const int before = 2 * __popc(bv0 & ((1u << lane) - 1u));           // both pixels of a lane share validity
        const int c0v = (codes0 >> (4 * bestMode)) & 15, c1v = (codes1 >> (4 * bestMode)) & 15;
        const int n0 = nibOff + before;                                     // nibble index of pixel 0; pixel 1 follows
        uint32_t* W = S.r1Nib[plane];
        atomicOr(&W[n0 >> 3], (uint32_t)c0v << (4 * (n0 & 7)));             // low nibble first, EC.cpp:1180-1184
        atomicOr(&W[(n0 + 1) >> 3], (uint32_t)c1v << (4 * ((n0 + 1) & 7)));
        if (S.r1Dst) {
            const int* L = T + (bestMode < 3 ? 16 * bestMode : 48 + 8 * (bestMode - 3));
            int* d = S.r1Dst + (size_t)(y + r) * S.w + x + c0;
            d[0] = __ldg(L + c0v); d[1] = __ldg(L + c1v);                    // EC.cpp:4448-4457 (offset 0 for full-resolution planes)
        }
    }
    if (lane == 0) S.r1Defs[plane][defOff] = (uint16_t)((bestMode << 13) | (r7 << 7) | b6);    // EncodeTileType, YAIK_private.h:358
}

// ------------------------------------------------------------------------------------------------------------------
// launch wrappers
void yk_launch_analyze(const YkSlotDev* slotsDev, int slot0, int nSlots, int nRegions, const YkRun& run, cudaStream_t st) {
    YK_LAUNCH(yk_k_analyze, dim3(nRegions, nSlots), dim3(YK_THREADS), 0, st, slotsDev, slot0, run);
}
void yk_launch_fold_touch(const YkSlotDev* slotsDev, int slot0, int nSlots, int nWords, cudaStream_t st) {
    YK_LAUNCH(yk_k_fold_touch, dim3((nWords + 255) / 256, nSlots), dim3(256), 0, st, slotsDev, slot0, nWords);
}
void yk_launch_emit(const YkSlotDev* slotsDev, int slot0, int nSlots, int gradGroups, int r2Groups, const YkRun& run, cudaStream_t st) {
    YK_LAUNCH(yk_k_emit, dim3(gradGroups + r2Groups, nSlots), dim3(YK_EMIT_THREADS), 0, st, slotsDev, slot0, run, gradGroups, r2Groups);
}
void yk_launch_range1d(const YkSlotDev* slotsDev, int slot0, int nSlots, int nSegs, cudaStream_t st) {
    const int per = YK_THREADS / 32;
    YK_LAUNCH(yk_k_range1d, dim3((nSegs + per - 1) / per, nSlots), dim3(YK_THREADS), 0, st, slotsDev, slot0, nSegs);
}
void yk_launch_state(const YkSlotDev* slotsDev, int slot, int nRegions, int32_t* smoothMap, int32_t* mipmapMask,
                     int32_t* mappedRGB, int32_t* recon0, int32_t* recon1, int32_t* recon2, cudaStream_t st) {
    YK_LAUNCH(yk_k_state, dim3(nRegions), dim3(YK_THREADS), 0, st, slotsDev, slot, smoothMap, mipmapMask, mappedRGB, recon0, recon1, recon2);
}
void yk_launch_range_dyn_count(const YkSlotDev* slotsDev, int slot, int cx, int cy, int cw, int ch, int nBlocks, cudaStream_t st) {
    YK_LAUNCH(yk_k_r1_count, dim3((nBlocks + 255) / 256), dim3(256), 0, st, slotsDev, slot, cx, cy, cw, ch, nBlocks);
}
void yk_launch_range_dyn_scan(const YkSlotDev* slotsDev, int slot, int nBlocks, int plane, cudaStream_t st) {
    YK_LAUNCH(yk_k_r1_scan, dim3(1), dim3(1024), 0, st, slotsDev, slot, nBlocks, plane);
}
void yk_launch_range_dyn_encode(const YkSlotDev* slotsDev, int slot, int plane, int mode3, int cx, int cy, int cw, int ch,
                                int nBlocks, const int* lutDev, cudaStream_t st) {
    const int per = YK_THREADS / 32;
    YK_LAUNCH(yk_k_r1_encode, dim3((nBlocks + per - 1) / per), dim3(YK_THREADS), 0, st, slotsDev, slot, plane, mode3, cx, cy, cw, ch, nBlocks, lutDev);
}
