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
//   yk_k_emit_count   corner ownership (EC.cpp:4001-4021, 4115-4132): a lattice point is emitted by the
//                     first pass that touches it, by the accepted tile with the smallest stream position.
//   yk_k_scan         exclusive scans that turn per-swizzle-block byte counts into stream offsets.
//   yk_k_emit_write   writes rgbStream in stream order.
//   yk_k_range1d      DynamicTileCompressor (EC.cpp:8398-8522), one warp per 8x8 tile and plane.
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
static __device__ __forceinline__ YkGeomC yk_geom(int pid) {
    const YkGeomC t[YK_NPASS] = YK_PASS_TABLE;
    return t[pid];
}

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

static __device__ __forceinline__ unsigned yk_pack4(int4 v) {
    return (unsigned)(v.x & 255) | ((unsigned)(v.y & 255) << 8) | ((unsigned)(v.z & 255) << 16) | ((unsigned)(v.w & 255) << 24);
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

// One FittingQuadSmooth pass (tile 1<<SHX by 1<<SHY, swizzle block BW x BH) over the staged region.
// Step 1: one thread per tile — eligibility (top-left cell unclaimed, EC.cpp:3871-3875; tile fully inside,
//         EC.cpp:3818/3826) and a one-quad pre-test that can only prove rejection; survivors are compacted into sList.
// Step 2: G = N/16 lanes per surviving tile evaluate 16 pixels each, family by family, voting after every quad.
template <int SHX, int SHY, int BW, int BH>
static __device__ void yk_pass(const uint8_t (*pix)[65 * YK_RS], uint32_t* sCell, uint32_t* sBits, int* sStat,
                               uint16_t* sList, int* sCount, int X0, int Y0, int w, int h, int yOrg, int R) {
    constexpr int TW = 1 << SHX, TH = 1 << SHY, N = TW * TH, NX = 64 / TW, NY = 64 / TH, NT = NX * NY;
    constexpr int G = N / 16, QR = TW / 4, BITS = (BW / TW) * (BH / TH);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hiT = (2 * R + 1) * N;                    // exclusive upper bound of U for the truncated variant
    const int loR = -(N / 2 - 1);                       // inclusive lower bound of U for the rounded variant
    // a family's corners differ from the raw ones by -3..+4 (Round6: -3..+3, Round6P: -2..+4), so no variant can
    // accept a pixel whose raw-family U is outside [loWide, hiWide)
    const int loWide = -(4 * N + N / 2 - 1), hiWide = hiT + 3 * N;

    bool cand = false;
    if (tid < NT) {
        int lx0 = (tid % NX) * TW, ly0 = (tid / NX) * TH;
        bool inside = (X0 + lx0 + TW <= w) && (Y0 + ly0 + TH <= h);
        bool claimed = (sCell[ly0 >> 2] >> (lx0 >> 2)) & 1u;
        if (inside && !claimed) {
            int umin = INT_MAX, umax = INT_MIN;
            constexpr int dx0 = 4 * (QR / 2), dy = TH / 2;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const uint8_t* p = pix[c];
                int tl = p[ly0 * YK_RS + lx0], tr = p[ly0 * YK_RS + lx0 + TW];
                int bl = p[(ly0 + TH) * YK_RS + lx0], br = p[(ly0 + TH) * YK_RS + lx0 + TW];
                yk_quad<N>(p, (ly0 + dy) * YK_RS + lx0 + dx0, dx0, dy, tl * N + R * N, TH * (tr - tl), TW * (bl - tl), tl - tr - bl + br, umin, umax);
            }
            cand = !(umin < loWide || umax >= hiWide);
        }
    }
    {
        unsigned b = __ballot_sync(YK_FULL, cand);
        int base = 0;
        if (lane == 0 && b) base = atomicAdd(sCount, __popc(b));
        base = __shfl_sync(YK_FULL, base, 0);
        if (cand) sList[base + __popc(b & ((1u << lane) - 1u))] = (uint16_t)tid;
    }
    __syncthreads();

    const int nCand = *sCount;
    constexpr int TPW = 32 / G;
    const int j = lane % G, slot = lane / G;
    const unsigned gmask = (G == 32) ? YK_FULL : (((1u << (G & 31)) - 1u) << (slot * G));
    for (int base = warp * TPW; base < nCand; base += (YK_THREADS / 32) * TPW) {
        const int ci = base + slot;
        const bool active = ci < nCand;
        const int t = active ? sList[ci] : 0;
        const int lx0 = (t % NX) * TW, ly0 = (t / NX) * TH;
        int cr[3][4];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const uint8_t* p = pix[c];
            cr[c][0] = p[ly0 * YK_RS + lx0]; cr[c][1] = p[ly0 * YK_RS + lx0 + TW];
            cr[c][2] = p[(ly0 + TH) * YK_RS + lx0]; cr[c][3] = p[(ly0 + TH) * YK_RS + lx0 + TW];
        }
        bool accepted = false, resolved = !active;
#pragma unroll
        for (int fam = 0; fam < 3; fam++) {
            int A3[3], B[3], C[3], D[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                int tl, tr, bl, br;
                if (fam == 0) { tl = cr[c][0]; tr = cr[c][1]; bl = cr[c][2]; br = cr[c][3]; }
                else if (fam == 1) { tl = yk_round6(cr[c][0]); tr = yk_round6(cr[c][1]); bl = yk_round6(cr[c][2]); br = yk_round6(cr[c][3]); }
                else { tl = yk_round6p(cr[c][0]); tr = yk_round6p(cr[c][1]); bl = yk_round6p(cr[c][2]); br = yk_round6p(cr[c][3]); }
                A3[c] = tl * N + R * N; B[c] = TH * (tr - tl); C[c] = TW * (bl - tl); D[c] = tl - tr - bl + br;
            }
            int umin = INT_MAX, umax = INT_MIN;
            bool famDead = false;
            for (int k = 0; k < 4; k++) {
                const int q = j + G * k, dx0 = 4 * (q % QR), dy = q / QR;
                if (!resolved && !famDead) {
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        yk_quad<N>(pix[c], (ly0 + dy) * YK_RS + lx0 + dx0, dx0, dy, A3[c], B[c], C[c], D[c], umin, umax);
                }
                const bool dT = (umin < 0) || (umax >= hiT);
                const bool dR = (umin < loR) || (umax >= hiT + loR);
                const unsigned bT = __ballot_sync(YK_FULL, dT), bR = __ballot_sync(YK_FULL, dR);
                famDead = ((bT & gmask) != 0u) && ((bR & gmask) != 0u);
                if (fam == 0) {
                    const unsigned bH = __ballot_sync(YK_FULL, (umin < loWide) || (umax >= hiWide));
                    if (bH & gmask) resolved = true;        // no family can accept this tile
                }
                if (!__any_sync(YK_FULL, !resolved && !famDead)) break;
            }
            if (!resolved && !famDead) { accepted = true; resolved = true; }     // EC.cpp:3998: any surviving variant accepts
            if (!__any_sync(YK_FULL, !resolved)) break;
        }
        if (accepted && j == 0) {
            const int sub = (ly0 / BH) * (64 / BW) + (lx0 / BW);
            const int li = sub * BITS + ((ly0 % BH) / TH) * (BW / TW) + (lx0 % BW) / TW;
            atomicOr(&sBits[li >> 5], 1u << (li & 31));                          // EC.cpp:4026
#pragma unroll
            for (int r = 0; r < TH / 4; r++)                                     // EC.cpp:4029-4037
                atomicOr(&sCell[(ly0 >> 2) + r], ((1u << (TW / 4)) - 1u) << (lx0 >> 2));
            atomicAdd(&sStat[YK_ST_TILEDONE], 1);                                // EC.cpp:4039-4044 (mins stored as extent - value)
            atomicMax(&sStat[YK_ST_MINX], w - (X0 + lx0));
            atomicMax(&sStat[YK_ST_MINY], INT_MAX / 2 - (yOrg + Y0 + ly0));
            atomicMax(&sStat[YK_ST_MAXX], X0 + lx0 + TW);
            atomicMax(&sStat[YK_ST_MAXY], yOrg + Y0 + ly0 + TH);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(YK_THREADS, 3)
yk_k_analyze(const YkSlotDev* __restrict__ slots, int slot0, YkRun run) {
    __shared__ __align__(16) uint8_t pix[3][65 * YK_RS];
    __shared__ uint32_t sCell[16];
    __shared__ uint32_t sBits[YK_NPASS][8];
    __shared__ int sStat[YK_NPASS][YK_ST_STRIDE];
    __shared__ uint16_t sList[256];
    __shared__ int sCount[YK_NPASS];
    __shared__ uint32_t sAlpha;

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
    if (tid < YK_NPASS) sCount[tid] = 0;
    if (tid == 0) sAlpha = 0;
    __syncthreads();

    unsigned bad = 0;
    yk_stage_pixels(S, X0, Y0, pix, bad);

    // ---- alpha-zero tile rejection: all(alpha == 0) per 16x16 tile (EC.cpp:357-430 restated per tile) by ballot
    if (run.doAlpha && S.nPlanes == 4) {
        const int32_t* __restrict__ P = S.plane[3];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int ly = (tid >> 4) + 16 * k, lx = (tid & 15) * 4;
            int y = Y0 + ly, x = X0 + lx;
            int nz = 0;
            if (y < h && x + 3 < w) {
                int4 v = __ldg(reinterpret_cast<const int4*>(P + (size_t)y * w + x));
                nz = v.x | v.y | v.z | v.w;
            } else if (y < h) {
                for (int i = 0; i < 4; i++) if (x + i < w) nz |= __ldg(P + (size_t)y * w + x + i);
            }
            unsigned b = __ballot_sync(YK_FULL, nz != 0);
            if (lane == 0 && b) {
                unsigned m = 0;
#pragma unroll
                for (int tx = 0; tx < 4; tx++) if (b & (0x000F000Fu << (4 * tx))) m |= 1u << (k * 4 + tx);
                atomicOr(&sAlpha, m);
            }
        }
    }
    if (bad & ~255u) atomicOr(&S.hdr[YK_HD_ERR], 1);
    __syncthreads();

    const int R = run.rejectFactor;
    for (int rp = 0; rp < run.nPasses; rp++) {
        const int pid = run.passId[rp];
        switch (pid) {      // Convert()'s order, EC.cpp:9057-9093
        case 0: yk_pass<4, 4, 64, 64>(pix, sCell, sBits[0], sStat[0], sList, &sCount[rp], X0, Y0, w, h, S.y0, R); break;
        case 1: yk_pass<4, 3, 64, 64>(pix, sCell, sBits[1], sStat[1], sList, &sCount[rp], X0, Y0, w, h, S.y0, R); break;
        case 2: yk_pass<3, 4, 64, 64>(pix, sCell, sBits[2], sStat[2], sList, &sCount[rp], X0, Y0, w, h, S.y0, R); break;
        case 3: yk_pass<3, 3, 64, 64>(pix, sCell, sBits[3], sStat[3], sList, &sCount[rp], X0, Y0, w, h, S.y0, R); break;
        case 4: yk_pass<3, 2, 64, 32>(pix, sCell, sBits[4], sStat[4], sList, &sCount[rp], X0, Y0, w, h, S.y0, R); break;
        case 5: yk_pass<2, 3, 32, 64>(pix, sCell, sBits[5], sStat[5], sList, &sCount[rp], X0, Y0, w, h, S.y0, R); break;
        default: yk_pass<2, 2, 32, 32>(pix, sCell, sBits[6], sStat[6], sList, &sCount[rp], X0, Y0, w, h, S.y0, R); break;
        }
    }

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
    // per 8-tile segment of DynamicTileCompressor's row-major tile order: 16-byte chunks to code and coded tiles
    if (tid >= 32 && tid < 40) {
        const int t = tid - 32, ty = (Y0 >> 3) + t;
        if (ty * 8 < h) {
            const uint32_t c0 = sCell[2 * t], c1 = sCell[2 * t + 1];
            int chunks = 0, tiles = 0;
#pragma unroll
            for (int x = 0; x < 8; x++) {
                int n = 4 - __popc(((c0 >> (2 * x)) & 3u) | (((c1 >> (2 * x)) & 3u) << 2));
                chunks += n; tiles += (n > 0);
            }
            S.r2Seg[(size_t)ty * nbx + bx] = chunks | (tiles << 16);
        }
    }
    // corner colours at every 4-pixel lattice point of the region (what an accepted tile would emit, EC.cpp:4115-4132)
    {
        const int iMax = (bx == nbx - 1) ? 17 : 16, jMax = (by == S.nby - 1) ? 17 : 16;
        for (int idx = tid; idx < 17 * 17; idx += YK_THREADS) {
            const int i = idx % 17, jj = idx / 17;
            const int gx = (X0 >> 2) + i, gy = (Y0 >> 2) + jj;
            if (i < iMax && jj < jMax && gx < S.latW && gy < S.latH) {
                uint8_t* d = S.latRGB + ((size_t)gy * S.latW + gx) * 3;
#pragma unroll
                for (int c = 0; c < 3; c++) d[c] = (uint8_t)yk_compress250(yk_round6(pix[c][(4 * jj) * YK_RS + 4 * i]));
            }
        }
    }
    if (run.doAlpha && S.nPlanes == 4 && tid < 32) {
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

// ------------------------------------------------------------------------------------------------------------------
// Corner ownership.  The reference walks tiles in stream order and lets a tile emit a corner colour only if no
// earlier tile (of this or an earlier pass) touched that lattice point (mappedRGB, EC.cpp:4001-4021, 4115-4132).
// Order-free: firstClaim[L] = first pass of this launch with an accepted tile touching L (0 = claimed before the
// launch); in that pass the accepted toucher with the smallest stream position emits L.
static __device__ __forceinline__ int yk_accept_bit(const YkSlotDev& S, int pid, const YkGeomC& g, int gtx, int gty) {
    if (gtx < 0 || gty < 0 || ((gtx + 1) << g.shx) > S.w || ((gty + 1) << g.shy) > S.h) return 0;
    int pos = yk_tile_pos(g, S.w, gtx, gty);
    return (S.bitmap[pid][pos >> 3] >> (pos & 7)) & 1;
}

__global__ void __launch_bounds__(YK_THREADS)
yk_k_emit_count(const YkSlotDev* __restrict__ slots, int slot0, YkRun run) {
    __shared__ int fc[17 * 17];
    __shared__ uint8_t acc[YK_NPASS][18 * 18];
    __shared__ int unitCnt[YK_NPASS][4];
    const YkSlotDev& S = slots[slot0 + blockIdx.y];
    const int tid = threadIdx.x;
    const int w = S.w, h = S.h, nbx = S.nbx;
    const int bx = blockIdx.x % nbx, by = blockIdx.x / nbx;
    const int X0 = bx * YK_REGION, Y0 = by * YK_REGION;

    for (int idx = tid; idx < 17 * 17; idx += YK_THREADS) {
        const int gx = (X0 >> 2) + idx % 17, gy = (Y0 >> 2) + idx / 17;
        int claimed = 1;
        if (gx < S.latW && gy < S.latH) claimed = (S.cornerMask[(size_t)gy * S.cornerWords + (gx >> 5)] >> (gx & 31)) & 1u;
        fc[idx] = claimed ? 0 : 127;
    }
    if (tid < YK_NPASS * 4) (&unitCnt[0][0])[tid] = 0;
    __syncthreads();

    for (int rp = 0; rp < run.nPasses; rp++) {
        const int pid = run.passId[rp];
        const YkGeomC g = yk_geom(pid);
        const int NX = 64 >> g.shx, NY = 64 >> g.shy, EW = NX + 2;
        const int cw4 = (1 << g.shx) >> 2, ch4 = (1 << g.shy) >> 2;
        for (int idx = tid; idx < EW * (NY + 2); idx += YK_THREADS) {
            const int ex = idx % EW - 1, ey = idx / EW - 1;
            const int a = yk_accept_bit(S, pid, g, bx * NX + ex, by * NY + ey);
            acc[rp][idx] = (uint8_t)a;
            if (a) {
                const int cx0 = ex * cw4, cy0 = ey * ch4;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int cx = cx0 + (k & 1) * cw4, cy = cy0 + (k >> 1) * ch4;
                    if (cx >= 0 && cx <= 16 && cy >= 0 && cy <= 16) atomicMin(&fc[cy * 17 + cx], rp + 1);
                }
            }
        }
    }
    __syncthreads();

    for (int rp = 0; rp < run.nPasses; rp++) {
        const int pid = run.passId[rp];
        const YkGeomC g = yk_geom(pid);
        const int NX = 64 >> g.shx, NY = 64 >> g.shy, EW = NX + 2;
        const int cw4 = (1 << g.shx) >> 2, ch4 = (1 << g.shy) >> 2;
        const int tw = 1 << g.shx, th = 1 << g.shy;
        for (int idx = tid; idx < NX * NY; idx += YK_THREADS) {
            const int tx = idx % NX, ty = idx / NX;
            const int gtx = bx * NX + tx, gty = by * NY + ty;
            if (((gtx + 1) << g.shx) > w || ((gty + 1) << g.shy) > h) continue;
            const int myPos = yk_tile_pos(g, w, gtx, gty);
            int mask = 0;
            if (acc[rp][(ty + 1) * EW + tx + 1]) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int tX = tx + (k & 1), tY = ty + (k >> 1);           // lattice point in tile units
                    if (fc[(tY * ch4) * 17 + tX * cw4] != rp + 1) continue;
                    bool owner = true;
#pragma unroll
                    for (int o = 0; o < 4; o++) {
                        const int ox = tX - 1 + (o & 1), oy = tY - 1 + (o >> 1);
                        if (ox == tx && oy == ty) continue;
                        if (acc[rp][(oy + 1) * EW + ox + 1] && yk_tile_pos(g, w, bx * NX + ox, by * NY + oy) < myPos) owner = false;
                    }
                    if (owner) mask |= 1 << k;
                }
            }
            S.emitMask[pid][myPos] = (uint8_t)mask;
            if (mask) {
                const int lx0 = tx * tw, ly0 = ty * th;
                atomicAdd(&unitCnt[rp][(ly0 / g.bh) * (64 / g.bw) + lx0 / g.bw], 3 * __popc(mask));
            }
        }
    }
    __syncthreads();

    if (tid < run.nPasses * 4) {
        const int rp = tid >> 2, sub = tid & 3, pid = run.passId[rp];
        const YkGeomC g = yk_geom(pid);
        if (sub < (64 / g.bw) * (64 / g.bh)) {
            const int sx = X0 + (sub % (64 / g.bw)) * g.bw, sy = Y0 + (sub / (64 / g.bw)) * g.bh;
            if (sx < w && sy < h) S.unitOff[pid][(sy / g.bh) * ((w + g.bw - 1) / g.bw) + sx / g.bw] = unitCnt[rp][sub];
        }
    }
    // lattice points claimed by this launch
    if (tid >= 64 && tid < 64 + 17) {
        const int jj = tid - 64, gy = (Y0 >> 2) + jj;
        if (gy < S.latH) {
            unsigned long long bits = 0;
            for (int i = 0; i < 17; i++) {
                const int v = fc[jj * 17 + i];
                if (v > 0 && v < 127 && (X0 >> 2) + i < S.latW) bits |= 1ull << i;
            }
            if (bits) {
                const int gx0 = X0 >> 2;                 // multiple of 16
                bits <<= (gx0 & 31);
                uint32_t* row = S.cornerNew + (size_t)gy * S.cornerWords + (gx0 >> 5);
                if ((uint32_t)bits) atomicOr(row, (uint32_t)bits);
                if ((uint32_t)(bits >> 32)) atomicOr(row + 1, (uint32_t)(bits >> 32));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Exclusive scans: blockIdx.x < YK_NPASS -> rgb byte counts of that pass per swizzle block (stream order);
// blockIdx.x == YK_NPASS -> DynamicTileCompressor segments (chunks and coded tiles).
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

__global__ void __launch_bounds__(1024)
yk_k_scan(const YkSlotDev* __restrict__ slots, int slot0, YkRun run) {
    __shared__ int sWarp[33];
    const YkSlotDev& S = slots[slot0 + blockIdx.y];
    const int tid = threadIdx.x;
    if (blockIdx.x < YK_NPASS) {
        if ((int)blockIdx.x >= run.nPasses) return;
        const int pid = run.passId[blockIdx.x];
        const YkGeomC g = yk_geom(pid);
        const int n = ((S.w + g.bw - 1) / g.bw) * ((S.h + g.bh - 1) / g.bh);
        int* a = S.unitOff[pid];
        const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
        const int b = min(n, tid * per), e = min(n, b + per);
        int sum = 0;
        for (int i = b; i < e; i++) sum += a[i];
        int total;
        int runv = yk_block_exclusive(sum, sWarp, total);
        for (int i = b; i < e; i++) { int v = a[i]; a[i] = runv; runv += v; }
        if (tid == 0) S.hdr[YK_HD_PASS0 + pid * YK_ST_STRIDE + YK_ST_RGBBYTES] = total;
    } else {
        const int n = (S.h >> 3) * S.nbx;
        int* a = S.r2Seg; int* t2 = S.r2SegTiles;
        const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
        const int b = min(n, tid * per), e = min(n, b + per);
        int sc = 0, stl = 0;
        for (int i = b; i < e; i++) { int v = a[i]; sc += v & 0xFFFF; stl += v >> 16; }
        int totC, totT;
        int rc = yk_block_exclusive(sc, sWarp, totC);
        int rt = yk_block_exclusive(stl, sWarp, totT);
        for (int i = b; i < e; i++) { int v = a[i]; a[i] = rc; t2[i] = rt; rc += v & 0xFFFF; rt += v >> 16; }
        if (tid == 0) { S.hdr[YK_HD_R2_CHUNKS] = totC; S.hdr[YK_HD_R2_TILES] = totT; }
    }
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(YK_THREADS)
yk_k_emit_write(const YkSlotDev* __restrict__ slots, int slot0, YkRun run) {
    __shared__ int sWarp[33];
    __shared__ int sP[256];
    const YkSlotDev& S = slots[slot0 + blockIdx.y];
    const int tid = threadIdx.x;
    const int w = S.w, h = S.h, nbx = S.nbx;
    const int bx = blockIdx.x % nbx, by = blockIdx.x / nbx;
    const int X0 = bx * YK_REGION, Y0 = by * YK_REGION;

    for (int rp = 0; rp < run.nPasses; rp++) {
        const int pid = run.passId[rp];
        const YkGeomC g = yk_geom(pid);
        const int tw = 1 << g.shx, th = 1 << g.shy;
        const int nsub = (64 / g.bw) * (64 / g.bh), NT = nsub * g.bits;
        // thread = tile in stream order inside the region: li = sub * bits + row * (bw/tw) + col
        int m = 0, lx0 = 0, ly0 = 0, sub = 0;
        if (tid < NT) {
            sub = tid / g.bits;
            const int within = tid % g.bits, tpr = g.bw / tw;
            lx0 = (sub % (64 / g.bw)) * g.bw + (within % tpr) * tw;
            ly0 = (sub / (64 / g.bw)) * g.bh + (within / tpr) * th;
            if (X0 + lx0 + tw <= w && Y0 + ly0 + th <= h)
                m = S.emitMask[pid][yk_tile_pos(g, w, (X0 + lx0) >> g.shx, (Y0 + ly0) >> g.shy)];
        }
        int total;
        const int ex = yk_block_exclusive(3 * __popc(m), sWarp, total);
        sP[tid] = ex;
        __syncthreads();
        if (m) {
            const int sx = X0 + (sub % (64 / g.bw)) * g.bw, sy = Y0 + (sub / (64 / g.bw)) * g.bh;
            const int gb = (sy / g.bh) * ((w + g.bw - 1) / g.bw) + sx / g.bw;
            int off = S.unitOff[pid][gb] + ex - sP[sub * g.bits];
            uint8_t* out = S.rgb[pid];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if ((m >> k) & 1) {
                    const int gx = ((X0 + lx0) >> 2) + (k & 1) * (tw >> 2), gy = ((Y0 + ly0) >> 2) + (k >> 1) * (th >> 2);
                    const uint8_t* s = S.latRGB + ((size_t)gy * S.latW + gx) * 3;
                    out[off] = s[0]; out[off + 1] = s[1]; out[off + 2] = s[2];
                    off += 3;
                }
            }
        }
        __syncthreads();
    }
    // fold this launch's claims into the persistent corner mask (mappedRGB, EC.cpp:4005-4019)
    if (tid < 17 * 2) {
        const int jj = tid >> 1, gy = (Y0 >> 2) + jj, wi = ((X0 >> 2) >> 5) + (tid & 1);
        if (gy < S.latH && wi < S.cornerWords) {
            const uint32_t v = S.cornerNew[(size_t)gy * S.cornerWords + wi];
            if (v) atomicOr(&S.cornerMask[(size_t)gy * S.cornerWords + wi], v);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// DynamicTileCompressor (EC.cpp:8398-8522): one warp per (8x8 tile, plane); lane = 2 pixels.
__global__ void __launch_bounds__(YK_THREADS)
yk_k_range1d(const YkSlotDev* __restrict__ slots, int slot0) {
    __shared__ uint32_t hist[YK_THREADS / 32][256];
    __shared__ uint32_t sCell[16];
    const YkSlotDev& S = slots[slot0 + blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = S.w, h = S.h, nbx = S.nbx;
    const int bx = blockIdx.x % nbx, by = blockIdx.x / nbx;
    const int X0 = bx * YK_REGION, Y0 = by * YK_REGION;
    if (tid < 16) {
        int cy = (Y0 >> 2) + tid;
        uint32_t v = 0xFFFFu;
        if (cy * 4 < h) {
            v = S.cellMask[(size_t)cy * nbx + bx];
            int cellsIn = (w - X0) >> 2;
            if (cellsIn < 16) v |= (0xFFFFu << cellsIn) & 0xFFFFu;
        }
        sCell[tid] = v;
    }
    for (int i = tid; i < (YK_THREADS / 32) * 256; i += YK_THREADS) (&hist[0][0])[i] = 0;
    __syncthreads();

    const int r = lane >> 2, c0 = (lane & 3) * 2;           // pixel row / first column of this lane inside the tile
    const int band = r >> 2, right = c0 >> 2;
    for (int item = warp; item < 64 * 3; item += YK_THREADS / 32) {
        const int tile = item % 64, plane = item / 64;
        const int tx = tile & 7, ty = tile >> 3;
        if (X0 + 8 * tx + 8 > w || Y0 + 8 * ty + 8 > h) continue;
        const uint32_t r0 = sCell[2 * ty], r1 = sCell[2 * ty + 1];
        // quadrant needs coding iff its top-left map pixel is 0 (EC.cpp:8420-8430) == its 4x4 cell is unclaimed
        const unsigned q = (~(((r0 >> (2 * tx)) & 3u) | (((r1 >> (2 * tx)) & 3u) << 2))) & 15u;   // bit0 TL, 1 TR, 2 BL, 3 BR
        if (q == 0) continue;
        // offsets: scanned segment base + tiles to the left inside the segment
        int chunkOff, tileOff;
        {
            const size_t seg = (size_t)((Y0 >> 3) + ty) * nbx + bx;
            chunkOff = S.r2Seg[seg]; tileOff = S.r2SegTiles[seg];
            for (int x = 0; x < tx; x++) {
                int n = 4 - __popc(((r0 >> (2 * x)) & 3u) | (((r1 >> (2 * x)) & 3u) << 2));
                chunkOff += n; tileOff += (n > 0);
            }
        }
        const bool valid = (q >> (band * 2 + right)) & 1u;
        int vx = 0, vy = 0;
        if (valid) {
            int2 p = __ldg(reinterpret_cast<const int2*>(S.plane[plane] + (size_t)(Y0 + 8 * ty + r) * w + X0 + 8 * tx + c0));
            vx = p.x & 255; vy = p.y & 255;             // CompressF(v, 255) == v (EC.cpp:8442)
            atomicAdd(&hist[warp][vx], 1u); atomicAdd(&hist[warp][vy], 1u);
        }
        __syncwarp();
        // FindAndRemoveMostUsedColor (EC.cpp:8335-8356): highest index among the maximal counts
        unsigned key = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { unsigned cnt = hist[warp][lane * 8 + i]; key = max(key, (cnt << 8) | (unsigned)(lane * 8 + i)); hist[warp][lane * 8 + i] = 0; }
        key = __reduce_max_sync(YK_FULL, key);
        int color0 = (int)(key & 255u);
        color0 = min(max(color0, 1), 254);
        // Model1 (EC.cpp:8358-8381) over what is left of the histogram
        const bool remx = valid && (vx < color0 - 1 || vx > color0 + 1), remy = valid && (vy < color0 - 1 || vy > color0 + 1);
        int mn = min(remx ? vx : 999, remy ? vy : 999), mx = max(remx ? vx : -1, remy ? vy : -1);
        mn = __reduce_min_sync(YK_FULL, mn); mx = __reduce_max_sync(YK_FULL, mx);
        int minCol = 0, delta = 0;
        if (mn != 999) { minCol = mn; delta = mx - mn; }
        if (valid) {
            // GetValueModel1 (EC.cpp:8383-8391): C division; numerator -1 only when delta == 1
            int bxv = 0, byv = 0;
            if (remx) bxv = 1 + (delta ? ((vx - minCol) * 15 + ((delta >> 1) - 1)) / delta : 0);
            if (remy) byv = 1 + (delta ? ((vy - minCol) * 15 + ((delta >> 1) - 1)) / delta : 0);
            const unsigned qb = (q >> (band * 2)) & 3u;                     // coded quadrants of this band: bit0 left, bit1 right
            const int lengthX = (qb == 3u) ? 8 : 4, x2 = (qb == 2u) ? 4 : 0;
            const int bandBase = band ? 16 * __popc(q & 3u) : 0;
            const int pos = bandBase + (r & 3) * lengthX + (c0 - x2);
            uint8_t* d = S.r2Idx[plane] + (size_t)chunkOff * 16 + pos;
            *reinterpret_cast<uint16_t*>(d) = (uint16_t)((bxv & 255) | ((byv & 255) << 8));
        }
        if (lane == 0) {
            uint8_t* t = S.r2Type[plane] + (size_t)tileOff * 3;            // EC.cpp:8503-8505
            t[0] = (uint8_t)color0; t[1] = (uint8_t)minCol; t[2] = (uint8_t)delta;
        }
        __syncwarp();
    }
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
            if (!(x & 3) && !(y & 3)) v = ((S.cornerMask[(size_t)(y >> 2) * S.cornerWords + (x >> 7)] >> ((x >> 2) & 31)) & 1u) ? 255 : 0;
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
void yk_launch_emit_count(const YkSlotDev* slotsDev, int slot0, int nSlots, int nRegions, const YkRun& run, cudaStream_t st) {
    YK_LAUNCH(yk_k_emit_count, dim3(nRegions, nSlots), dim3(YK_THREADS), 0, st, slotsDev, slot0, run);
}
void yk_launch_scan(const YkSlotDev* slotsDev, int slot0, int nSlots, const YkRun& run, cudaStream_t st) {
    YK_LAUNCH(yk_k_scan, dim3(YK_NPASS + 1, nSlots), dim3(1024), 0, st, slotsDev, slot0, run);
}
void yk_launch_emit_write(const YkSlotDev* slotsDev, int slot0, int nSlots, int nRegions, const YkRun& run, cudaStream_t st) {
    YK_LAUNCH(yk_k_emit_write, dim3(nRegions, nSlots), dim3(YK_THREADS), 0, st, slotsDev, slot0, run);
}
void yk_launch_range1d(const YkSlotDev* slotsDev, int slot0, int nSlots, int nRegions, cudaStream_t st) {
    YK_LAUNCH(yk_k_range1d, dim3(nRegions, nSlots), dim3(YK_THREADS), 0, st, slotsDev, slot0);
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
