// yaik_b200 — auxiliary sm_100a kernels of the YAIK encoder-analysis stage (the hot path is yk_analyze.cu + yk_emit.cu):
//
//   yk_k_state        expands the compact masks into the reference's int32 state planes (compat download).
//   yk_k_r1_encode    DynamicTileEncode (EC.cpp:4365-4503, 747-1212), LUT search at 3/4 bits per pixel, on the colour planes
//                     or on the Y / reduced Co / Cg planes of the chroma front-end.
//   yk_k_chroma       RGB -> YCoCg + SampleDown of the chroma planes (Image.cpp:285-321, Plane.cpp:278-369).
//
// Reference line numbers are KLab/YAIK's encoder/EncoderContext.cpp ("EC.cpp") unless another file is named.
#include "yk_device.h"



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
        if (S.rowBelow[c])                                       // strip mode: the real row below (typed like the uploaded planes)
            return S.isU8 ? (int)__ldg(reinterpret_cast<const uint8_t*>(S.rowBelow[c]) + x) : __ldg(reinterpret_cast<const int32_t*>(S.rowBelow[c]) + x);
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

// One launch per plane.  A CTA takes a unit of YK_R1_UNIT consecutive blocks of the walk (global ticket): one thread
// per block counts its valid pixels, a block scan + one decoupled look-back turn the counts into nibble / tile-def
// offsets, the blocks that hold valid pixels are compacted into a list in shared memory and the warps encode them one
// block per warp.  Small units: the per-block work is a long dependent chain (min/max -> table -> search -> ordered
// float sum), so the kernel wants as many warps in flight as the SM holds.
//
// lut: [64 base6][176 range7][144] ints.  [0,72) = the six LUTs (16,16,16,8,8,8 entries) of DynamicTile::buildTable
// (EC.cpp:625-699); [72,144) = their decision thresholds, same layout.  Every LUT is non-decreasing, so the reference's
// "first strict minimum of |entry - value|" scan (EC.cpp:873-881) picks entry n over all earlier ones exactly when
// value > floor((L[n-1] + L[n]) / 2) with L[n] > L[n-1]; a repeated entry is never picked and inherits the threshold of
// the next distinct one (INT_MAX if none), which keeps the thresholds non-decreasing: code = #{n >= 1 : value > thr[n]}.
#ifndef YK_R1_UNIT
#define YK_R1_UNIT 32
#endif
#ifndef YK_R1_THREADS
#define YK_R1_THREADS 128
#endif
#define YK_R1_LUT_INTS 144

// Validity of one full-resolution pixel as DynamicTileEncode sees it: mipmapMask != 0 (kept by the alpha stage) and
// smoothMap == 0 (its 4x4 cell not claimed by a gradient tile).
static __device__ __forceinline__ bool yk_r1_mask_at(const YkSlotDev& S, int fx, int fy) {
    return !(S.alphaValid && !S.alphaReset && !S.alphaKept[(size_t)(fy >> 4) * ((S.w + 15) >> 4) + (fx >> 4)]);
}
static __device__ __forceinline__ bool yk_r1_smooth_at(const YkSlotDev& S, int fx, int fy) {
    const int cx = fx >> 2;
    return (S.cellMask[(size_t)(fy >> 2) * S.nbx + (cx >> 4)] >> (cx & 15)) & 1u;
}

// mipmapMask != 0 at linear index i of the full-size mask plane
static __device__ __forceinline__ bool yk_r1_valid_at(const YkSlotDev& S, size_t i) {
    const int fx = (int)(i % (size_t)S.w), fy = (int)(i / (size_t)S.w);
    return yk_r1_mask_at(S, fx, fy) && !yk_r1_smooth_at(S, fx, fy);
}

// The walk of LeftRightOrder (framework.h:228-256) over the constraint box of a plane of pw x ph samples: rows of
// nbw = ceil(cw / 8) blocks while y < cy + ch, then one more block at the start of the next row when that row still
// lies inside the plane (HasNextBlock tests `y < h` after the wrap).  Block sizes: the reference compares with the
// constraint's width / height, not its right / bottom edge, and falls back to x % 8 (framework.h:251-252) - 0 for the
// full-resolution planes, possibly 4 for a reduced chroma plane whose box starts on an odd multiple of 4.
struct YkR1Block { int x, y, rw, rh; };
static __device__ __forceinline__ YkR1Block yk_r1_block(const YkR1Args& A, int i) {
    YkR1Block b;
    b.x = A.cx + 8 * (i % A.nbw); b.y = A.cy + 8 * (i / A.nbw);
    b.rw = (b.x + 8 > A.cw) ? (b.x & 7) : 8;
    b.rh = (b.y + 8 > A.ch) ? (b.y & 7) : 8;
    return b;
}

#ifndef YK_R1_MINB
#define YK_R1_MINB 10      // 48 registers: 3-bit-only launches 44.6 -> 40.6 us, six-mode launches unchanged
#endif
__global__ void __launch_bounds__(YK_R1_THREADS, YK_R1_MINB)
yk_k_r1_encode(const YkSlotDev* __restrict__ slots, int slot, const YkR1Args A, const int* __restrict__ lut) {
    __shared__ __align__(16) float sTerm[YK_R1_THREADS / 32][6][64];
    __shared__ int sBlock[YK_R1_UNIT], sNibOff[YK_R1_UNIT];      // sBlock: block index of the walk
    __shared__ unsigned sMask[YK_R1_UNIT];                       // coded pixel pairs of the block: bit = 4 * row + pair
    __shared__ int sWarp[33];
    __shared__ int sUnit;
    __shared__ unsigned sBase[2];
    const YkSlotDev& S = slots[slot];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nUnits = (A.nBlocks + YK_R1_UNIT - 1) / YK_R1_UNIT;
    const bool reduced = (A.shX | A.shY) != 0;
    unsigned long long* status = S.r1Status;
    if (tid == 0) sUnit = (int)atomicAdd(reinterpret_cast<unsigned*>(&status[nUnits]), 1u);
    __syncthreads();
    const int u = sUnit;
    // ---- count: which pixel pairs of my block are coded (EC.cpp:826-861; the 2x2 / 2x1 / 1x2 mask samples a reduced
    // pixel covers lie in one 16x16 alpha tile, so "all of them" is the top-left one)
    const int i = u * YK_R1_UNIT + tid;
    unsigned pairs = 0;
    if (tid < YK_R1_UNIT && i < A.nBlocks) {
        const YkR1Block b = yk_r1_block(A, i);
        if (!reduced) {
            if (b.rw == 8 && b.rh == 8 && b.x + 8 <= S.w && b.y + 8 <= S.h) {
                const unsigned cells = yk_r1_cells(S, b.x, b.y);
                pairs = ((cells & 1u) ? 0x00003333u : 0u) | ((cells & 2u) ? 0x0000CCCCu : 0u) | ((cells & 4u) ? 0x33330000u : 0u) | ((cells & 8u) ? 0xCCCC0000u : 0u);
            }
        } else {
            for (int r = 0; r < b.rh; r++)
                for (int p = 0; 2 * p < b.rw; p++) {
                    const int fx = (b.x + 2 * p) << A.shX, fy = (b.y + r) << A.shY;
                    if (yk_r1_mask_at(S, fx, fy) && !yk_r1_smooth_at(S, fx, fy)) pairs |= 1u << (4 * r + p);
                }
        }
    }
    const int n = 2 * __popc(pairs);
    // ---- offsets: pixels in the low 16 bits (<= 64 * YK_R1_UNIT per unit), blocks with pixels above
    int tot;
    const int ex = yk_block_exclusive(n | ((n > 0) << 16), sWarp, tot);
    if (warp == 0) {
        const unsigned long long base = yk_lookback64(status, u, (unsigned)(tot & 0xFFFF), (unsigned)(tot >> 16));
        if (lane == 0) {
            sBase[0] = (unsigned)(base >> 32); sBase[1] = (unsigned)base;
            if (u == nUnits - 1) {
                S.hdr[YK_HD_R1_NIB0 + A.out] = (int)(base >> 32) + (tot & 0xFFFF);
                S.hdr[YK_HD_R1_DEF0 + A.out] = (int)(unsigned)base + (tot >> 16);
            }
        }
    }
    if (n > 0) { sBlock[ex >> 16] = i; sMask[ex >> 16] = pairs; sNibOff[ex >> 16] = ex & 0xFFFF; }
    __syncthreads();
    const int nList = tot >> 16;
    const int nibBase = (int)sBase[0], defBase = (int)sBase[1];
    const int r = lane >> 2, c0 = (lane & 3) * 2;
    const int startMode = A.mode3 ? 3 : 0;
    for (int k = warp; k < nList; k += YK_R1_THREADS / 32) {
        const YkR1Block b = yk_r1_block(A, sBlock[k]);
        const int x = b.x, y = b.y;
        const bool valid = (sMask[k] >> lane) & 1u;
        // ---- min / max of the block (Plane::GetMinMax_Y, Plane.cpp:489-587).  Full resolution: over the coded pixels.
        // Reduced planes: over "not smooth and any covered mask sample set", where the reference addresses the
        // full-size mask with the REDUCED plane's width as row stride (Plane.cpp:516, 538-553) - restated literally.
        bool m0 = valid, m1 = valid;
        const bool inRect = r < b.rh && c0 < b.rw;
        if (reduced) {
            m0 = m1 = false;
            if (inRect) {
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const size_t vi = ((size_t)(x + c0 + q) << A.shX) + (size_t)((y + r) << A.shY) * A.pw;
                    const int W = S.w;
                    // every covered sample is a mipmapMask value: kept by the alpha stage AND not claimed by a gradient
                    // tile (FittingQuadSmooth zeroes the mask over accepted tiles, EC.cpp:4035); with the reduced width as
                    // row stride the samples of the second row lie in another cell than vi
                    bool any = yk_r1_valid_at(S, vi);
                    if (A.shX) any |= yk_r1_valid_at(S, vi + 1);
                    if (A.shY) any |= yk_r1_valid_at(S, vi + A.pw);
                    if (A.shX && A.shY) any |= yk_r1_valid_at(S, vi + A.pw + 1);
                    const bool ok = any && !yk_r1_smooth_at(S, (int)(vi % W), (int)(vi / W));
                    if (q) m1 = ok; else m0 = ok;
                }
            }
        }
        int v0 = 0, v1 = 0;
        if (valid || m0 || m1) {
            int2 p = __ldg(reinterpret_cast<const int2*>(A.src + (size_t)(y + r) * A.pw + x + c0));
            v0 = p.x; v1 = p.y;
        }
        int mn = __reduce_min_sync(YK_FULL, min(m0 ? v0 : INT_MAX, m1 ? v1 : INT_MAX));
        int mx = __reduce_max_sync(YK_FULL, max(m0 ? v0 : INT_MIN, m1 ? v1 : INT_MIN));
        if (mn == INT_MAX) { mn = 0; mx = 0; }                                 // Plane.cpp:579-585
        int sgn = 0;
        if (mn < 0) { mn += 128; mx += 128; sgn = 128; }                       // EC.cpp:764-768
        mn = min(max(mn, 0), 255); mx = min(max(mx, mn), 255);
        // DynamicTile::buildTable index (EC.cpp:635-650)
        const int m = min(mn, 224);
        const int diff = max(mx - m, 16);
        const int b6 = (m * 63 + 112) / 224, BN = (b6 * 224) / 63;
        const int scale = 223 - BN;
        const int r7 = ((max(diff, 32) - 32) * 127 + scale - 1) / scale;
        const int* T = lut + ((size_t)b6 * 176 + min(r7, 175)) * YK_R1_LUT_INTS;
        const int o0 = v0 + sgn, o1 = v1 + sgn;
        unsigned codes0 = 0, codes1 = 0;                                       // 4 bits per mode
#pragma unroll
        for (int mode = 0; mode < 6; mode++) {
            if (mode < startMode) continue;
            const int off = mode < 3 ? 16 * mode : 48 + 8 * (mode - 3);
            const int4* H = reinterpret_cast<const int4*>(T + 72 + off);
            int f0 = 0, f1 = 0;
#pragma unroll
            for (int q = 0; q < (mode < 3 ? 4 : 2); q++) {
                const int4 t = __ldg(H + q);
                if (q) { f0 += (o0 > t.x); f1 += (o1 > t.x); }                  // threshold 0 of a LUT is unused
                f0 += (o0 > t.y); f1 += (o1 > t.y);
                f0 += (o0 > t.z); f1 += (o1 > t.z);
                f0 += (o0 > t.w); f1 += (o1 > t.w);
            }
            const int d0 = abs(__ldg(T + off + f0) - o0), d1 = abs(__ldg(T + off + f1) - o1);
            codes0 |= (unsigned)f0 << (4 * mode); codes1 |= (unsigned)f1 << (4 * mode);
            // cumulated relative error term, float32 (EC.cpp:884-886); invalid pixels add +0.0f which leaves the sum unchanged.
            // The quotient is taken on operands that are never zero (a zero on either side sends the whole warp down the
            // slow path of the IEEE division) and dropped afterwards: 0 / o is +-0 and adds nothing either.
            const float q0 = (float)(d0 ? d0 : 1) / (float)(o0 ? o0 : 1), q1 = (float)(d1 ? d1 : 1) / (float)(o1 ? o1 : 1);
            sTerm[warp][mode][2 * lane] = (valid && o0 != 0 && d0 != 0) ? q0 : 0.0f;
            sTerm[warp][mode][2 * lane + 1] = (valid && o1 != 0 && d1 != 0) ? q1 : 0.0f;
        }
        __syncwarp();
        float err = 0.0f;
        if (lane >= startMode && lane < 6) {
            // the reference sums in row-major valid-pixel order; float addition is not associative, so one lane per mode
            const float4* P = reinterpret_cast<const float4*>(sTerm[warp][lane]);
#pragma unroll 4
            for (int q = 0; q < 16; q++) {
                const float4 t = P[q];
                err = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(err, t.x), t.y), t.z), t.w);
            }
        }
        int bestMode = -1;
        float bestErr = 99999999.0f;
        for (int mode = startMode; mode < 6; mode++) {                          // `<=`: later modes win ties, EC.cpp:897-905
            const float e = __shfl_sync(YK_FULL, err, mode);
            if (e <= bestErr) { bestErr = e; bestMode = mode; }
        }
        __syncwarp();                                                           // sTerm is rewritten by the next block
        const unsigned bv0 = __ballot_sync(YK_FULL, valid);
        if (valid) {
            const int before = 2 * __popc(bv0 & ((1u << lane) - 1u));           // both pixels of a lane share validity
            const int c0v = (codes0 >> (4 * bestMode)) & 15, c1v = (codes1 >> (4 * bestMode)) & 15;
            const int n0 = nibBase + sNibOff[k] + before;                       // even: the two nibbles share a byte
            reinterpret_cast<uint8_t*>(S.r1Nib[A.out])[n0 >> 1] = (uint8_t)(c0v | (c1v << 4));   // low nibble first, EC.cpp:1180-1184
            if (S.r1Dst) {
                // EC.cpp:4441-4502: chroma blocks with a negative minimum go back minus 128; reduced planes are
                // written at their top-left full-size position only ("interpolation comes later")
                const int* L = T + (bestMode < 3 ? 16 * bestMode : 48 + 8 * (bestMode - 3));
                const int offset = (A.chroma && sgn) ? -128 : 0;
                int* d = S.r1Dst + (size_t)((y + r) << A.shY) * S.w + ((x + c0) << A.shX);
                d[0] = __ldg(L + c0v) + offset; d[1 << A.shX] = __ldg(L + c1v) + offset;
            }
        }
        if (lane == 0) S.r1Defs[A.out][defBase + k] = (uint16_t)((bestMode << 13) | (r7 << 7) | b6);    // EncodeTileType, YAIK_private.h:358
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Chroma front-end: Image::ConvertToRGB2YCoCg (Image.cpp:285-321, RGBtoYCoCg EC.cpp:53-67) fused
// with EncoderContext::chromaReduction (EC.cpp:2770-2782) = Plane::SampleDown (Plane.cpp:278-369) of Co and of Cg.
// One thread per 2x2 quad of the image: reads the three colour planes once, writes Y at full size and each chroma
// plane at its own reduction.  C `/` throughout (truncation towards zero: the chroma samples are signed).
static __device__ __forceinline__ void yk_reduce_quad(int a, int b, int c, int d, int hx, int hy, int mode,
                                                      int32_t* __restrict__ out, int ow, int qx, int qy) {
    // a b / c d = the quad at (2qx, 2qy); EDownSample (framework.h:60-66): 0 NEAREST_TL 1 NEAREST_BR 2 AVERAGE_BOX 3 MAX_BOX 4 MIN_BOX
    if (hx && hy) {
        int v = a;
        if (mode == 2) v = (a + b + c + d) / 4;
        else if (mode == 1) v = d;
        else if (mode == 3) v = max(max(a, b), max(c, d));
        else if (mode == 4) v = min(min(a, b), min(c, d));
        out[(size_t)qy * ow + qx] = v;
    } else if (hx) {                          // modes 0 and 2 only (the API refuses the others on one axis)
        out[(size_t)(2 * qy) * ow + qx] = mode == 2 ? (a + b) / 2 : a;
        out[(size_t)(2 * qy + 1) * ow + qx] = mode == 2 ? (c + d) / 2 : c;
    } else if (hy) {
        *reinterpret_cast<int2*>(out + (size_t)qy * ow + 2 * qx) = mode == 2 ? make_int2((a + c) / 2, (b + d) / 2) : make_int2(a, b);
    } else {
        *reinterpret_cast<int2*>(out + (size_t)(2 * qy) * ow + 2 * qx) = make_int2(a, b);
        *reinterpret_cast<int2*>(out + (size_t)(2 * qy + 1) * ow + 2 * qx) = make_int2(c, d);
    }
}

__global__ void __launch_bounds__(256)
yk_k_chroma(const YkSlotDev* __restrict__ slots, int slot, const YkChromaArgs A) {
    const YkSlotDev& S = slots[slot];
    const int qw = S.w >> 1, qh = S.h >> 1;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= qw * qh) return;
    const int qx = q % qw, qy = q / qw;
    int Y[4], Co[4], Cg[4];
#pragma unroll
    for (int row = 0; row < 2; row++) {
        const size_t at = (size_t)(2 * qy + row) * S.w + 2 * qx;
        const int2 R = __ldg(reinterpret_cast<const int2*>(S.plane[0] + at));
        const int2 G = __ldg(reinterpret_cast<const int2*>(S.plane[1] + at));
        const int2 B = __ldg(reinterpret_cast<const int2*>(S.plane[2] + at));
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int r = k ? R.y : R.x, g = k ? G.y : G.x, b = k ? B.y : B.x;
            const int co = r - b, tmp = b + co / 2, cg = g - tmp;             // EC.cpp:55-66
            Y[2 * row + k] = tmp + cg / 2; Co[2 * row + k] = co / 2; Cg[2 * row + k] = cg / 2;
        }
    }
    *reinterpret_cast<int2*>(A.y + (size_t)(2 * qy) * S.w + 2 * qx) = make_int2(Y[0], Y[1]);
    *reinterpret_cast<int2*>(A.y + (size_t)(2 * qy + 1) * S.w + 2 * qx) = make_int2(Y[2], Y[3]);
    yk_reduce_quad(Co[0], Co[1], Co[2], Co[3], A.half[0], A.half[1], A.mode[0], A.co, A.half[0] ? qw : S.w, qx, qy);
    yk_reduce_quad(Cg[0], Cg[1], Cg[2], Cg[3], A.half[2], A.half[3], A.mode[1], A.cg, A.half[2] ? qw : S.w, qx, qy);
}

// ------------------------------------------------------------------------------------------------------------------// launch wrappers
void yk_launch_state(const YkSlotDev* slotsDev, int slot, int nRegions, int32_t* smoothMap, int32_t* mipmapMask,
                     int32_t* mappedRGB, int32_t* recon0, int32_t* recon1, int32_t* recon2, cudaStream_t st) {
    YK_LAUNCH(yk_k_state, dim3(nRegions), dim3(YK_THREADS), 0, st, slotsDev, slot, smoothMap, mipmapMask, mappedRGB, recon0, recon1, recon2);
}
void yk_launch_range_dyn_encode(const YkSlotDev* slotsDev, int slot, const YkR1Args& args, const int* lutDev, cudaStream_t st) {
    YK_LAUNCH(yk_k_r1_encode, dim3((args.nBlocks + YK_R1_UNIT - 1) / YK_R1_UNIT), dim3(YK_R1_THREADS), 0, st, slotsDev, slot, args, lutDev);
}
void yk_launch_chroma(const YkSlotDev* slotsDev, int slot, int w, int h, const YkChromaArgs& args, cudaStream_t st) {
    const int quads = (w >> 1) * (h >> 1);
    YK_LAUNCH(yk_k_chroma, dim3((quads + 255) / 256), dim3(256), 0, st, slotsDev, slot, args);
}
