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

// Two launches code up to three planes that share their geometry and validity (the colour planes R, G, B; or one plane:
// Y / reduced Co / Cg of the chroma front-end):
//   yk_k_r1_offsets   one thread per block of the walk: which pixel pairs of the block are coded (that only depends on
//                     the alpha / claimed-cell masks, not on the samples), a block scan + one decoupled look-back per
//                     256 blocks give the nibble / tile-def offsets - the same for every plane - and the blocks that
//                     hold valid pixels are listed in walk order {x | y << 16, pairs, nibble offset}.
//   yk_k_r1_encode    warps take (listed block, plane) items with a grid stride and code them independently of each
//                     other (no block barrier, no waiting): min / max, table, mode search, ordered error sums, output
//                     at the final stream position.
//
// lut: [64 base6][176 range7][144] ints.  [0,72) = the six LUTs (16,16,16,8,8,8 entries) of DynamicTile::buildTable
// (EC.cpp:625-699); [72,144) = their decision thresholds, same layout.  Every LUT is non-decreasing, so the reference's
// "first strict minimum of |entry - value|" scan (EC.cpp:873-881) picks entry n over all earlier ones exactly when
// value > floor((L[n-1] + L[n]) / 2) with L[n] > L[n-1]; a repeated entry is never picked and inherits the threshold of
// the next distinct one (INT_MAX if none), which keeps the thresholds non-decreasing: code = #{n >= 1 : value > thr[n]}
// = the largest n with value > thr[n] (binary search).
// rtab: [64][128 range7 < 128][6 modes][384] u16 = code << 8 | LUT entry that search picks, for every value 0..383 (the
// domain of the reference's tables: samples -128..255, shifted by 128 when the block's minimum is negative), built on
// the host from the same thresholds: the per-pixel, per-mode work of the mode search is one 16-bit load.
#ifndef YK_R1_UNIT
#define YK_R1_UNIT 512
#endif
#ifndef YK_R1_THREADS
#define YK_R1_THREADS 128
#endif
#define YK_R1_LUT_INTS 144
#define YK_R1_RTAB_VALUES 384
#define YK_R1_RTAB_R7 128

// Validity of one full-resolution pixel as DynamicTileEncode sees it: mipmapMask != 0 (kept by the alpha stage) and
// smoothMap == 0 (its 4x4 cell not claimed by a gradient tile).
static __device__ __forceinline__ bool yk_r1_mask_at(const YkSlotDev& S, bool maskActive, int fx, int fy) {
    return !(maskActive && !S.alphaKept[(size_t)(fy >> 4) * ((S.w + 15) >> 4) + (fx >> 4)]);
}
static __device__ __forceinline__ bool yk_r1_smooth_at(const YkSlotDev& S, int fx, int fy) {
    const int cx = fx >> 2;
    return (S.cellMask[(size_t)(fy >> 2) * S.nbx + (cx >> 4)] >> (cx & 15)) & 1u;
}
// mipmapMask != 0 at linear index i of the full-size mask plane
static __device__ __forceinline__ bool yk_r1_valid_at(const YkSlotDev& S, bool maskActive, size_t i) {
    const int fx = (int)(i % (size_t)S.w), fy = (int)(i / (size_t)S.w);
    return yk_r1_mask_at(S, maskActive, fx, fy) && !yk_r1_smooth_at(S, fx, fy);
}
// unclaimed-and-kept 4x4 cells of the 8x8 block at (x, y): bit0 TL, bit1 TR, bit2 BL, bit3 BR
static __device__ __forceinline__ unsigned yk_r1_cells(const YkSlotDev& S, bool maskActive, int x, int y) {
    if (!yk_r1_mask_at(S, maskActive, x, y)) return 0u;
    const int cx = x >> 2, cy = y >> 2;
    const unsigned r0 = S.cellMask[(size_t)cy * S.nbx + (cx >> 4)], r1 = S.cellMask[(size_t)(cy + 1) * S.nbx + (cx >> 4)];
    return (~(((r0 >> (cx & 15)) & 3u) | (((r1 >> (cx & 15)) & 3u) << 2))) & 15u;
}

// The walk of LeftRightOrder (framework.h:228-256) over the constraint box of a plane of pw x ph samples: rows of
// nbw = ceil(cw / 8) blocks while y < cy + ch, then one more block at the start of the next row when that row still
// lies inside the plane (HasNextBlock tests `y < h` after the wrap).  Block sizes: the reference compares with the
// constraint's width / height, not its right / bottom edge, and falls back to x % 8 (framework.h:251-252) - 0 for the
// full-resolution planes, possibly 4 for a reduced chroma plane whose box starts on an odd multiple of 4.
struct YkR1Geom { int cx, cy, cw, ch, nbw, nBlocks, maskActive, pad; };
struct YkR1Block { int x, y, rw, rh; };
static __device__ __forceinline__ YkR1Block yk_r1_block(const YkR1Geom& G, int i) {
    YkR1Block b;
    b.x = G.cx + 8 * (i % G.nbw); b.y = G.cy + 8 * (i / G.nbw);
    b.rw = (b.x + 8 > G.cw) ? (b.x & 7) : 8;
    b.rh = (b.y + 8 > G.ch) ? (b.y & 7) : 8;
    return b;
}

// float(d) / float(o), correctly rounded, for the small integers of this stage: with r = RN(1 / o) the residual
// correction below reproduces IEEE division for all 1 <= d, o <= 2048 (checked exhaustively on the host);
// anything larger takes the division itself.
static __device__ __forceinline__ float yk_r1_quot(int d, float fo, float r) {
    const float fd = __int2float_rn(d);
    const float q0 = __fmul_rn(fd, r);
    return __fmaf_rn(__fmaf_rn(-q0, fo, fd), r, q0);                // r == 0 (a dropped term) gives exactly +0
}

struct YkR1Item { unsigned xy, pairs, nibOff, pad; };        // one listed block: see yk_k_r1_offsets

static __device__ YkR1Geom yk_r1_geometry(const YkSlotDev& S, const YkR1Launch& LP) {
    const YkR1Args& A0 = LP.job[0];
    YkR1Geom G;
    G.cx = A0.cx; G.cy = A0.cy; G.cw = A0.cw; G.ch = A0.ch; G.nbw = A0.nbw; G.nBlocks = A0.nBlocks;
    G.maskActive = S.alphaValid && !S.alphaReset; G.pad = 0;
    if (LP.fromHdr) {
        // the bound box of the alpha stage straight from the image header the analysis kernel has just written
        // (MipPrefilter, EC.cpp:1287-1291, 1400-1403; CheckMipmapMask's full image without an alpha stage)
        int L = 0, T = 0, R = S.w, B = S.h, any = 1;
        G.maskActive = 0;
        if (LP.useAlpha && S.nPlanes == 4) {
            const int* h = S.hdr;
            any = h[YK_HD_ALPHA_KEPT0] > 0;
            if (any) {
                L = S.w - h[YK_HD_ALPHA_MINX]; T = INT_MAX / 2 - h[YK_HD_ALPHA_MINY]; R = h[YK_HD_ALPHA_MAXX]; B = h[YK_HD_ALPHA_MAXY];
                G.maskActive = !(L == 0 && T == 0 && R == S.w && B == S.imgH);
            }
        }
        G.cx = (L >> 3) << 3; G.cy = (T >> 3) << 3;                                   // EC.cpp:4386-4391
        G.cw = (((R + 7) >> 3) << 3) - G.cx; G.ch = (((B + 7) >> 3) << 3) - G.cy;
        G.nbw = (G.cw + 7) >> 3;
        const int rows = (G.ch + 7) >> 3;
        G.nBlocks = any ? G.nbw * rows : 0;
        if (G.nBlocks > 0 && G.cy + 8 * rows < A0.ph) G.nBlocks += 1;
    }
    return G;
}

__global__ void __launch_bounds__(YK_R1_UNIT)
yk_k_r1_offsets(const YkSlotDev* __restrict__ slots, int slot, const YkR1Launch LP) {
    __shared__ int sWarp[33];
    __shared__ int sUnit;
    __shared__ unsigned sBase[2];
    __shared__ YkR1Geom sG;
    const YkSlotDev& S = slots[slot];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const YkR1Args& A0 = LP.job[0];
    const bool reduced = (A0.shX | A0.shY) != 0;
    unsigned long long* status = S.r1Status;
    if (tid == 0) {
        const YkR1Geom G = yk_r1_geometry(S, LP);
        sG = G;
        const int nU = (G.nBlocks + YK_R1_UNIT - 1) / YK_R1_UNIT;
        sUnit = (int)atomicAdd(reinterpret_cast<unsigned*>(&status[nU]), 1u);
    }
    __syncthreads();
    const YkR1Geom G = sG;
    const bool maskActive = G.maskActive != 0;
    const int nUnits = (G.nBlocks + YK_R1_UNIT - 1) / YK_R1_UNIT;
    const int u = sUnit;
    if (u >= nUnits) {                                            // launched for the largest possible box
        if (nUnits == 0 && u == 0 && tid == 0) {
            for (int j = 0; j < LP.nJobs; j++) { S.hdr[YK_HD_R1_NIB0 + LP.job[j].out] = 0; S.hdr[YK_HD_R1_DEF0 + LP.job[j].out] = 0; }
        }
        return;
    }
    // ---- count: which pixel pairs of my block are coded (EC.cpp:826-861; the 2x2 / 2x1 / 1x2 mask samples a reduced
    // pixel covers lie in one 16x16 alpha tile, so "all of them" is the top-left one)
    const int i = u * YK_R1_UNIT + tid;
    unsigned pairs = 0;
    YkR1Block b = { 0, 0, 0, 0 };
    if (i < G.nBlocks) {
        b = yk_r1_block(G, i);
        if (!reduced) {
            if (b.rw == 8 && b.rh == 8 && b.x + 8 <= S.w && b.y + 8 <= S.h) {
                const unsigned cells = yk_r1_cells(S, maskActive, b.x, b.y);
                pairs = ((cells & 1u) ? 0x00003333u : 0u) | ((cells & 2u) ? 0x0000CCCCu : 0u) | ((cells & 4u) ? 0x33330000u : 0u) | ((cells & 8u) ? 0xCCCC0000u : 0u);
            }
        } else {
            for (int r = 0; r < b.rh; r++)
                for (int p = 0; 2 * p < b.rw; p++) {
                    const int fx = (b.x + 2 * p) << A0.shX, fy = (b.y + r) << A0.shY;
                    if (yk_r1_mask_at(S, maskActive, fx, fy) && !yk_r1_smooth_at(S, fx, fy)) pairs |= 1u << (4 * r + p);
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
            if (u == nUnits - 1)
                for (int j = 0; j < LP.nJobs; j++) {
                    S.hdr[YK_HD_R1_NIB0 + LP.job[j].out] = (int)(base >> 32) + (tot & 0xFFFF);
                    S.hdr[YK_HD_R1_DEF0 + LP.job[j].out] = (int)(unsigned)base + (tot >> 16);
                }
        }
    }
    __syncthreads();
    if (n > 0) {
        YkR1Item it;
        it.xy = (unsigned)b.x | ((unsigned)b.y << 16); it.pairs = pairs; it.nibOff = sBase[0] + (unsigned)(ex & 0xFFFF); it.pad = (unsigned)b.rw | ((unsigned)b.rh << 8);
        reinterpret_cast<uint4*>(S.r1List)[sBase[1] + (unsigned)(ex >> 16)] = make_uint4(it.xy, it.pairs, it.nibOff, it.pad);
    }
}

#ifndef YK_R1_MINB
#define YK_R1_MINB 9
#endif
#ifndef YK_R1_CAP
#define YK_R1_CAP 12
#endif
__global__ void __launch_bounds__(YK_R1_THREADS, YK_R1_MINB)
yk_k_r1_encode(const YkSlotDev* __restrict__ slots, int slot, const YkR1Launch LP, const int* __restrict__ lut, const unsigned short* __restrict__ rtab, const short* __restrict__ r7tab) {
    __shared__ __align__(16) float sTerm[YK_R1_THREADS / 32][6][64];
    const YkSlotDev& S = slots[slot];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const YkR1Args& A0 = LP.job[0];
    const bool reduced = (A0.shX | A0.shY) != 0;
    const bool maskActive = reduced ? (yk_r1_geometry(S, LP).maskActive != 0) : false;      // only the reduced planes look at the masks again
    const int nList = __ldg(&S.hdr[YK_HD_R1_DEF0 + A0.out]);                                // written by yk_k_r1_offsets
    const int r = lane >> 2, c0 = (lane & 3) * 2;
    const int startMode = A0.mode3 ? 3 : 0;
    const int nItems = nList * LP.nJobs;
    float (*term)[64] = sTerm[warp];
    // the next item's list entry and (full-resolution planes) its samples are fetched while the current item is coded
    const int stride = gridDim.x * (YK_R1_THREADS / 32);
    int it = blockIdx.x * (YK_R1_THREADS / 32) + warp;
    uint4 nItem = make_uint4(0u, 0u, 0u, 0u);
    int nv0 = 0, nv1 = 0;
    auto fetch = [&](int at) {
        if (at >= nItems) return;
        const int kk = LP.nJobs == 3 ? at / 3 : at / LP.nJobs;
        const YkR1Args& An = LP.job[at - kk * LP.nJobs];
        nItem = __ldg(reinterpret_cast<const uint4*>(S.r1List) + kk);
        nv0 = 0; nv1 = 0;
        if (!reduced && ((nItem.y >> lane) & 1u)) {
            const int xx = nItem.x & 0xFFFF, yy = nItem.x >> 16;
            if (An.srcU8) {
                const unsigned short two = __ldg(reinterpret_cast<const unsigned short*>(An.srcU8 + (size_t)(yy + r) * An.pitchU8 + xx + c0));
                nv0 = two & 255; nv1 = two >> 8;
            } else {
                const int2 p = __ldg(reinterpret_cast<const int2*>(An.src + (size_t)(yy + r) * An.pw + xx + c0));
                nv0 = p.x; nv1 = p.y;
            }
        }
    };
    fetch(it);
    for (; it < nItems; it += stride) {
        const int k = LP.nJobs == 3 ? it / 3 : it / LP.nJobs, j = it - k * LP.nJobs;
        const YkR1Args& A = LP.job[j];
        const uint4 item = nItem;
        int v0 = nv0, v1 = nv1;
        fetch(it + stride);
        const int x = item.x & 0xFFFF, y = item.x >> 16;
        const int brw = item.w & 255, brh = (item.w >> 8) & 255;
        const bool valid = (item.y >> lane) & 1u;
        // ---- min / max of the block (Plane::GetMinMax_Y, Plane.cpp:489-587).  Full resolution: over the coded pixels.
        // Reduced planes: over "not smooth and any covered mask sample set", where the reference addresses the
        // full-size mask with the REDUCED plane's width as row stride (Plane.cpp:516, 538-553) - restated literally.
        bool m0 = valid, m1 = valid;
        if (reduced) {
            m0 = m1 = false;
            if (r < brh && c0 < brw) {
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const size_t vi = ((size_t)(x + c0 + q) << A.shX) + (size_t)((y + r) << A.shY) * A.pw;
                    const int W = S.w;
                    // every covered sample is a mipmapMask value: kept by the alpha stage AND not claimed by a gradient
                    // tile (FittingQuadSmooth zeroes the mask over accepted tiles, EC.cpp:4035); with the reduced width as
                    // row stride the samples of the second row lie in another cell than vi
                    bool any = yk_r1_valid_at(S, maskActive, vi);
                    if (A.shX) any |= yk_r1_valid_at(S, maskActive, vi + 1);
                    if (A.shY) any |= yk_r1_valid_at(S, maskActive, vi + A.pw);
                    if (A.shX && A.shY) any |= yk_r1_valid_at(S, maskActive, vi + A.pw + 1);
                    const bool ok = any && !yk_r1_smooth_at(S, (int)(vi % W), (int)(vi / W));
                    if (q) m1 = ok; else m0 = ok;
                }
            }
            if (valid || m0 || m1) {
                const int2 p = __ldg(reinterpret_cast<const int2*>(A.src + (size_t)(y + r) * A.pw + x + c0));
                v0 = p.x; v1 = p.y;
            }
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
        (void)BN;
        const int r7 = __ldg(r7tab + b6 * 224 + (max(diff, 32) - 32));          // ((max(diff,32) - 32) * 127 + scale - 1) / scale, scale = 223 - BN (EC.cpp:643-650)
        const int* T = lut + (size_t)(b6 * 176 + min(r7, 175)) * YK_R1_LUT_INTS;
        const unsigned short* RT = rtab + (size_t)(b6 * YK_R1_RTAB_R7 + min(r7, YK_R1_RTAB_R7 - 1)) * (6 * YK_R1_RTAB_VALUES);
        const int o0 = v0 + sgn, o1 = v1 + sgn;
        // the per-value table covers range codes below 128, bases below 63 (the other LUTs have entries above 255: reference
        // quirks) and samples 0..383; anything else takes the search and the division themselves
        const bool tabled = r7 < YK_R1_RTAB_R7 && b6 < 63 && __all_sync(YK_FULL, (unsigned)o0 < (unsigned)YK_R1_RTAB_VALUES && (unsigned)o1 < (unsigned)YK_R1_RTAB_VALUES);
        // cumulated relative error term, float32 (EC.cpp:884-886): float(|entry - value|) / float(value) summed in
        // row-major valid-pixel order.  Invalid pixels and zero samples add +0.0f, which leaves the sum unchanged: their
        // reciprocal is set to zero here, and a zero difference gives a zero quotient by itself.
        const bool use0 = valid && o0 != 0, use1 = valid && o1 != 0;
        const float f0 = __int2float_rn(use0 ? o0 : 1), f1 = __int2float_rn(use1 ? o1 : 1);
        const float rc0 = use0 ? __frcp_rn(f0) : 0.0f, rc1 = use1 ? __frcp_rn(f1) : 0.0f;
        if (tabled) {
#pragma unroll
            for (int mode = 0; mode < 6; mode++) {
                if (mode < startMode) continue;
                const int L0 = __ldg(RT + mode * YK_R1_RTAB_VALUES + o0) & 255, L1 = __ldg(RT + mode * YK_R1_RTAB_VALUES + o1) & 255;
                *reinterpret_cast<float2*>(&term[mode][2 * lane]) = make_float2(yk_r1_quot(abs(L0 - o0), f0, rc0), yk_r1_quot(abs(L1 - o1), f1, rc1));
            }
        } else {
#pragma unroll 1
            for (int mode = startMode; mode < 6; mode++) {
                const int off = mode < 3 ? 16 * mode : 48 + 8 * (mode - 3);
                int g0 = 0, g1 = 0;
                for (int st = (mode < 3 ? 8 : 4); st > 0; st >>= 1) {
                    if (o0 > __ldg(T + 72 + off + g0 + st)) g0 += st;
                    if (o1 > __ldg(T + 72 + off + g1 + st)) g1 += st;
                }
                const int d0 = abs(__ldg(T + off + g0) - o0), d1 = abs(__ldg(T + off + g1) - o1);
                term[mode][2 * lane] = (use0 && d0 != 0) ? __fdiv_rn(__int2float_rn(d0), f0) : 0.0f;
                term[mode][2 * lane + 1] = (use1 && d1 != 0) ? __fdiv_rn(__int2float_rn(d1), f1) : 0.0f;
            }
        }
        __syncwarp();
        float err = 0.0f;
        if (lane >= startMode && lane < 6) {
            // the reference sums in row-major valid-pixel order; float addition is not associative, so one lane per mode
            const float4* P = reinterpret_cast<const float4*>(term[lane]);
#pragma unroll 4
            for (int q = 0; q < 16; q++) {
                const float4 t = P[q];
                err = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(err, t.x), t.y), t.z), t.w);
            }
        }
        // `<=`: later modes win ties (EC.cpp:897-905).  Sums can be negative on signed chroma planes: compare through the
        // order-preserving map of the bit patterns (-0 counted as +0, as `<=` does)
        const unsigned eraw = __float_as_uint(__fadd_rn(err, 0.0f));
        const unsigned ebits = (lane >= startMode && lane < 6) ? (eraw ^ ((eraw >> 31) ? 0xFFFFFFFFu : 0x80000000u)) : 0xFFFFFFFFu;
        const unsigned emin = __reduce_min_sync(YK_FULL, ebits);
        const int bestMode = 31 - __clz((int)__ballot_sync(YK_FULL, ebits == emin && lane < 6));
        __syncwarp();                                                           // the terms are rewritten by the next item
        const unsigned bv0 = __ballot_sync(YK_FULL, valid);
        if (valid) {
            // the codes of the chosen mode: from the table again, or the largest n with value > threshold[n]
            const int boff = bestMode < 3 ? 16 * bestMode : 48 + 8 * (bestMode - 3);
            int c0v = 0, c1v = 0;
            if (tabled) {
                c0v = __ldg(RT + bestMode * YK_R1_RTAB_VALUES + o0) >> 8; c1v = __ldg(RT + bestMode * YK_R1_RTAB_VALUES + o1) >> 8;
            } else {
                for (int st = (bestMode < 3 ? 8 : 4); st > 0; st >>= 1) {
                    if (o0 > __ldg(T + 72 + boff + c0v + st)) c0v += st;
                    if (o1 > __ldg(T + 72 + boff + c1v + st)) c1v += st;
                }
            }
            const int before = 2 * __popc(bv0 & ((1u << lane) - 1u));           // both pixels of a lane share validity
            const int n0 = (int)item.z + before;                                // even: the two nibbles share a byte
            reinterpret_cast<uint8_t*>(S.r1Nib[A.out])[n0 >> 1] = (uint8_t)(c0v | (c1v << 4));   // low nibble first, EC.cpp:1180-1184
            if (A.dst) {
                // EC.cpp:4441-4502: chroma blocks with a negative minimum go back minus 128; reduced planes are
                // written at their top-left full-size position only ("interpolation comes later")
                const int off = boff;
                const int offset = (A.chroma && sgn) ? -128 : 0;
                int* d = A.dst + (size_t)((y + r) << A.shY) * S.w + ((x + c0) << A.shX);
                d[0] = __ldg(T + off + c0v) + offset; d[1 << A.shX] = __ldg(T + off + c1v) + offset;
            }
        }
        if (lane == 0) S.r1Defs[A.out][k] = (uint16_t)((bestMode << 13) | (r7 << 7) | b6);    // EncodeTileType, YAIK_private.h:358
    }
}

// The three full-resolution colour planes of one launch (nJobs == 3, the fused yk_analyze path and yk_range_dyn on R, G,
// B): a warp takes a listed block and codes its three planes in one go.  What a block costs besides its per-pixel work is
// shared by the planes (list entry, geometry, validity ballots), and the ordered error sums - 64 dependent float adds
// that one lane per mode has to do whatever the other lanes are up to - run for the 18 (plane, mode) pairs side by side
// in 18 lanes; the 16-pixel quadrants of the block that hold no coded pixel only add +0.0f and are skipped (exact).
#ifndef YK_R1_MINB3
#define YK_R1_MINB3 8
#endif
#ifndef YK_R1_CAP3
#define YK_R1_CAP3 8
#endif
#ifndef YK_R1_NO3
#define YK_R1_NO3 0
#endif
__global__ void __launch_bounds__(YK_R1_THREADS, YK_R1_MINB3)
yk_k_r1_encode3(const YkSlotDev* __restrict__ slots, int slot, const YkR1Launch LP, const int* __restrict__ lut, const unsigned short* __restrict__ rtab, const short* __restrict__ r7tab) {
    __shared__ __align__(16) float sTerm[YK_R1_THREADS / 32][3][6][64];
    const YkSlotDev& S = slots[slot];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const YkR1Args& A0 = LP.job[0];
    const int nList = __ldg(&S.hdr[YK_HD_R1_DEF0 + A0.out]);                                // written by yk_k_r1_offsets
    const int r = lane >> 2, c0 = (lane & 3) * 2;
    const int startMode = A0.mode3 ? 3 : 0;
    float (*term)[6][64] = sTerm[warp];
    // the lane that sums (plane sj, mode sm)
    const int sj = lane / 6, sm = lane - 6 * sj;
    const bool summing = lane < 18 && sm >= startMode;
    const int stride = gridDim.x * (YK_R1_THREADS / 32);
    int k = blockIdx.x * (YK_R1_THREADS / 32) + warp;
    uint4 nItem = make_uint4(0u, 0u, 0u, 0u);
    int nv[3][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 } };
    auto fetch = [&](int at) {
        if (at >= nList) return;
        nItem = __ldg(reinterpret_cast<const uint4*>(S.r1List) + at);
        const bool on = (nItem.y >> lane) & 1u;
        const int xx = nItem.x & 0xFFFF, yy = nItem.x >> 16;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const YkR1Args& An = LP.job[j];
            nv[j][0] = 0; nv[j][1] = 0;
            if (on) {
                if (An.srcU8) {
                    const unsigned short two = __ldg(reinterpret_cast<const unsigned short*>(An.srcU8 + (size_t)(yy + r) * An.pitchU8 + xx + c0));
                    nv[j][0] = two & 255; nv[j][1] = two >> 8;
                } else {
                    const int2 p = __ldg(reinterpret_cast<const int2*>(An.src + (size_t)(yy + r) * An.pw + xx + c0));
                    nv[j][0] = p.x; nv[j][1] = p.y;
                }
            }
        }
    };
    fetch(k);
    for (; k < nList; k += stride) {
        const uint4 item = nItem;
        int o[3][2];
#pragma unroll
        for (int j = 0; j < 3; j++) { o[j][0] = nv[j][0]; o[j][1] = nv[j][1]; }
        fetch(k + stride);
        const int x = item.x & 0xFFFF, y = item.x >> 16;
        const bool valid = (item.y >> lane) & 1u;
        int meta[3];                                                           // b6 | r7 << 8 | sgn << 16 | tabled << 24
#pragma unroll
        for (int j = 0; j < 3; j++) {
            // ---- min / max over the coded pixels (Plane::GetMinMax_Y, Plane.cpp:489-587), table index (EC.cpp:635-650, 764-768)
            const int v0 = o[j][0], v1 = o[j][1];
            int mn = __reduce_min_sync(YK_FULL, valid ? min(v0, v1) : INT_MAX);
            int mx = __reduce_max_sync(YK_FULL, valid ? max(v0, v1) : INT_MIN);
            if (mn == INT_MAX) { mn = 0; mx = 0; }
            int sgn = 0;
            if (mn < 0) { mn += 128; mx += 128; sgn = 128; }
            mn = min(max(mn, 0), 255); mx = min(max(mx, mn), 255);
            const int m = min(mn, 224);
            const int diff = max(mx - m, 16);
            const int b6 = (m * 63 + 112) / 224;
            const int r7 = __ldg(r7tab + b6 * 224 + (max(diff, 32) - 32));
            const int o0 = v0 + sgn, o1 = v1 + sgn;
            o[j][0] = o0; o[j][1] = o1;
            const bool tabled = r7 < YK_R1_RTAB_R7 && b6 < 63 && __all_sync(YK_FULL, (unsigned)o0 < (unsigned)YK_R1_RTAB_VALUES && (unsigned)o1 < (unsigned)YK_R1_RTAB_VALUES);
            meta[j] = b6 | (r7 << 8) | (sgn << 16) | ((int)tabled << 24);
            // relative error terms, float32 (EC.cpp:884-886); invalid pixels and zero samples add +0.0f
            const bool use0 = valid && o0 != 0, use1 = valid && o1 != 0;
            const float f0 = __int2float_rn(use0 ? o0 : 1), f1 = __int2float_rn(use1 ? o1 : 1);
            const float rc0 = use0 ? __frcp_rn(f0) : 0.0f, rc1 = use1 ? __frcp_rn(f1) : 0.0f;
            if (tabled) {
                const unsigned short* RT = rtab + (size_t)(b6 * YK_R1_RTAB_R7 + r7) * (6 * YK_R1_RTAB_VALUES);
#pragma unroll
                for (int mode = 0; mode < 6; mode++) {
                    if (mode < startMode) continue;
                    const int L0 = __ldg(RT + mode * YK_R1_RTAB_VALUES + o0) & 255, L1 = __ldg(RT + mode * YK_R1_RTAB_VALUES + o1) & 255;
                    *reinterpret_cast<float2*>(&term[j][mode][2 * lane]) = make_float2(yk_r1_quot(abs(L0 - o0), f0, rc0), yk_r1_quot(abs(L1 - o1), f1, rc1));
                }
            } else {
                const int* T = lut + (size_t)(b6 * 176 + min(r7, 175)) * YK_R1_LUT_INTS;
#pragma unroll 1
                for (int mode = startMode; mode < 6; mode++) {
                    const int off = mode < 3 ? 16 * mode : 48 + 8 * (mode - 3);
                    int g0 = 0, g1 = 0;
                    for (int st = (mode < 3 ? 8 : 4); st > 0; st >>= 1) {
                        if (o0 > __ldg(T + 72 + off + g0 + st)) g0 += st;
                        if (o1 > __ldg(T + 72 + off + g1 + st)) g1 += st;
                    }
                    const int d0 = abs(__ldg(T + off + g0) - o0), d1 = abs(__ldg(T + off + g1) - o1);
                    term[j][mode][2 * lane] = (use0 && d0 != 0) ? __fdiv_rn(__int2float_rn(d0), f0) : 0.0f;
                    term[j][mode][2 * lane + 1] = (use1 && d1 != 0) ? __fdiv_rn(__int2float_rn(d1), f1) : 0.0f;
                }
            }
        }
        __syncwarp();
        // ---- ordered sums: row-major pixel order, one lane per (plane, mode).  float4 group q = pixels 4q .. 4q + 3 = row
        // q >> 1, left (q even) or right half: a quadrant without coded pixels holds zeros only
        float err = 0.0f;
        if (summing) {
            const float4* P = reinterpret_cast<const float4*>(term[sj][sm]);
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const unsigned hm = (item.y >> (16 * half)) & 0xFFFFu;
                if (hm) {
                    const bool Lq = (hm & 0x3333u) != 0u, Rq = (hm & 0xCCCCu) != 0u;
                    if (Lq && Rq) {
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const float4 t = P[8 * half + q];
                            err = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(err, t.x), t.y), t.z), t.w);
                        }
                    } else {
                        const float4* Q = P + 8 * half + (Rq ? 1 : 0);
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const float4 t = Q[2 * q];
                            err = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(err, t.x), t.y), t.z), t.w);
                        }
                    }
                }
            }
        }
        // `<=`: later modes win ties (EC.cpp:897-905); order-preserving map of the bit patterns (-0 counted as +0)
        const unsigned eraw = __float_as_uint(__fadd_rn(err, 0.0f));
        const unsigned ebits = summing ? (eraw ^ ((eraw >> 31) ? 0xFFFFFFFFu : 0x80000000u)) : 0xFFFFFFFFu;
        int best[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const bool mine = summing && sj == j;
            const unsigned emin = __reduce_min_sync(YK_FULL, mine ? ebits : 0xFFFFFFFFu);
            best[j] = 31 - __clz((int)__ballot_sync(YK_FULL, mine && ebits == emin)) - 6 * j;
        }
        __syncwarp();                                                           // the terms are rewritten by the next block
        const unsigned bv0 = item.y;                                            // coded pairs == lanes with pixels
        const int before = 2 * __popc(bv0 & ((1u << lane) - 1u));
        const int n0 = (int)item.z + before;                                    // even: the two nibbles share a byte
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const YkR1Args& A = LP.job[j];
            const int b6 = meta[j] & 255, r7 = (meta[j] >> 8) & 255, sgn = (meta[j] >> 16) & 255;
            const bool tabled = (meta[j] >> 24) & 1;
            const int bestMode = best[j];
            if (valid) {
                const int boff = bestMode < 3 ? 16 * bestMode : 48 + 8 * (bestMode - 3);
                const int* T = lut + (size_t)(b6 * 176 + min(r7, 175)) * YK_R1_LUT_INTS;
                const int o0 = o[j][0], o1 = o[j][1];
                int c0v = 0, c1v = 0;
                if (tabled) {
                    const unsigned short* RT = rtab + (size_t)(b6 * YK_R1_RTAB_R7 + r7) * (6 * YK_R1_RTAB_VALUES);
                    c0v = __ldg(RT + bestMode * YK_R1_RTAB_VALUES + o0) >> 8; c1v = __ldg(RT + bestMode * YK_R1_RTAB_VALUES + o1) >> 8;
                } else {
                    for (int st = (bestMode < 3 ? 8 : 4); st > 0; st >>= 1) {
                        if (o0 > __ldg(T + 72 + boff + c0v + st)) c0v += st;
                        if (o1 > __ldg(T + 72 + boff + c1v + st)) c1v += st;
                    }
                }
                reinterpret_cast<uint8_t*>(S.r1Nib[A.out])[n0 >> 1] = (uint8_t)(c0v | (c1v << 4));   // low nibble first, EC.cpp:1180-1184
                if (A.dst) {
                    const int offset = (A.chroma && sgn) ? -128 : 0;
                    int* d = A.dst + (size_t)(y + r) * S.w + (x + c0);
                    d[0] = __ldg(T + boff + c0v) + offset; d[1] = __ldg(T + boff + c1v) + offset;
                }
            }
            if (lane == 0) S.r1Defs[A.out][k] = (uint16_t)((bestMode << 13) | (r7 << 7) | b6);    // EncodeTileType, YAIK_private.h:358
        }
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
void yk_launch_range_dyn_encode(const YkSlotDev* slotsDev, int slot, const YkR1Launch& launch, int maxBlocks, int numSMs, const int* lutDev, const uint16_t* rtabDev, const int16_t* r7Dev, cudaStream_t st) {
    const int units = (maxBlocks + YK_R1_UNIT - 1) / YK_R1_UNIT;
    YK_LAUNCH(yk_k_r1_offsets, dim3(units > 0 ? units : 1), dim3(YK_R1_UNIT), 0, st, slotsDev, slot, launch);
    // warps take items with a grid stride: enough CTAs to fill every SM, no more than there can be items
    const int perCta = YK_R1_THREADS / 32;
    long long want = ((long long)maxBlocks * launch.nJobs + perCta - 1) / perCta;
    const long long cap = (long long)numSMs * YK_R1_CAP;
    if (want > cap) want = cap;
    const bool three = launch.nJobs == 3 && !(launch.job[0].shX | launch.job[0].shY) && !YK_R1_NO3;
    if (three) {
        // a warp codes the three planes of a block: one third of the items
        want = ((long long)maxBlocks + perCta - 1) / perCta;
        const long long cap3 = (long long)numSMs * YK_R1_CAP3;
        if (want > cap3) want = cap3;
        YK_LAUNCH(yk_k_r1_encode3, dim3(want > 0 ? (unsigned)want : 1u), dim3(YK_R1_THREADS), 0, st, slotsDev, slot, launch, lutDev, reinterpret_cast<const unsigned short*>(rtabDev), reinterpret_cast<const short*>(r7Dev));
        return;
    }
    YK_LAUNCH(yk_k_r1_encode, dim3(want > 0 ? (unsigned)want : 1u), dim3(YK_R1_THREADS), 0, st, slotsDev, slot, launch, lutDev, reinterpret_cast<const unsigned short*>(rtabDev), reinterpret_cast<const short*>(r7Dev));
}
void yk_launch_chroma(const YkSlotDev* slotsDev, int slot, int w, int h, const YkChromaArgs& args, cudaStream_t st) {
    const int quads = (w >> 1) * (h >> 1);
    YK_LAUNCH(yk_k_chroma, dim3((quads + 255) / 256), dim3(256), 0, st, slotsDev, slot, args);
}

// CUDA loads kernels lazily, and loading one synchronises the context: a kernel that waits for another stream (yk_strip_run)
// would then never be released by a kernel that is launched for the first time.  Every kernel is loaded up front.
int yk_preload_aux() {
#ifndef YK_EMULATE
    cudaFuncAttributes fa;
    { const cudaError_t e = cudaFuncGetAttributes(&fa, yk_k_state); if (e != cudaSuccess) return (int)e; }
    { const cudaError_t e = cudaFuncGetAttributes(&fa, yk_k_r1_offsets); if (e != cudaSuccess) return (int)e; }
    { const cudaError_t e = cudaFuncGetAttributes(&fa, yk_k_r1_encode); if (e != cudaSuccess) return (int)e; }
    { const cudaError_t e = cudaFuncGetAttributes(&fa, yk_k_r1_encode3); if (e != cudaSuccess) return (int)e; }
    { const cudaError_t e = cudaFuncGetAttributes(&fa, yk_k_chroma); if (e != cudaSuccess) return (int)e; }
#endif
    return 0;
}
