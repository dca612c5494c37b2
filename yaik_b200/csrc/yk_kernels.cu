// yaik_b200 — auxiliary sm_100a kernels of the YAIK encoder-analysis stage (the hot path is yk_analyze.cu + yk_emit.cu):
//
//   yk_k_state        expands the compact masks into the reference's int32 state planes (compat download).
//   yk_k_r1_*         DynamicTileEncode (EC.cpp:4365-4503, 747-1212), LUT search at 3/4 bits per pixel.
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

// ------------------------------------------------------------------------------------------------------------------// launch wrappers
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
