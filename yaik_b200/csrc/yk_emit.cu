// yaik_b200 — corner ownership, rgbStream emission and the gather of the range stage's streams (sm_100a).
//
// The reference walks tiles in stream order and lets a tile emit a corner colour only if no earlier tile (of this or an
// earlier pass) touched that lattice point (mappedRGB, EC.cpp:4001-4021, 4115-4132).  Order-free: a lattice point is
// emitted in the first pass that touches it, by the accepted toucher with the smallest stream position — all of which
// the point's touch word says (written by yk_k_analyze).
//
//   yk_k_owner   one thread per lattice point: resolves the owner and sets the owner tile's bit in emitNib
//                (4 bits per tile in stream order: it emits TL / TR / BL / BR).
//   yk_k_emit    one thread per 8 tiles of a pass, groups of 8 x YK_EMIT_THREADS tiles taken in stream order by ticket: byte counts from
//                the nibbles (popc), block scan + one decoupled look-back per group, then the colours are copied from
//                latRGB.  Tickets past the gradient groups gather DynamicTileCompressor's per-tile output (written at
//                fixed places by yk_k_analyze) into the reference's row-major tile order (EC.cpp:8412-8413), offsets by
//                the same scan + look-back.
#include "yk_device.h"

// Strips (one large image split into tile-row strips over several GPUs): the lattice row on a strip
// boundary is touched by tiles of both strips.  Each strip ORs in the words its neighbour computed (touchInTop /
// touchInBottom); a tile of the upper strip always precedes a tile of the lower strip in a pass's stream, and only the
// strip that holds the owner tile marks it.
#ifndef YK_EMIT_SLIM
#define YK_EMIT_SLIM 0
#endif
#ifndef YK_OWNER_THREADS
#define YK_OWNER_THREADS 256
#endif
#if YK_EMIT_SLIM
__global__ void __launch_bounds__(YK_OWNER_THREADS, 16)
#else
__global__ void __launch_bounds__(YK_OWNER_THREADS)
#endif
yk_k_owner(const YkSlotDev* __restrict__ slots, int slot0, int nPoints, YkRun run) {
    const YkSlotDev& S = slots[slot0 + blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nPoints) return;
    uint32_t word = __ldg(&S.touchMap[idx]);
    const int gy = idx / S.latW, gx = idx - gy * S.latW;
    if (S.hasAbove && gy == 0) word |= __ldg(&S.touchInTop[gx]);
    if (S.hasBelow && gy == S.latH - 1) word |= __ldg(&S.touchInBottom[gx]);
    if (word == 0u || (word & 0x80000000u)) return;                 // untouched / claimed by an earlier launch
    const int rp = (__ffs((int)(word & 0x0FFFFFFFu)) - 1) >> 2;     // first pass of this launch that touches the point
    const unsigned nib = (word >> (4 * rp)) & 15u;                  // roles present in that pass
    const int pid = run.passId[rp];
    const YkGeomS g = yk_geom_s(pid);
    const int nSwzX = (S.w + (1 << g.lbw) - 1) >> g.lbw;
    const int LX = (4 * gx) >> g.shx, LY = (4 * gy) >> g.shy;       // the point in tile units of that pass
    const int tileRows = S.h >> g.shy;
    int best = INT_MAX, bestK = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if ((nib >> k) & 1u) {                                      // an accepted tile has the point as its corner k
            const int ty = LY - (k >> 1);
            int pos;
            if (ty < 0) pos = k - 5;                                // a tile of the strip above: before every tile of this strip (BR's tile before BL's)
            else if (ty >= tileRows) pos = INT_MAX - 1;             // a tile of the strip below: after every tile of this strip
            else pos = yk_pos_s(g, nSwzX, LX - (k & 1), ty);
            if (pos < best) { best = pos; bestK = k; }
        }
    }
    if (best < 0 || best == INT_MAX - 1) return;                    // the neighbouring strip holds the owner
    atomicOr(&S.emitNib[pid][best >> 3], 1u << (4 * (best & 7) + bestK));
}

// 512 threads at 64 registers (two CTAs per SM): the kernel is a chain of dependent global accesses per CTA (ticket, masks,
// per-tile output, look-back), and in the pipelined step its CTAs hold SMs that the next analysis launch is waiting for.
// Per pipelined step: 256 threads x 3 CTAs per SM at 78 registers 47.6 us, 256 x 4 at 64 registers 45.8, 512 x 2 44.7
// (half the tickets and look-backs), 1024 x 1 the same with a longer lone launch, 512 x 3 at 42 registers (spills) 46.0.
#ifndef YK_EMIT_MINB
#define YK_EMIT_MINB 2
#endif
__global__ void __launch_bounds__(YK_EMIT_THREADS, YK_EMIT_MINB)
yk_k_emit(const YkSlotDev* __restrict__ slots, int slot0, YkRun run, int gradGroups, int r2Groups) {
    __shared__ int sTicket;
    __shared__ unsigned sA[YK_EMIT_THREADS / 32 + 1], sB[YK_EMIT_THREADS / 32 + 1];
    const YkSlotDev& S = slots[slot0 + blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = YK_EMIT_THREADS / 32;
    const int w = S.w, h = S.h;
    if (tid == 0) sTicket = atomicAdd(&S.hdr[YK_HD_TICKET_EMIT], 1);
    __syncthreads();
    const int ticket = sTicket;
    if (ticket >= gradGroups + r2Groups) return;

    if (ticket >= gradGroups) {
        // ---- DynamicTileCompressor streams: tiles in row-major order, 16 bytes per coded quadrant, 3 type bytes per coded tile
        const int grp = ticket - gradGroups, nbx = S.nbx, tilesW = w >> 3, nTiles = tilesW * (h >> 3);
        const int t = grp * YK_EMIT_THREADS + tid;
        unsigned chunks = 0;
        if (t < nTiles) {
            const int ty = t / tilesW, tx = t - ty * tilesW;
            const uint32_t r0 = S.cellMask[(size_t)(2 * ty) * nbx + (tx >> 3)], r1 = S.cellMask[(size_t)(2 * ty + 1) * nbx + (tx >> 3)];
            const int s2 = 2 * (tx & 7);
            chunks = 4u - (unsigned)__popc(((r0 >> s2) & 3u) | (((r1 >> s2) & 3u) << 2));       // quadrants whose top-left map pixel is 0 (EC.cpp:8420-8430)
        }
        const unsigned tiles = chunks > 0;
#if !YK_EMIT_SLIM
        // the tile's output is fetched while the offsets are being scanned
        uint4 v[3][4];
        uint32_t ty3[3] = { 0, 0, 0 };
        if (chunks) {
#pragma unroll
            for (int p = 0; p < 3; p++) {
                const uint4* src = reinterpret_cast<const uint4*>(S.r2Raw[p] + (size_t)t * 64);
#pragma unroll
                for (int k = 0; k < 4; k++) if (k < (int)chunks) v[p][k] = src[k];
                ty3[p] = S.r2RawType[p][t];
            }
        }
#endif
        unsigned ic = chunks, it = tiles;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned a0 = __shfl_up_sync(YK_FULL, ic, d), a1 = __shfl_up_sync(YK_FULL, it, d);
            if (lane >= d) { ic += a0; it += a1; }
        }
        if (lane == 31) { sA[warp] = ic; sB[warp] = it; }
        __syncthreads();
        if (warp == 0) {
            const unsigned a = lane < NW ? sA[lane] : 0u, b2 = lane < NW ? sB[lane] : 0u;
            unsigned ia = a, ib = b2;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned a0 = __shfl_up_sync(YK_FULL, ia, d), a1 = __shfl_up_sync(YK_FULL, ib, d);
                if (lane >= d) { ia += a0; ib += a1; }
            }
            const unsigned totC = __shfl_sync(YK_FULL, ia, 31), totT = __shfl_sync(YK_FULL, ib, 31);
            const unsigned long long base = yk_lookback64(S.r2Status, grp, totC, totT);
            const unsigned bc = (unsigned)(base >> 32), bt = (unsigned)(base & 0xFFFFFFFFull);
            if (lane < NW) { sA[lane] = bc + ia - a; sB[lane] = bt + ib - b2; }
            if (grp == r2Groups - 1 && lane == 0) { S.hdr[YK_HD_R2_CHUNKS] = (int)(bc + totC); S.hdr[YK_HD_R2_TILES] = (int)(bt + totT); }
        }
        __syncthreads();
        if (chunks) {
            const unsigned chunkOff = sA[warp] + ic - chunks, tileOff = sB[warp] + it - tiles;
#if YK_EMIT_SLIM
            // the slim build (32 registers, 128 threads: a CTA fits on an SM next to a resident analysis CTA) copies after the
            // scan instead of holding the tile's 12 x 16 bytes in registers across it
#pragma unroll 1
            for (int p = 0; p < 3; p++) {
                const uint4* src = reinterpret_cast<const uint4*>(S.r2Raw[p] + (size_t)t * 64);
                uint4* dst = reinterpret_cast<uint4*>(S.r2Idx[p] + (size_t)chunkOff * 16);
                for (int k = 0; k < (int)chunks; k++) dst[k] = src[k];
                const uint32_t ty = S.r2RawType[p][t];
                uint8_t* td = S.r2Type[p] + (size_t)tileOff * 3;                                // EC.cpp:8503-8505
                td[0] = (uint8_t)ty; td[1] = (uint8_t)(ty >> 8); td[2] = (uint8_t)(ty >> 16);
            }
#else
#pragma unroll
            for (int p = 0; p < 3; p++) {
                uint4* dst = reinterpret_cast<uint4*>(S.r2Idx[p] + (size_t)chunkOff * 16);
#pragma unroll
                for (int k = 0; k < 4; k++) if (k < (int)chunks) dst[k] = v[p][k];
                uint8_t* td = S.r2Type[p] + (size_t)tileOff * 3;                                // EC.cpp:8503-8505
                td[0] = (uint8_t)ty3[p]; td[1] = (uint8_t)(ty3[p] >> 8); td[2] = (uint8_t)(ty3[p] >> 16);
            }
#endif
        }
        return;
    }

    // ---- rgbStream of one pass: ticket -> (pass position, group of 8 x YK_EMIT_THREADS tiles)
    int rp = 0, grp = ticket, nWords = 0, nSwzX = 0, nGroups = 0;
    YkGeomS g = yk_geom_s(run.passId[0]);
    for (;;) {
        g = yk_geom_s(run.passId[rp]);
        nSwzX = (w + (1 << g.lbw) - 1) >> g.lbw;
        nWords = (nSwzX * ((h + (1 << g.lbh) - 1) >> g.lbh) * g.bits) >> 3;
        nGroups = (nWords + YK_EMIT_THREADS - 1) / YK_EMIT_THREADS;
        if (grp < nGroups) break;
        grp -= nGroups; rp++;
    }
    const int pid = run.passId[rp];
    const int wi = grp * YK_EMIT_THREADS + tid;
    uint32_t word = wi < nWords ? __ldg(&S.emitNib[pid][wi]) : 0u;
    const unsigned cnt = 3u * (unsigned)__popc(word);
    unsigned inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned a0 = __shfl_up_sync(YK_FULL, inc, d); if (lane >= d) inc += a0; }
    if (lane == 31) sA[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned a = lane < NW ? sA[lane] : 0u;
        unsigned ia = a;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned a0 = __shfl_up_sync(YK_FULL, ia, d); if (lane >= d) ia += a0; }
        const unsigned tot = __shfl_sync(YK_FULL, ia, 31);
        const unsigned base = yk_lookback32(S.emitStatus[pid], grp, tot);
        if (lane < NW) sA[lane] = base + ia - a;
        if (grp == nGroups - 1 && lane == 0) S.hdr[YK_HD_PASS0 + pid * YK_ST_STRIDE + YK_ST_RGBBYTES] = (int)(base + tot);
    }
    __syncthreads();
    if (word == 0u) return;
    unsigned off = sA[warp] + inc - cnt;
    // the 8 tiles of a word share a swizzle block: invert the stream position once
    const int lbits = g.bits == 16 ? 4 : (g.bits == 32 ? 5 : 6);
    const int pos0 = wi * 8, blk = pos0 >> lbits, within0 = pos0 & (g.bits - 1);
    const int tprShift = g.lbw - g.shx;
    const int bys = blk / nSwzX, bxs = blk - bys * nSwzX;
    uint8_t* out = S.rgb[pid];
    while (word) {
        const int b = __ffs((int)word) - 1;
        word &= word - 1u;
        const int within = within0 + (b >> 2), k = b & 3;                         // TL, TR, BL, BR (EC.cpp:4115-4132)
        const int gtx = (bxs << tprShift) + (within & ((1 << tprShift) - 1)), gty = (bys << (g.lbh - g.shy)) + (within >> tprShift);
        const int gx = ((gtx + (k & 1)) << g.shx) >> 2, gy = ((gty + (k >> 1)) << g.shy) >> 2;
        const uint8_t* sp = S.latRGB + ((size_t)gy * S.latW + gx) * 3;
        out[off] = sp[0]; out[off + 1] = sp[1]; out[off + 2] = sp[2];
        off += 3;
    }
}

void yk_launch_owner(const YkSlotDev* slotsDev, int slot0, int nSlots, int nPoints, const YkRun& run, cudaStream_t st) {
    YK_LAUNCH(yk_k_owner, dim3((nPoints + YK_OWNER_THREADS - 1) / YK_OWNER_THREADS, nSlots), dim3(YK_OWNER_THREADS), 0, st, slotsDev, slot0, nPoints, run);
}
void yk_launch_emit(const YkSlotDev* slotsDev, int slot0, int nSlots, int gradGroups, int r2Groups, const YkRun& run, cudaStream_t st) {
    YK_LAUNCH(yk_k_emit, dim3(gradGroups + r2Groups, nSlots), dim3(YK_EMIT_THREADS), 0, st, slotsDev, slot0, run, gradGroups, r2Groups);
}

// Strips (yk_strip_run): epoch flags in halo memory.  The setter runs after the copies it announces (stream order) and
// makes them visible system-wide before the flag; the waiter spins on its own GPU's memory.
__global__ void yk_k_flag_set(unsigned* a, unsigned* b, unsigned value) {
    __threadfence_system();
    if (a) yk_stv(a, value);
    if (b) yk_stv(b, value);
}
__global__ void yk_k_flag_wait(const unsigned* a, const unsigned* b, unsigned value) {
    while (a && (int)(yk_ldv(a) - value) < 0) yk_spin();
    while (b && (int)(yk_ldv(b) - value) < 0) yk_spin();
    __threadfence_system();
}
void yk_launch_flag_set(unsigned* a, unsigned* b, unsigned value, cudaStream_t st) { YK_LAUNCH(yk_k_flag_set, dim3(1), dim3(1), 0, st, a, b, value); }
void yk_launch_flag_wait(const unsigned* a, const unsigned* b, unsigned value, cudaStream_t st) { YK_LAUNCH(yk_k_flag_wait, dim3(1), dim3(1), 0, st, a, b, value); }

// CUDA loads kernels lazily, and loading one synchronises the context: a kernel that waits for another stream (yk_strip_run)
// would then never be released by a kernel that is launched for the first time.  Every kernel is loaded up front.
int yk_preload_emit() {
#ifndef YK_EMULATE
    cudaFuncAttributes fa;
    { const cudaError_t e = cudaFuncGetAttributes(&fa, yk_k_owner); if (e != cudaSuccess) return (int)e; }
    { const cudaError_t e = cudaFuncGetAttributes(&fa, yk_k_emit); if (e != cudaSuccess) return (int)e; }
    { const cudaError_t e = cudaFuncGetAttributes(&fa, yk_k_flag_set); if (e != cudaSuccess) return (int)e; }
    { const cudaError_t e = cudaFuncGetAttributes(&fa, yk_k_flag_wait); if (e != cudaSuccess) return (int)e; }
#endif
    return 0;
}
