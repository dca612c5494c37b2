// Internal declarations shared by the kernels (yk_kernels.cu) and the C-ABI layer (yk_api.cu).
// Vocabulary follows the reference (KLab/YAIK): planes, tiles, swizzle blocks, passes, streams.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifndef YK_EMULATE
#include <cuda_runtime.h>
#define YK_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#define YK_NPASS 7
#define YK_REGION 64            // the largest swizzle block (YAIK_private.h:212-276): regions tile the image, strips start on them
#define YK_THREADS 256          // CTA size of the auxiliary kernels (state download, DynamicTileEncode)

// Swizzle geometry of pass id p (HeaderGradientTile::getSwizzleSize, include/YAIK_private.h:212-276),
// in Convert()'s order (EC.cpp:9057-9093): 16x16 16x8 8x16 8x8 8x4 4x8 4x4.
struct YkPassGeom { int shx, shy, bw, bh, bits; };
#define YK_PASS_TABLE { {4,4,64,64,16}, {4,3,64,64,32}, {3,4,64,64,32}, {3,3,64,64,64}, {3,2,64,32,64}, {2,3,32,64,64}, {2,2,32,32,64} }

// per-pass counters written by the analysis kernels (ints)
enum { YK_ST_TILEDONE = 0, YK_ST_MINX, YK_ST_MINY, YK_ST_MAXX, YK_ST_MAXY, YK_ST_RGBBYTES, YK_ST_PAD0, YK_ST_PAD1, YK_ST_STRIDE };
// per-slot header ints (device memory, zeroed/initialised before each run)
enum {
    YK_HD_ERR = 0,                 // bit0: sample outside 0..255
    YK_HD_ALPHA_KEPT0,             // alpha stage: kept tiles, then their box (count first, like a pass's counters)
    YK_HD_ALPHA_MINX, YK_HD_ALPHA_MINY, YK_HD_ALPHA_MAXX, YK_HD_ALPHA_MAXY,
    YK_HD_R2_CHUNKS, YK_HD_R2_TILES,        // totals from the scan: 16-byte chunks per plane, coded tiles per plane
    YK_HD_R1_NIB0, YK_HD_R1_NIB1, YK_HD_R1_NIB2, YK_HD_R1_DEF0, YK_HD_R1_DEF1, YK_HD_R1_DEF2,
    YK_HD_TICKET_EMIT, YK_HD_TICKET_ANALYZE, // work tickets: emission groups in stream order / regions of the persistent analysis kernel
    YK_HD_PASS0 = 16,              // YK_NPASS * YK_ST_STRIDE ints
    YK_HD_INTS = YK_HD_PASS0 + YK_NPASS * YK_ST_STRIDE
};

// TMA descriptor of one int32 plane as a 2-D tensor [h][w] (a CUtensorMap: 128 bytes, 64-byte aligned), encoded on
// the host by cuTensorMapEncodeTiled.  Box = 132 x 17 samples for the colour planes (one macro-tile row of two
// neighbouring 64x64 regions plus its right / bottom corner samples: the TMA unit's cost is per box row, so rows are
// made long), 128 x 16 for alpha; out-of-image samples arrive as zeros and are never used unclamped.
struct alignas(64) YkTmap { unsigned long long opaque[16]; };
#ifndef YK_EMIT_THREADS
#define YK_EMIT_THREADS 512    // threads of a yk_k_emit CTA == nibble words / range tiles per look-back group (host and device agree on it)
#endif
#define YK_UNIT_W 128           // pixels a unit of the analysis kernel is wide (two regions)
#define YK_RAW_PITCH 132        // samples per row of a staged colour box
#define YK_RAW_ROWS 17

// Device-visible description of one slot (one image + all results of its analysis).
#define YK_U8_BOX 144           // bytes per row of a staged colour box of a packed (u8) plane: 128 + corner column, 16-byte granular

struct alignas(128) YkSlotDev {
    YkTmap tmap[4];             // per plane (R, G, B, alpha): int32 planes, or the packed u8 planes when isU8
    const int32_t* plane[4];    // int32 row-major planes, pitch == w (Plane::GetPixels(), framework.h:81); with isU8 they are
                                //   only filled on demand (yk_k_expand) for the kernels outside the hot path
    const uint8_t* planeU8[4];  // packed upload (yk_set_image): one byte per sample, row pitch pitchU8
    int isU8, pitchU8;
    const void* rowBelow[3];    // strip mode: the pixel row under the strip (w samples, int32 or u8 like the planes), else NULL
    int w, h, nPlanes;
    int nbx, nby;               // 64x64 regions
    int imgH, y0;               // strip mode: height of the whole image / first row of this strip (else h, 0)
    int hasAbove, hasBelow;     // strip mode: there is a strip above / below this one
    const uint32_t* touchInTop;     // strip mode: touch words of the boundary lattice row as the strip above / below saw them
    const uint32_t* touchInBottom;  //   (latW words each, written by the neighbour over NVLink P2P between the two phases)
    int latW, latH;             // lattice of 4-pixel points: w/4+1, h/4+1
    // ---- compact state (replaces the reference's int32 state planes, EncoderContext.h:300-323)
    uint16_t* cellMask;         // [h/4][nbx]   bit i = 4x4 cell (16*bx+i) claimed   == smoothMap / mapSmoothTile != 0
    uint32_t* touchMap;         // [latH][latW] per lattice point: bit 4*rp+k = in pass position rp of the running launch an
                                //   accepted tile has this point as corner k (0 TL,1 TR,2 BL,3 BR); bit 31 = claimed by an
                                //   earlier launch.  != 0  ==  mappedRGB != 0
    uint8_t*  alphaKept;        // [ceil(h/16)][ceil(w/16)] 1 = tile has a non-zero alpha sample
    int*      hdr;              // YK_HD_INTS ints
    // ---- gradient results
    uint8_t*  bitmap[YK_NPASS];     // pFillBitMap in the reference's swizzled layout
    uint32_t* emitStatus[YK_NPASS]; // per swizzle block: decoupled look-back word (rgb bytes << 2 | flag)
    uint8_t*  rgb[YK_NPASS];        // rgbStream
    uint8_t*  latRGB;               // [latH][latW][3]  CompressF(Round6(clamped pixel),250) at every lattice point
    uint32_t* emitNib[YK_NPASS];    // per tile in stream order 4 bits: which of TL,TR,BL,BR the tile emits (8 tiles per word)
    // ---- range stage R2 (DynamicTileCompressor)
    unsigned long long* r2Status;   // per group of YK_EMIT_THREADS tiles (row-major tile order): look-back word (chunks << 32 | codedTiles << 2 | flag)
    uint8_t*  r2Raw[3];         // [h/8][w/8][64] index bytes of every coded tile at a fixed place (written by the analysis kernel)
    uint32_t* r2RawType[3];     // [h/8][w/8] color0 | minCol << 8 | delta << 16
    uint8_t*  r2Idx[3];         // the streams in the reference's order (gathered by yk_k_emit)
    uint8_t*  r2Type[3];
    // ---- range stage R1 (DynamicTileEncode)
    unsigned long long* r1Status;   // look-back words of yk_k_r1_offsets' units (256 blocks of the walk each) + the unit ticket; zeroed per call
    void*     r1List;           // the blocks that hold valid pixels, in walk order (16 bytes each, see yk_k_r1_offsets)
    uint32_t* r1Nib[3];         // nibble stream (two nibbles per byte, written pairwise)
    uint16_t* r1Defs[3];
    int       alphaReset;       // set by the host after the alpha stage: bbox == full image -> mask all 255 (EC.cpp:1400-1403)
    int       alphaValid;       // alpha stage has been run for this image
};

struct YkRun {
    int nPasses;
    int passId[YK_NPASS];       // which passes this launch runs, in order
    int rejectFactor;
    int doAlpha;
    int doR2;                   // code the DynamicTileCompressor tiles of every region after its cascade
    int endgameUnits;           // units per CTA at the end of a launch that are taken on demand instead of ahead (load balance of the tail)
    int fresh;                  // no cell is claimed yet (first gradient launch after yk_reset_state): cellMask need not be read
};

#ifdef __cplusplus
extern "C++" {
#endif
// launch wrappers (yk_kernels.cu); `slots` is a device array, grid.y indexes it from slot0
int  yk_analyze_setup(int* numSMs);
int  yk_preload_analyze();              // force-load every kernel of a translation unit (see yk_analyze.cu); return a cudaError_t
int  yk_preload_emit();
int  yk_preload_aux();      // opt-in shared memory of the persistent kernel; returns a cudaError_t
void yk_launch_analyze(const YkSlotDev* slotsDev, int slot0, int nSlots, int nRegions, int gridCtas, bool packedU8, const YkRun& run, cudaStream_t st);   // nRegions: of one image (nbx * nby)
void yk_launch_expand(const YkSlotDev* slotsDev, int slot, int nPlanes, int w, int h, int32_t* const* dst, cudaStream_t st);
void yk_launch_fold_touch(const YkSlotDev* slotsDev, int slot0, int nSlots, int nWords, cudaStream_t st);
// one-thread kernels on a strip's stream: publish / wait for an epoch number in (peer-mapped) halo memory
void yk_launch_flag_set(unsigned* a, unsigned* b, unsigned value, cudaStream_t st);
void yk_launch_flag_wait(const unsigned* a, const unsigned* b, unsigned value, cudaStream_t st);
void yk_launch_owner(const YkSlotDev* slotsDev, int slot0, int nSlots, int nPoints, const YkRun& run, cudaStream_t st);
void yk_launch_emit(const YkSlotDev* slotsDev, int slot0, int nSlots, int gradGroups, int r2Groups, const YkRun& run, cudaStream_t st);
void yk_launch_state(const YkSlotDev* slotsDev, int slot, int nRegions, int32_t* smoothMap, int32_t* mipmapMask,
                     int32_t* mappedRGB, int32_t* recon0, int32_t* recon1, int32_t* recon2, cudaStream_t st);
// DynamicTileEncode on one plane: the source samples, the (possibly halved) constraint box and the walk over it
struct YkR1Args {
    const int32_t* src; int pw, ph;      // plane coded: pw x ph int32 samples on the device
    const uint8_t* srcU8; int pitchU8;   // ... or, when not NULL, one byte per sample (a colour plane uploaded packed)
    int32_t* dst;                        // optional write-back plane (w x h int32, full size), else NULL
    int shX, shY;                        // 1 = the plane is SampleDown'ed on that axis (isHalfX / isHalfY)
    int chroma, mode3;                   // isCo | isCg; mode3BitOnly
    int cx, cy, cw, ch;                  // constraint box in the plane's own coordinates (EC.cpp:4386-4401)
    int nbw, nBlocks;                    // blocks per row of the walk; all blocks (rows * nbw + the extra one, see yk_r1_block)
    int out;                             // index of r1Nib / r1Defs / header counters written
};
struct YkChromaArgs { int32_t *y, *co, *cg; int half[4]; int mode[2]; };
// one launch: 1..3 planes with the same geometry and validity (job[0]'s); fromHdr: the constraint box is derived on the
// device from the alpha stage's box in the image header (the fused yk_analyze path, no host round trip)
struct YkR1Launch { int nJobs, fromHdr, useAlpha, pad; YkR1Args job[3]; };
void yk_launch_range_dyn_encode(const YkSlotDev* slotsDev, int slot, const YkR1Launch& launch, int maxBlocks, int numSMs, const int* lutDev, const uint16_t* rtabDev, const int16_t* r7Dev, cudaStream_t st);
void yk_launch_chroma(const YkSlotDev* slotsDev, int slot, int w, int h, const YkChromaArgs& args, cudaStream_t st);
#ifdef __cplusplus
}
#endif
