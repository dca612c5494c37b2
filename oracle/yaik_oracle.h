/* TEST INFRASTRUCTURE ONLY — the CPU oracle of the YAIK encoder-analysis hot path.
 *
 * A plain-C restatement of what the reference (KLab/YAIK, /root/reference) computes in
 *   MipPrefilter / quadRecursion      encoder/EncoderContext.cpp:1257-1427, 357-430
 *   FittingQuadSmooth                 encoder/EncoderContext.cpp:3710-4363
 *   DynamicTileCompressor (R2)        encoder/EncoderContext.cpp:8335-8522
 *   DynamicTileEncode (R1)            encoder/EncoderContext.cpp:4365-4602, 517-906, Plane.cpp:489-587
 * Pinned against the compiled reference itself (oracle/_ref, tests/test_oracle_vs_ref.py) and against
 * golden vectors generated from it (tests/golden/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the product never does.
 */
#ifndef YAIK_ORACLE_H
#define YAIK_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct yko_ctx yko_ctx;

/* planes[c] is w*h int32 row-major (Plane::GetPixels(), framework.h:81); nplanes is 3 or 4. Copies. */
yko_ctx* yko_create(int w, int h, int nplanes, const int32_t* const* planes);
void     yko_destroy(yko_ctx* c);

/* MipPrefilter(true).  bitmap needs ceil((w/16)*(h/16)/8) bytes.  Returns 0, or -1 outside the parity
 * domain (w == h == 2^k >= 16, 4 planes). */
int yko_alpha_reject(yko_ctx* c, uint8_t* bitmap, int* bitmapBytes, int boundPx[4],
                     int* remainingPixels, int* wroteChunk, int chunkBBoxTiles[4]);

/* One FittingQuadSmooth(rejectFactor, R, G, B, output, false, shX, shY) call.
 * bitmap: getBitmapSwizzleSize()/8 bytes (zero-filled here); rgb: 3*(w/tsx+1)*(h/tsy+1) bytes max.
 * bbox = {minX, minY, maxX, maxY} of accepted tiles ({w,h,0,0} when none). */
int yko_gradient_pass(yko_ctx* c, int rejectFactor, int shX, int shY, uint8_t* bitmap, int* bitmapBytes,
                      uint8_t* rgb, int* rgbBytes, int bbox[4], int* tileDone);

/* DynamicTileCompressor(stream, src=plane, map=mapSmoothTile[plane], debug).  idx: <= w*h bytes,
 * type: <= 3*(w/8)*(h/8) bytes, debug: w*h int32 or NULL (only written where a pixel is coded). */
int yko_range1d(yko_ctx* c, int plane, uint8_t* idx, int* idxBytes, uint8_t* type, int* typeBytes, int32_t* debug);

/* DynamicTileEncode(mode3BitOnly, plane, dst, false,false,false,false) (full resolution).
 * nibbles: packed low-nibble-first, (w/8)*(h/8)*32 bytes max, *nNibbles = number of 4-bit codes;
 * defs: (w/8)*(h/8) u16 max; dst: w*h int32 or NULL (written only at valid pixels);
 * constraint = {x,y,w,h} of the 8-aligned bound box. */
int yko_range_dyn(yko_ctx* c, int plane, int mode3BitOnly, uint8_t* nibbles, int* nNibbles,
                  uint16_t* defs, int* nDefs, int32_t* dst, int constraint[4]);

/* Chroma front-end: Image::ConvertToRGB2YCoCg(true) (Image.cpp:285-321, RGBtoYCoCg EC.cpp:53-67),
 * Plane::SampleDown (Plane.cpp:278-369; mode = EDownSample 0 NEAREST_TL, 1 NEAREST_BR, 2 AVERAGE_BOX, 3 MAX_BOX,
 * 4 MIN_BOX; -1 for the combinations whose result reads outside the plane in the reference), and DynamicTileEncode on
 * any plane: the Y plane, or a chroma plane reduced by SampleDown (isChroma = isCo | isCg, halfX, halfY). */
void yko_rgb_to_ycocg(yko_ctx* c, int32_t* oY, int32_t* oCo, int32_t* oCg);
int  yko_sample_down(const int32_t* src, int w, int h, int halfX, int halfY, int mode, int32_t* dst);
int  yko_range_dyn_plane(yko_ctx* c, const int32_t* src, int pw, int ph, int mode3BitOnly, int isChroma, int halfX, int halfY,
                         uint8_t* nibbles, int* nNibbles, uint16_t* defs, int* nDefs, int32_t* dst, int constraint[4]);

/* State planes for comparison with the reference's: 0 smoothMap, 1 mipmapMask, 2..4 mapSmoothTile[c]
 * (w*h), 5..7 mappedRGB[c] ((w+1)*(h+1)), 8..10 recon/testOutput[c] (w*h). */
const int32_t* yko_state_plane(yko_ctx* c, int which);

/* The LUT of one (min,max) table: mode 0..5 -> count entries (16,16,16,8,8,8); returns count and
 * fills base6/range7 (buildTable, EC.cpp:628-697). */
int yko_dyn_table(int minV, int maxV, int mode, int lut[16], int* base6, int* range7);

#ifdef __cplusplus
}
#endif
#endif
