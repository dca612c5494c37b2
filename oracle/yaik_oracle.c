/* TEST INFRASTRUCTURE ONLY — see yaik_oracle.h.  Plain-C restatement of the reference's encoder-analysis
 * hot path in sequential, reference-order form (the CUDA product uses an order-free form; the two must
 * agree bit for bit).  Every function cites the reference lines it follows (paths under /root/reference,
 * "EC.cpp" = encoder/EncoderContext.cpp).  Pinned by tests/test_oracle_vs_ref.py against the compiled,
 * unmodified reference (oracle/_ref) and by tests/golden/ vectors generated from it.
 */
#include "yaik_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

struct yko_ctx {
    int w, h, nplanes;
    int32_t* src[4];
    /* EncoderContext state, EncoderContext.h:300-323 */
    int32_t* smoothMap;
    int32_t* mipmapMask;
    int32_t* mapSmoothTile[3];
    int32_t* mappedRGB[3];      /* (w+1)*(h+1) */
    int32_t* recon[3];          /* `output` image handed to FittingQuadSmooth, EC.cpp:9057 */
    int boundX0, boundY0, boundX1, boundY1;
    int remainingPixels, mipMapTileSize;
};

static int32_t* plane_new(size_t n, int32_t v) {
    int32_t* p = (int32_t*)malloc(n * sizeof(int32_t) + 16);
    if (p) for (size_t i = 0; i < n; i++) p[i] = v;
    return p;
}

/* Plane::GetPixelValue, encoder/framework.h:116-121 (clamping accessor) */
static int pix_clamped(const int32_t* p, int w, int h, int x, int y) {
    if (x >= w) x = w - 1;
    if (x < 0) x = 0;
    if (y >= h) y = h - 1;
    if (y < 0) y = 0;
    return p[x + (size_t)y * w];
}

yko_ctx* yko_create(int w, int h, int nplanes, const int32_t* const* planes) {
    if (w <= 0 || h <= 0 || nplanes < 3 || nplanes > 4) return NULL;
    yko_ctx* c = (yko_ctx*)calloc(1, sizeof *c);
    if (!c) return NULL;
    c->w = w; c->h = h; c->nplanes = nplanes;
    size_t n = (size_t)w * h, n1 = (size_t)(w + 1) * (h + 1);
    for (int k = 0; k < nplanes; k++) {
        c->src[k] = plane_new(n, 0);
        memcpy(c->src[k], planes[k], n * sizeof(int32_t));
    }
    /* CheckMipmapMask, EC.cpp:2784-2794: all-255 mask, bound = full image */
    c->mipmapMask = plane_new(n, 255);
    c->boundX0 = 0; c->boundY0 = 0; c->boundX1 = w; c->boundY1 = h;
    c->remainingPixels = w * h; c->mipMapTileSize = 16;
    /* lazily created on the first FittingQuadSmooth call, EC.cpp:3739-3749; created here, all zero */
    c->smoothMap = plane_new(n, 0);
    for (int k = 0; k < 3; k++) {
        c->mapSmoothTile[k] = plane_new(n, 0);
        c->mappedRGB[k] = plane_new(n1, 0);
        c->recon[k] = plane_new(n, 0);
    }
    return c;
}

void yko_destroy(yko_ctx* c) {
    if (!c) return;
    for (int k = 0; k < 4; k++) free(c->src[k]);
    free(c->smoothMap); free(c->mipmapMask);
    for (int k = 0; k < 3; k++) { free(c->mapSmoothTile[k]); free(c->mappedRGB[k]); free(c->recon[k]); }
    free(c);
}

const int32_t* yko_state_plane(yko_ctx* c, int which) {
    if (which == 0) return c->smoothMap;
    if (which == 1) return c->mipmapMask;
    if (which >= 2 && which <= 4) return c->mapSmoothTile[which - 2];
    if (which >= 5 && which <= 7) return c->mappedRGB[which - 5];
    if (which >= 8 && which <= 10) return c->recon[which - 8];
    return NULL;
}

/* ------------------------------------------------------------------------------------------------
 * Alpha-zero tile rejection: MipPrefilter (EC.cpp:1257-1427) + quadRecursion (EC.cpp:357-430).
 * On the parity domain (w == h == 2^k >= 16) the recursion is: a node is "all zero" iff all alpha
 * samples under it are 0; all-zero nodes of size >= 16 zero the mask (EC.cpp:394-414), and a 16x16 node
 * that is not all-zero grows the bounding box (EC.cpp:416-422).  Restated per 16x16 tile. */
int yko_alpha_reject(yko_ctx* c, uint8_t* bitmap, int* bitmapBytes, int boundPx[4],
                     int* remainingPixels, int* wroteChunk, int chunkBBoxTiles[4]) {
    int w = c->w, h = c->h;
    if (c->nplanes != 4 || w != h || w < 16 || (w & (w - 1))) return -1;
    const int32_t* a = c->src[3];
    int tw = w / 16, th = h / 16;
    int L = 9999999, T = 9999999, R = -1, B = -1;               /* EC.cpp:1280-1283 */
    for (size_t i = 0; i < (size_t)w * h; i++) c->mipmapMask[i] = 255;      /* EC.cpp:1270-1273 */
    for (int ty = 0; ty < th; ty++)
        for (int tx = 0; tx < tw; tx++) {
            int allZero = 1;
            for (int y = 0; y < 16 && allZero; y++)
                for (int x = 0; x < 16; x++)
                    if (a[(tx * 16 + x) + (size_t)(ty * 16 + y) * w] != 0) { allZero = 0; break; }   /* EC.cpp:428 */
            if (allZero) {
                for (int y = 0; y < 16; y++)
                    for (int x = 0; x < 16; x++) c->mipmapMask[(tx * 16 + x) + (size_t)(ty * 16 + y) * w] = 0;
            } else {
                if (L > tx * 16) L = tx * 16;
                if (T > ty * 16) T = ty * 16;
                if (R < tx * 16 + 16) R = tx * 16 + 16;
                if (B < ty * 16 + 16) B = ty * 16 + 16;
            }
        }
    if (R < 0) return -1;       /* fully transparent: the reference runs into negative sizes (SURVEY hazard 7) */
    c->boundX0 = L; c->boundX1 = R; c->boundY0 = T; c->boundY1 = B; c->mipMapTileSize = 16;     /* EC.cpp:1287-1291 */
    *wroteChunk = 0; *bitmapBytes = 0;
    if (L != 0 || T != 0 || R != w || B != h) {                 /* EC.cpp:1294 */
        int bx0 = L >> 4, bx1 = R >> 4, by0 = T >> 4, by1 = B >> 4;
        int tWB = bx1 - bx0, tHB = by1 - by0;
        int sizeByte = (tWB * tHB + 7) / 8;                     /* EC.cpp:1306 */
        memset(bitmap, 0, (size_t)sizeByte);
        int bitPos = 0; c->remainingPixels = 0;
        for (int y = 0; y < tHB; y++)
            for (int x = 0; x < tWB; x++) {                     /* EC.cpp:1317-1327 */
                int v = c->mipmapMask[(x + bx0) * 16 + (size_t)((y + by0) * 16) * w];
                if (v != 0) { bitmap[bitPos >> 3] |= (uint8_t)(1 << (bitPos & 7)); c->remainingPixels += 256; }
                bitPos++;
            }
        *wroteChunk = 1; *bitmapBytes = sizeByte;
        chunkBBoxTiles[0] = bx0; chunkBBoxTiles[1] = by0; chunkBBoxTiles[2] = tWB; chunkBBoxTiles[3] = tHB;  /* EC.cpp:1387-1390 */
    } else {
        for (size_t i = 0; i < (size_t)w * h; i++) c->mipmapMask[i] = 255;      /* EC.cpp:1401: rejection discarded */
        c->remainingPixels = R * B;                                             /* EC.cpp:1402 */
    }
    boundPx[0] = c->boundX0; boundPx[1] = c->boundY0; boundPx[2] = c->boundX1; boundPx[3] = c->boundY1;
    *remainingPixels = c->remainingPixels;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Gradient pass: FittingQuadSmooth (EC.cpp:3710-4363), RGB (PlaneBit 7), useYCoCg = false. */
static int Round6(int v)  { int r = v >> 2; return (r << 2) | (r >> 4); }                  /* EC.cpp:3183-3189 */
static int Round6P(int v) { v++; if (v > 255) v = 255; int r = v >> 2; return (r << 2) | (r >> 4); }  /* EC.cpp:3202-3207 */
static int CompressF(int v, int rate) { return (v * rate + 127) / 255; }                   /* EC.cpp:3191-3194 */

/* HeaderGradientTile::getSwizzleSize, include/YAIK_private.h:212-276 */
static int swizzle_size(int shX, int shY, int* bw, int* bh, int* bits) {
    int sx = 0, sy = 0;
    if (shX == 4 && (shY == 4 || shY == 3)) { sx = 64; sy = 64; }
    else if (shX == 3 && (shY == 4 || shY == 3)) { sx = 64; sy = 64; }
    else if (shX == 3 && shY == 2) { sx = 64; sy = 32; }
    else if (shX == 2 && shY == 3) { sx = 32; sy = 64; }
    else if (shX == 2 && shY == 2) { sx = 32; sy = 32; }
    *bw = sx; *bh = sy; *bits = (sx >> shX) * (sy >> shY);
    return sx != 0;
}

int yko_gradient_pass(yko_ctx* c, int rejectFactor, int shX, int shY, uint8_t* bitmap, int* bitmapBytes,
                      uint8_t* rgb, int* rgbBytes, int bbox[4], int* tileDone) {
    int W = c->w, H = c->h;
    int tsx = 1 << shX, tsy = 1 << shY;
    int bw, bh, bits;
    if (!swizzle_size(shX, shY, &bw, &bh, &bits)) return -1;
    static const int weight4[]  = { 1024, 768, 512, 256, 0 };                                       /* EC.cpp:3735-3737 */
    static const int weight8[]  = { 1024, 896, 768, 640, 512, 384, 256, 128, 0 };
    static const int weight16[] = { 1024, 960, 896, 832, 768, 704, 640, 576, 512, 448, 384, 320, 256, 192, 128, 64, 0 };
    const int* wX = tsx == 4 ? weight4 : tsx == 8 ? weight8 : weight16;
    const int* wY = tsy == 4 ? weight4 : tsy == 8 ? weight8 : weight16;

    int xBB = (W + bw - 1) / bw, yBB = (H + bh - 1) / bh;       /* EC.cpp:3770-3771 */
    int sizeBitmap = (xBB * yBB * bits) >> 3;                   /* EC.cpp:3775 */
    memset(bitmap, 0, (size_t)sizeBitmap);
    *bitmapBytes = sizeBitmap;
    uint8_t* wr = rgb;
    int done = 0;
    int minX = W, maxX = 0, minY = H, maxY = 0;                 /* EC.cpp:3798-3799 */
    int stepYS = bits * xBB, stepXS = bits, stepY = bw / tsx;   /* EC.cpp:3803-3805 */
    int recT[3][256];

    int posYS = 0;
    for (int sy = 0; sy < H; sy += bh, posYS += stepYS) {
        int posXS = posYS;
        for (int sx = 0; sx < W; sx += bw, posXS += stepXS) {
            int posY = posXS;
            for (int y = sy; y < sy + bh; y += tsy, posY += stepY) {
                if (y >= H || y + tsy > H) break;               /* EC.cpp:3818 */
                int pos = posY;
                for (int x = sx; x < sx + bw; x += tsx, pos++) {
                    if (x >= W || x + tsx > W) break;           /* EC.cpp:3826 */
                    int cr[4][3], c6[4][3], c6p[4][3];          /* TL TR BL BR */
                    for (int n = 0; n < 3; n++) {               /* EC.cpp:3845-3868 */
                        cr[0][n] = pix_clamped(c->src[n], W, H, x, y);
                        cr[1][n] = pix_clamped(c->src[n], W, H, x + tsx, y);
                        cr[2][n] = pix_clamped(c->src[n], W, H, x, y + tsy);
                        cr[3][n] = pix_clamped(c->src[n], W, H, x + tsx, y + tsy);
                        for (int k = 0; k < 4; k++) { c6[k][n] = Round6(cr[k][n]); c6p[k][n] = Round6P(cr[k][n]); }
                    }
                    size_t i0 = x + (size_t)y * W;
                    if (c->mapSmoothTile[0][i0] || c->mapSmoothTile[1][i0] || c->mapSmoothTile[2][i0]) continue;  /* EC.cpp:3871-3875 */
                    int rej = 0, rej6 = 0, rejO = 0, rej6O = 0, rej6OE = 0, rej6E = 0;
                    for (int dy = 0; dy < tsy; dy++) {
                        int tF = wY[dy], bF = 1024 - tF;
                        for (int dx = 0; dx < tsx; dx++) {
                            int lF = wX[dx], rF = 1024 - lF;
                            const int rounding = (1 << 19) - 1;
                            for (int ch = 0; ch < 3; ch++) {    /* EC.cpp:3936-3991 */
                                int cur = c->src[ch][(x + dx) + (size_t)(y + dy) * W];
                                int bT = cr[0][ch] * lF + cr[1][ch] * rF, bB = cr[2][ch] * lF + cr[3][ch] * rF;
                                int num = bT * tF + bB * bF;
                                int blendC = (num + rounding) / (1024 * 1024), blendCO = num / (1024 * 1024);
                                bT = c6[0][ch] * lF + c6[1][ch] * rF; bB = c6[2][ch] * lF + c6[3][ch] * rF;
                                num = bT * tF + bB * bF;
                                int blendC6 = (num + rounding) / (1024 * 1024), blendC6O = num / (1024 * 1024);
                                bT = c6p[0][ch] * lF + c6p[1][ch] * rF; bB = c6p[2][ch] * lF + c6p[3][ch] * rF;
                                num = bT * tF + bB * bF;
                                int blendC6E = (num + rounding) / (1024 * 1024), blendC6OE = num / (1024 * 1024);
                                recT[ch][dx + dy * tsx] = blendC6E;                         /* EC.cpp:3969-3971 */
                                if (abs(cur - blendC) > rejectFactor) rej = 1;
                                if (abs(cur - blendC6) > rejectFactor) rej6 = 1;
                                if (abs(cur - blendCO) > rejectFactor) rejO = 1;
                                if (abs(cur - blendC6O) > rejectFactor) rej6O = 1;
                                if (abs(cur - blendC6OE) > rejectFactor) rej6OE = 1;
                                if (abs(cur - blendC6E) > rejectFactor) rej6E = 1;
                            }
                        }
                    }
                    if (!((!rej || !rejO) || (!rej6 || !rej6O) || (!rej6OE || !rej6E))) continue;   /* EC.cpp:3998 */
                    int enc[4][3];
                    const int lx[4] = { x, x + tsx, x, x + tsx }, ly[4] = { y, y, y + tsy, y + tsy };
                    for (int n = 0; n < 3; n++)                 /* EC.cpp:4001-4021: corner claim on the (W+1)x(H+1) lattice */
                        for (int k = 0; k < 4; k++) {
                            size_t li = lx[k] + (size_t)ly[k] * (W + 1);
                            enc[k][n] = c->mappedRGB[n][li];
                            if (!enc[k][n]) c->mappedRGB[n][li] = 255;
                        }
                    bitmap[pos >> 3] |= (uint8_t)(1 << (pos & 7));      /* EC.cpp:4026 */
                    for (int dy = 0; dy < tsy; dy++)
                        for (int dx = 0; dx < tsx; dx++) {              /* EC.cpp:4029-4037, 4096-4104 */
                            size_t i = (x + dx) + (size_t)(y + dy) * W;
                            c->smoothMap[i] = 255;
                            for (int n = 0; n < 3; n++) { c->mapSmoothTile[n][i] = 255; c->recon[n][i] = recT[n][dx + dy * tsx]; }
                            c->mipmapMask[i] = 0;
                        }
                    if (minX > x) minX = x;
                    if (minY > y) minY = y;
                    if (maxX < x + tsx) maxX = x + tsx;
                    if (maxY < y + tsy) maxY = y + tsy;
                    done++;
                    for (int k = 0; k < 4; k++)                 /* EC.cpp:4115-4132: TL, TR, BL, BR x R, G, B */
                        for (int n = 0; n < 3; n++)
                            if (!enc[k][n]) *wr++ = (uint8_t)CompressF(c6[k][n], 250);
                }
            }
        }
    }
    *rgbBytes = (int)(wr - rgb);
    *tileDone = done;
    bbox[0] = minX; bbox[1] = minY; bbox[2] = maxX; bbox[3] = maxY;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Range stage R2: DynamicTileCompressor (EC.cpp:8398-8522) with FindAndRemoveMostUsedColor (8335-8356),
 * Model1 (8358-8381), GetValueModel1 (8383-8391), DecompModel1 (8393-8396). */
int yko_range1d(yko_ctx* c, int plane, uint8_t* idx, int* idxBytes, uint8_t* type, int* typeBytes, int32_t* debug) {
    int w = c->w, h = c->h;
    const int32_t* src = c->src[plane];
    const int32_t* map = c->mapSmoothTile[plane];
    uint8_t* st = idx; uint8_t* pt = type;
    for (int y = 0; y < h; y += 8)
        for (int x = 0; x < w; x += 8) {
            uint8_t histo[256]; uint8_t values[64], offX[64], offY[64];
            int n = 0;
            memset(histo, 0, sizeof histo);
            for (int y2 = 0; y2 < 8; y2 += 4) {
                int hasLeft = pix_clamped(map, w, h, x, y + y2) == 0;          /* EC.cpp:8422-8430 */
                int hasRight = pix_clamped(map, w, h, x + 4, y + y2) == 0;
                if (hasLeft | hasRight) {
                    int lengthX = (hasLeft && hasRight) ? 8 : 4;
                    int x2 = (lengthX == 4 && hasRight) ? 4 : 0;
                    for (int iy = 0; iy < 4; iy++)
                        for (int ix = 0; ix < lengthX; ix++) {
                            int v = pix_clamped(src, w, h, x + x2 + ix, y + y2 + iy);
                            v = CompressF(v, 255);                              /* EC.cpp:8442 */
                            histo[v & 255]++;
                            values[n] = (uint8_t)v; offX[n] = (uint8_t)(x2 + ix); offY[n] = (uint8_t)(y2 + iy);
                            n++;
                        }
                }
            }
            if (n == 0) continue;
            int best = -1, bestV = -1;
            for (int k = 0; k < 256; k++) if (histo[k] >= bestV) { bestV = histo[k]; best = k; }    /* EC.cpp:8339-8344 */
            if (best == 0) best = 1;
            if (best == 255) best = 254;
            histo[best - 1] = 0; histo[best] = 0; histo[best + 1] = 0;
            int mn = 99999, mx = -99999;
            for (int k = 0; k < 256; k++) if (histo[k]) { if (mn > k) mn = k; if (mx < k) mx = k; }
            int minCol = 0, delta = 0;
            if (mn != 99999) { minCol = mn; delta = mx - mn; }
            for (int k = 0; k < n; k++) {                                      /* EC.cpp:8487-8501 */
                int v = values[k];
                if (v >= best - 1 && v <= best + 1) {
                    *st++ = 0;
                    if (debug) debug[(x + offX[k]) + (size_t)(y + offY[k]) * w] = best;
                } else {
                    int r = delta ? (((v - minCol) * 15) + ((delta >> 1) - 1)) / delta : 0;       /* EC.cpp:8383-8391 */
                    *st++ = (uint8_t)(1 + r);
                    if (debug) debug[(x + offX[k]) + (size_t)(y + offY[k]) * w] = minCol + (r * delta) / 15;
                }
            }
            *pt++ = (uint8_t)best; *pt++ = (uint8_t)minCol; *pt++ = (uint8_t)delta;                /* EC.cpp:8503-8505 */
        }
    *idxBytes = (int)(st - idx); *typeBytes = (int)(pt - type);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Range stage R1 tables: DynamicTile::buildTable (EC.cpp:625-699), Min/DiffRange{En,De}code (587-623). */
int yko_dyn_table(int minV, int maxV, int mode, int lut[16], int* base6, int* range7) {
    if (minV < 0 || maxV < minV || maxV > 255 || mode < 0 || mode > 5) return 0;
    int mn = minV > 224 ? 224 : minV;
    int diff = maxV - mn;
    if (diff < 16) diff = 16;
    int b = (mn * 63 + 112) / 224;                  /* MinRangeEncode, EC.cpp:587-597 */
    int BN = (b * 224) / 63;                        /* MinRangeDecode, EC.cpp:599-602 */
    int d8 = diff < 32 ? 32 : diff;                 /* DiffRangeEncode, EC.cpp:604-618 */
    int scale = (255 - 32) - BN;
    int r7 = ((d8 - 32) * 127 + (scale - 1)) / scale;
    int D = (r7 * scale) / 127 + 32;                /* DiffRangeDecode, EC.cpp:620-623 */
    float DF = (float)D;
    int count = mode < 3 ? 16 : 8;
    for (int i = 0; i < count; i++) {               /* EC.cpp:662-696 */
        float pos = (float)i / (count == 16 ? 15.0f : 7.0f);
        float nv = mode % 3 == 0 ? pos : mode % 3 == 1 ? powf(pos, 1.4f) : 1.0f - powf(1.0f - pos, 1.4f);
        float out = nv * DF;
        lut[i] = (int)((float)BN + out);
    }
    *base6 = b; *range7 = r7;
    return count;
}

/* Chroma front-end of the range stage. */

/* RGBtoYCoCg (EC.cpp:53-67) over the image: Image::ConvertToRGB2YCoCg(true), Image.cpp:285-321.  C `/`: truncation. */
void yko_rgb_to_ycocg(yko_ctx* c, int32_t* oY, int32_t* oCo, int32_t* oCg) {
    size_t n = (size_t)c->w * c->h;
    for (size_t i = 0; i < n; i++) {
        int R = c->src[0][i], G = c->src[1][i], B = c->src[2][i];
        int Co = R - B;
        int tmp = B + Co / 2;
        int Cg = G - tmp;
        oY[i] = tmp + Cg / 2; oCo[i] = Co / 2; oCg[i] = Cg / 2;
    }
}

/* Plane::SampleDown (Plane.cpp:278-369).  mode = EDownSample (framework.h:60-66): 0 NEAREST_TL, 1 NEAREST_BR,
 * 2 AVERAGE_BOX, 3 MAX_BOX, 4 MIN_BOX.  dst is (halfX ? w/2 : w) x (halfY ? h/2 : h).  The reference loads the four
 * source samples of every output whatever the mode; the combinations whose RESULT depends on a sample outside the
 * plane (NEAREST_BR / MAX_BOX / MIN_BOX with one axis only) are outside the parity domain: returns -1. */
int yko_sample_down(const int32_t* src, int w, int h, int halfX, int halfY, int mode, int32_t* dst) {
    if (mode < 0 || mode > 4 || (w & 1) || (h & 1)) return -1;
    if ((mode == 1 || mode == 3 || mode == 4) && !(halfX && halfY)) return -1;
    int stepX = halfX ? 2 : 1, stepY = halfY ? 2 : 1;
    int strideDst = halfX ? w / 2 : w;
    for (int y = 0; y < h; y += stepY)
        for (int x = 0; x < w; x += stepX) {
            size_t idx = x + (size_t)y * w;
            int A = src[idx];
            int v = A;
            if (mode == 2) {
                if (halfX && halfY) v = (A + src[idx + 1] + src[idx + w] + src[idx + 1 + w]) / 4;
                else if (halfX) v = (A + src[idx + 1]) / 2;
                else if (halfY) v = (A + src[idx + w]) / 2;
            } else if (mode != 0) {                                  /* both axes halved here */
                int B = src[idx + 1], C = src[idx + w], D = src[idx + 1 + w];
                if (mode == 1) v = D;
                else if (mode == 3) { int a = A > B ? A : B, b = C > D ? C : D; v = a > b ? a : b; }
                else { int a = A < B ? A : B, b = C < D ? C : D; v = a < b ? a : b; }
            }
            dst[(size_t)(y >> (halfY ? 1 : 0)) * strideDst + (x >> (halfX ? 1 : 0))] = v;
        }
    return 0;
}

/* DynamicTileEncode (EC.cpp:4365-4602) with LeftRightOrder (framework.h:228-256), GetTileEncode_Y (EC.cpp:1214-1221),
 * Plane::GetMinMax_Y (Plane.cpp:489-587), GetTileDynamic_Y (EC.cpp:747-1212), on any plane `src` of pw x ph samples:
 * the full-resolution planes (pw == w) or a SampleDown'ed chroma plane (halfX / halfY).  Two validity rules coexist for
 * the reduced planes, as in the reference: the min/max of a block runs over "any of the covered mask pixels"
 * (Plane.cpp:528-555, whose row stride is the REDUCED plane's width - kept as is), the coding over "all of them"
 * (EC.cpp:831-861).  isChroma = isCo | isCg: blocks with a negative minimum are written back minus 128 (EC.cpp:4441). */
int yko_range_dyn_plane(yko_ctx* c, const int32_t* src, int pw, int ph, int mode3BitOnly, int isChroma, int halfX, int halfY,
                        uint8_t* nibbles, int* nNibbles, uint16_t* defs, int* nDefs, int32_t* dst, int constraint[4]) {
    int W = c->w;                                                                   /* validPixel->GetWidth() */
    int shX = halfX ? 1 : 0, shY = halfY ? 1 : 0;
    const int32_t* mask = c->mipmapMask; const int32_t* smooth = c->smoothMap;
    int cx = (c->boundX0 >> 3) << 3, cy = (c->boundY0 >> 3) << 3;                   /* EC.cpp:4386-4391 */
    int cw = (((c->boundX1 + 7) >> 3) << 3) - cx, chh = (((c->boundY1 + 7) >> 3) << 3) - cy;
    if (halfX) { cx >>= 1; cw >>= 1; }                                              /* EC.cpp:4393-4401 */
    if (halfY) { cy >>= 1; chh >>= 1; }
    constraint[0] = cx; constraint[1] = cy; constraint[2] = cw; constraint[3] = chh;
    size_t maxNib = (size_t)(pw / 8) * (ph / 8) * 32;
    memset(nibbles, 0, maxNib);                                                     /* EC.cpp:4422 */
    int indexGlobal = 0, nd = 0;
    int x = cx - 8, y = cy;                                                         /* LeftRightOrder::Start */
    for (;;) {
        int valid = y < cy + chh;                                                   /* framework.h:239-255 */
        if (valid) {
            x += 8;
            if (x >= cx + cw) { x = cx; y += 8; valid = y < ph; }
        }
        if (!valid) break;
        int rw = (x + 8 > cw) ? (x % 8) : 8;
        int rh = (y + 8 > chh) ? (y % 8) : 8;
        /* GetMinMax_Y, Plane.cpp:489-587 (rect clipped to the plane) */
        int mn = 99999999, mx = -99999999, any = 0;
        int maxY = y + rh > ph ? ph : y + rh, maxX = x + rw > pw ? pw : x + rw;
        for (int yy = y; yy < maxY; yy++)
            for (int xx = x; xx < maxX; xx++) {
                size_t vi = ((size_t)xx << shX) + (size_t)(yy << shY) * pw;        /* `w` of the sampled plane, Plane.cpp:516 */
                int ok;
                if (!halfX && !halfY) ok = mask[vi] && !smooth[vi];
                else {
                    int hasGrad = smooth[vi] != 0, a = mask[vi] != 0, b, cc, d;
                    if (halfX && halfY) { b = mask[vi + 1] != 0; cc = mask[vi + pw] != 0; d = mask[vi + pw + 1] != 0; ok = !hasGrad && (a | b | cc | d); }
                    else if (halfX) { b = mask[vi + 1] != 0; ok = !hasGrad && (a | b); }
                    else { b = mask[vi + pw] != 0; ok = !hasGrad && (a | b); }
                }
                if (ok) {
                    int V = src[xx + (size_t)yy * pw];
                    if (V < mn) mn = V;
                    if (V > mx) mx = V;
                    any = 1;
                }
            }
        if (!any) { mn = 0; mx = 0; }
        /* GetTileDynamic_Y, EC.cpp:747-1212 */
        int useSigned = 0;
        if (mn < 0) { mn += 128; mx += 128; useSigned = 1; }
        int best[64], bestCode[64], bestMode = -1;
        float bestErr = 99999999.0f;
        int b6 = 0, r7 = 0;
        int bT[64], bC[64];
        for (int k = 0; k < 64; k++) { bC[k] = 0; best[k] = -999; bestCode[k] = 0; }
        for (int mode = mode3BitOnly ? 3 : 0; mode < 6; mode++) {
            int lut[16];
            int count = yko_dyn_table(mn, mx, mode, lut, &b6, &r7);
            float err = 0.0f;
            for (int k = 0; k < 64; k++) bT[k] = -999;
            for (int yy = 0; yy < rh; yy++)
                for (int xx = 0; xx < rw; xx++) {
                    size_t i = ((size_t)(xx + x) << shX) + (size_t)((yy + y) << shY) * W;
                    int ok;                                                         /* EC.cpp:826-861 */
                    if (!halfX && !halfY) ok = mask[i] && !smooth[i];
                    else {
                        int hasGrad = smooth[i] != 0, a = mask[i] != 0, b, cc, d;
                        if (halfX && halfY) { b = mask[i + 1] != 0; cc = mask[i + W] != 0; d = mask[i + W + 1] != 0; ok = !hasGrad && (a & b & cc & d); }
                        else if (halfX) { b = mask[i + 1] != 0; ok = !hasGrad && (a & b); }
                        else { b = mask[i + W] != 0; ok = !hasGrad && (a & b); }
                    }
                    if (!ok) continue;
                    int vo = src[(xx + x) + (size_t)(yy + y) * pw] + (useSigned ? 128 : 0);
                    int minDiff = 99999, found = 0, vfound = 0;
                    for (int n = 0; n < count; n++) {                               /* EC.cpp:873-881 */
                        int d = abs(lut[n] - vo);
                        if (d < minDiff) { minDiff = d; found = n; vfound = lut[n]; }
                    }
                    if (vo != 0) err += ((float)minDiff / (float)vo);               /* EC.cpp:884-886 */
                    bT[xx + (yy << 3)] = vfound; bC[xx + (yy << 3)] = found;
                }
            if (err <= bestErr) {                                                   /* EC.cpp:897-905 */
                bestErr = err; bestMode = mode;
                memcpy(best, bT, sizeof best); memcpy(bestCode, bC, sizeof bestCode);
            }
        }
        int valueCount = 0;
        int offset = isChroma ? (useSigned ? -128 : 0) : 0;                         /* EC.cpp:4441-4442 */
        for (int yy = 0; yy < rh; yy++)
            for (int xx = 0; xx < rw; xx++) {                                       /* EC.cpp:1174-1190 */
                int v = best[xx + (yy << 3)];
                if (v != -999) {
                    if (indexGlobal & 1) nibbles[indexGlobal >> 1] |= (uint8_t)(bestCode[xx + (yy << 3)] << 4);
                    else nibbles[indexGlobal >> 1] |= (uint8_t)bestCode[xx + (yy << 3)];
                    indexGlobal++; valueCount++;
                    if (dst) dst[((size_t)(xx + x) << shX) + ((size_t)(yy + y) << shY) * W] = v + offset;   /* EC.cpp:4444-4502 */
                }
            }
        if (valueCount) defs[nd++] = (uint16_t)((bestMode << 13) | (r7 << 7) | b6); /* EC.cpp:4437; YAIK_private.h:358 */
    }
    *nNibbles = indexGlobal; *nDefs = nd;
    return 0;
}

int yko_range_dyn(yko_ctx* c, int plane, int mode3BitOnly, uint8_t* nibbles, int* nNibbles,
                  uint16_t* defs, int* nDefs, int32_t* dst, int constraint[4]) {
    return yko_range_dyn_plane(c, c->src[plane], c->w, c->h, mode3BitOnly, 0, 0, 0, nibbles, nNibbles, defs, nDefs, dst, constraint);
}
