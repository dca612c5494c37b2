// TEST DRIVER for the host-side C++ mirror (yaik_b200/host/EncoderContext.h): calls the stage members exactly the way
// the reference's Convert() does (EC.cpp:9027-9093, 9451-9460, 9542-9544) and dumps the results in the same record
// format as oracle/ref_harness.cpp, so tests/test_host_mirror.py can compare them with the golden vectors / the oracle.
// usage: host_mirror_test <in.ykin> <out.ykout> [alpha] [grad] [r2] [r1] [r1_3bit] [noprepare]
//                         [yaik=<file>] [zstdlib=<shared library with ZSTD_compress>] [async] [decodable]
// yaik=: the host tails are attached (PaletteCompressor, chunk serialisers) and the chunks go to <file> in Convert()'s order;
// the entropy coder behind the callback is ZSTD_compress of zstdlib= (e.g. the reference's own build in oracle/_ref), or a
// stand-in that stores the bytes as they are; async: tails on the worker thread.
#include "EncoderContext.h"
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

static FILE* g_out = NULL;
static void rec(const char* name, char dtype, const void* data, uint64_t count) {
    char nm[32]; memset(nm, 0, sizeof nm); strncpy(nm, name, 31);
    size_t esz = dtype == 'i' ? 4 : dtype == 'H' ? 2 : dtype == 'd' ? 8 : 1;
    fwrite(nm, 1, 32, g_out); fwrite(&dtype, 1, 1, g_out); fwrite(&count, 8, 1, g_out);
    if (count) fwrite(data, esz, count, g_out);
}
static void recPlane(const char* name, Plane* p) { rec(name, 'i', p->GetPixels(), (uint64_t)p->GetWidth() * p->GetHeight()); }
static void recInts(const char* name, std::vector<int> v) { rec(name, 'i', v.data(), v.size()); }

typedef size_t (*zstd_fn)(void*, size_t, const void*, size_t, int);
static zstd_fn g_zstd = NULL;
static size_t compressCallback(void*, void* dst, size_t cap, const void* src, size_t n, int level) {
    if (g_zstd) { const size_t r = g_zstd(dst, cap, src, n, level); return r > cap ? 0 : r; }     // an error code is a huge value
    if (n + 1 > cap) return 0;                                                                        // stand-in: one marker byte + the bytes
    ((unsigned char*)dst)[0] = 0x53; memcpy((unsigned char*)dst + 1, src, n);
    return n + 1;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    bool doAlpha = false, doGrad = false, doR2 = false, doR1 = false, r1_3bit = false, prepare = true, doChroma = false;
    int chromaCfg[4] = { 1, 0, 1, 0 }, chromaModes[2] = { 2, 2 };
    std::string yaikPath, zstdLib;
    bool asyncTails = false, decodable = false;
    for (int i = 3; i < argc; i++) {
        std::string a = argv[i];
        if (a == "alpha") doAlpha = true; else if (a == "grad") doGrad = true; else if (a == "r2") doR2 = true;
        else if (a == "r1") doR1 = true; else if (a == "r1_3bit") { doR1 = true; r1_3bit = true; }
        else if (a == "noprepare") prepare = false;
        else if (a.rfind("yaik=", 0) == 0) yaikPath = a.substr(5);
        else if (a.rfind("zstdlib=", 0) == 0) zstdLib = a.substr(8);
        else if (a == "async") asyncTails = true;
        else if (a == "decodable") decodable = true;
        else if (a.rfind("chroma=", 0) == 0 && a.size() == 14) {                  // chroma=XYXY:MM, as oracle/ref_harness.cpp
            doChroma = true;
            for (int k = 0; k < 4; k++) chromaCfg[k] = a[7 + k] == '1';
            chromaModes[0] = a[12] - '0'; chromaModes[1] = a[13] - '0';
        }
    }
    FILE* fi = fopen(argv[1], "rb");
    char magic[4]; int hdr[3];
    if (!fi || fread(magic, 1, 4, fi) != 4 || (memcmp(magic, "YKIN", 4) && memcmp(magic, "YKI4", 4)) || fread(hdr, 4, 3, fi) != 3) { fprintf(stderr, "bad input\n"); return 1; }
    const bool wide = !memcmp(magic, "YKI4", 4);       // int32 samples (yaik_b200/synth.py to_ykin)
    const int W = hdr[0], H = hdr[1], NP = hdr[2];
    std::vector<int> px((size_t)W * H * NP);
    if (wide) {
        if (fread(px.data(), 4, px.size(), fi) != px.size()) return 1;
    } else {
        std::vector<u8> b(px.size());
        if (fread(b.data(), 1, b.size(), fi) != b.size()) return 1;
        for (size_t i = 0; i < b.size(); i++) px[i] = b[i];
    }
    fclose(fi);
    g_out = fopen(argv[2], "wb");

    EncoderContext ctx(0);
    Image* img = Image::CreateImage(W, H, NP, false);
    for (int c = 0; c < NP; c++) {
        int* d = img->GetPlane(c)->GetPixels(); const int* s = px.data() + (size_t)c * W * H;
        for (size_t i = 0; i < (size_t)W * H; i++) d[i] = s[i];
    }
    ctx.SetImageToEncode(img);
    Image* output = Image::CreateImage(W, H, 3, true);
    if (!yaikPath.empty()) {
        if (!zstdLib.empty()) {
            void* h = dlopen(zstdLib.c_str(), RTLD_NOW | RTLD_LOCAL);
            g_zstd = h ? (zstd_fn)dlsym(h, "ZSTD_compress") : NULL;
            if (!g_zstd) { fprintf(stderr, "cannot load ZSTD_compress from %s\n", zstdLib.c_str()); return 1; }
        }
        ctx.outFile = fopen(yaikPath.c_str(), "wb");
        if (!ctx.outFile) { perror(yaikPath.c_str()); return 1; }
        ctx.SetCompressor(compressCallback, NULL, decodable ? YK_PALETTE_DECODABLE : YK_PALETTE_BUG_COMPATIBLE);
        ctx.SetAsyncTails(asyncTails);
        ctx.WriteFileHeader();                                                      // EC.cpp:9007-9016
    }

    if (doAlpha && NP == 4) {
        ctx.MipPrefilter(true);                                                     // EC.cpp:9027
        recInts("alpha.bound", { ctx.boundX0, ctx.boundY0, ctx.boundX1, ctx.boundY1 });
        recInts("alpha.remaining", { ctx.remainingPixels, ctx.mipMapTileSize });
        if (ctx.lastAlpha.wroteChunk) recInts("alpha.chunk_bbox", { ctx.lastAlpha.chunkBBoxTiles[0], ctx.lastAlpha.chunkBBoxTiles[1], ctx.lastAlpha.chunkBBoxTiles[2], ctx.lastAlpha.chunkBBoxTiles[3], 1, 4 });
        else recInts("alpha.chunk_bbox", {});
        rec("alpha.bitmap", 'B', ctx.lastAlpha.bitmap.data(), ctx.lastAlpha.bitmap.size());
    }
    if (doGrad) {
        if (prepare) ctx.PrepareQuadSmooth();                                       // EC.cpp:9045
        static const int order[7][2] = { {4,4},{4,3},{3,4},{3,3},{3,2},{2,3},{2,2} };   // EC.cpp:9057-9093
        for (int k = 0; k < 7; k++) {
            int done = ctx.FittingQuadSmooth(3, img->GetPlane(0), img->GetPlane(1), img->GetPlane(2), output, false, order[k][0], order[k][1]);
            auto& g = ctx.lastGradient;
            const int wrote = (g.maxX > g.minX && g.maxY > g.minY && !g.rgbStream.empty()) ? 1 : 0;      // EC.cpp:4239
            char nm[32];
            snprintf(nm, sizeof nm, "grad%d.tiledone", k); recInts(nm, { done, wrote });
            snprintf(nm, sizeof nm, "grad%d.bbox", k);     recInts(nm, { g.minX, g.minY, g.maxX - g.minX, g.maxY - g.minX });   // EC.cpp:4255-4258
            snprintf(nm, sizeof nm, "grad%d.bitmap", k);   rec(nm, 'B', g.bitmap.data(), g.bitmap.size());
            snprintf(nm, sizeof nm, "grad%d.rgb", k);      rec(nm, 'B', g.rgbStream.data(), g.rgbStream.size());
        }
        ctx.SyncStatePlanes();
        recPlane("state.smoothMap", ctx.smoothMap);
        recPlane("state.mipmapMask", ctx.mipmapMask);
        for (int c = 0; c < 3; c++) {
            char nm[32];
            snprintf(nm, sizeof nm, "state.mapSmoothTile%d", c); recPlane(nm, ctx.mapSmoothTile->GetPlane(c));
            snprintf(nm, sizeof nm, "state.mappedRGB%d", c);     recPlane(nm, ctx.mappedRGB->GetPlane(c));
            snprintf(nm, sizeof nm, "state.recon%d", c);         recPlane(nm, output->GetPlane(c));
        }
    }
    if (doR2) {
        if (!ctx.mapSmoothTile) ctx.SyncStatePlanes();
        std::vector<u8> stream((size_t)W * H * 3 + 64);
        streamType = new u8[(size_t)(W / 8 + 1) * (H / 8 + 1) * 9 + 64]; pType = streamType;
        u8* p = stream.data();
        for (int c = 0; c < 3; c++) {
            u8* p0 = p; u8* t0 = pType;
            p = ctx.DynamicTileCompressor(p, img->GetPlane(c), ctx.mapSmoothTile->GetPlane(c), output->GetPlane(c));       // EC.cpp:9451-9460
            char nm[32];
            snprintf(nm, sizeof nm, "r2.idx%d", c);  rec(nm, 'B', p0, p - p0);
            snprintf(nm, sizeof nm, "r2.type%d", c); rec(nm, 'B', t0, pType - t0);
        }
        ctx.GenerateDynamicTileChunk(stream.data(), (int)(p - stream.data()));      // EC.cpp:9465
    }
    if (doR1) {
        for (int c = 0; c < 3; c++) {
            Plane* dst = new Plane(W, H);
            BoundingBox all = dst->GetRect(); dst->Fill(all, -1);
            int ret = ctx.DynamicTileEncode(r1_3bit, img->GetPlane(c), dst, false, false, false, false);                    // EC.cpp:9542-9544
            auto& d = ctx.lastDynamic;
            char nm[32];
            snprintf(nm, sizeof nm, "r1.defs%d", c);    rec(nm, 'H', d.tileDefs.data(), d.tileDefs.size());
            snprintf(nm, sizeof nm, "r1.nibbles%d", c); rec(nm, 'B', d.nibbles.data(), d.nibbles.size());
            snprintf(nm, sizeof nm, "r1.hdr%d", c);     recInts(nm, { d.constraint.x, d.constraint.y, d.constraint.w, d.constraint.h, (int)d.nibbles.size(), 1, 0, ret });
            snprintf(nm, sizeof nm, "r1.dst%d", c);     recPlane(nm, dst);
            delete dst;
        }
    }
    if (doChroma) {                                                                 // EC.cpp:9539-9545
        ctx.halfCoW = chromaCfg[0]; ctx.halfCoH = chromaCfg[1]; ctx.halfCgW = chromaCfg[2]; ctx.halfCgH = chromaCfg[3];
        ctx.downSampleCo = (EDownSample)chromaModes[0]; ctx.downSampleCg = (EDownSample)chromaModes[1];
        ctx.convRGB2YCoCg(true);
        ctx.chromaReduction();
        if (!ctx.lastError) {
            recPlane("yc.Y", ctx.YCoCgImg->GetPlane(0)); recPlane("yc.workCo", ctx.workCo); recPlane("yc.workCg", ctx.workCg);
            Plane* src[3] = { ctx.YCoCgImg->GetPlane(0), ctx.workCo, ctx.workCg };
            const bool m3[3] = { false, false, true }, isCo[3] = { false, true, false }, isCg[3] = { false, false, true };
            const bool hx[3] = { false, ctx.halfCoW, ctx.halfCgW }, hy[3] = { false, ctx.halfCoH, ctx.halfCgH };
            for (int c = 0; c < 3; c++) {
                Plane* dst = new Plane(W, H);
                BoundingBox all = dst->GetRect(); dst->Fill(all, -1000);
                int ret = ctx.DynamicTileEncode(m3[c], src[c], dst, isCo[c], isCg[c], hx[c], hy[c]);
                auto& d = ctx.lastDynamic;
                char nm[32];
                snprintf(nm, sizeof nm, "yc.defs%d", c);    rec(nm, 'H', d.tileDefs.data(), d.tileDefs.size());
                snprintf(nm, sizeof nm, "yc.nibbles%d", c); rec(nm, 'B', d.nibbles.data(), d.nibbles.size());
                snprintf(nm, sizeof nm, "yc.hdr%d", c);     recInts(nm, { d.constraint.x, d.constraint.y, d.constraint.w, d.constraint.h, (int)d.nibbles.size(), 1, 0, ret });
                snprintf(nm, sizeof nm, "yc.dst%d", c);     recPlane(nm, dst);
                delete dst;
            }
        }
    }
    if (ctx.outFile) {
        ctx.WriteEndTag();                                                          // EC.cpp:9779-9782
        ctx.FinishTails();
        fclose(ctx.outFile); ctx.outFile = NULL;
    }
    recInts("meta", { W, H, NP, ctx.lastError });
    fclose(g_out);
    return ctx.lastError ? 3 : 0;
}
