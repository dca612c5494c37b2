// Host-side C++ mirror of the reference's `struct EncoderContext` for the encoder-analysis stage
// (KLab/YAIK encoder/EncoderContext.h:185-380, bodies in encoder/EncoderContext.cpp = "EC.cpp").
//
// Same member names, argument meaning and call order as the reference's Convert() uses (EC.cpp:9027-9093, 9451-9460,
// 9542-9544); each body forwards to the C ABI in include/yaik_b200.h, i.e. to the CUDA kernels.  What the reference
// keeps in function locals and hands to its host tail (PaletteCompressor / ZSTD / fwrite) is kept in `last*` members
// so the tail can be attached unchanged.  There is no CPU implementation behind these members.
#pragma once
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <stdio.h>
#include <thread>
#include <vector>
#include "framework.h"
#include "../../include/yaik_b200.h"

extern u8* streamType;      // EC.cpp:8217 — the R2 type stream the reference keeps in globals
extern u8* pType;           // EC.cpp:8218

struct EncoderContext {
    explicit EncoderContext(int cudaDevice = 0);
    ~EncoderContext();

    void SetImageToEncode(Image* img);              // EncoderContext.h:230
    // stage members, in the order Convert() calls them
    void MipPrefilter(bool active);                 // EncoderContext.h:332, EC.cpp:1257
    void PrepareQuadSmooth();                       // EncoderContext.h:327, EC.cpp:2796 (empty in the reference; here: the fused 7-pass launch)
    int  FittingQuadSmooth(int rejectFactor, Plane* a, Plane* b, Plane* c, Image* testOutput, bool useYCoCg,
                           int tileBitSizeX, int tileBitSizeY);        // EncoderContext.h:329, EC.cpp:3710
    u8*  DynamicTileCompressor(u8* stream, Plane* src, Plane* map, Plane* debug);     // EncoderContext.h:236, EC.cpp:8398
    int  DynamicTileEncode(bool mode3BitOnly, Plane* plane, Plane* dst, bool isCo, bool isCg, bool isHalfX, bool isHalfY);   // EncoderContext.h:370, EC.cpp:4365
    void CheckMipmapMask();                         // EC.cpp:2784-2794
    // chroma front-end of the range stage (Convert() holds it at EC.cpp:9539-9545).  convRGB2YCoCg only records the
    // request; chromaReduction runs RGB -> YCoCg and the SampleDown of Co / Cg in one kernel (yk_chroma_prepare) and
    // brings Y, workCo and workCg back.  DynamicTileEncode recognises those three planes by pointer.
    void convRGB2YCoCg(bool notUseRGBAsIs);         // EncoderContext.h:334, EC.cpp:2766
    void chromaReduction();                         // EncoderContext.h:336, EC.cpp:2770

    // expands the compact device state into the int32 state planes below (and into `testOutput` of the last
    // FittingQuadSmooth call) — the reference updates them in place on every call; here it is explicit because it
    // costs 44 bytes per pixel of download that the hot path does not need
    void SyncStatePlanes();

    // ---- state with the reference's names (EncoderContext.h:300-323)
    Image* original;
    Plane* mipmapMask;
    Plane* smoothMap;
    Image* mapSmoothTile;
    Image* mappedRGB;           // (W+1) x (H+1)
    int boundX0, boundY0, boundX1, boundY1;
    int remainingPixels, mipMapTileSize;
    int colorCompressionQuad, colorCompression1D, rangeCompression1D;   // 250, 255, 15 (EncoderContext.h:221-224)
    bool halfCoW, halfCoH, halfCgW, halfCgH;        // EncoderContext.h:265-271 (the CLI sets W only, AVERAGE_BOX: ImageEncoder.cpp:175-181)
    EDownSample downSampleCo, downSampleCg;
    Image* YCoCgImg;            // plane 0 = Y.  Co / Cg exist at their working size only: workCo / workCg
    Plane* workCo;
    Plane* workCg;
    int lastError;              // YK_* code of the last failing call (the reference printf()s and carries on)

    // ---- what the reference's host tails consume
    struct { std::vector<u8> bitmap; int bbox[4]; int remainingPixels; bool wroteChunk; int chunkBBoxTiles[4]; } lastAlpha;     // MIPM chunk, EC.cpp:1367-1396
    struct { std::vector<u8> bitmap, rgbStream; int minX, minY, maxX, maxY, tileDone, shX, shY; } lastGradient;                  // GTIL chunk, EC.cpp:4239-4350
    struct { std::vector<u8> nibbles; std::vector<u16> tileDefs; int nNibbles; BoundingBox constraint; } lastDynamic;            // PLNT chunk, EC.cpp:4515-4589

    // ---- host tails: what the reference does after each stage - PaletteCompressor, entropy coding,
    // fwrite of the chunk into outFile (EC.cpp:1367-1396, 4239-4350, 4515-4589, 8524-8576).  They are attached once a
    // compressor callback is set (the reference links zstd; this library links none): every stage member then appends its
    // chunk to outFile, in call order.  With SetAsyncTails(true) the tails run on a worker thread, so PaletteCompressor and
    // the entropy coder of one stage (or image) overlap the GPU analysis of the next; FinishTails() waits for them.
    FILE* outFile;              // EncoderContext.h:306 of the reference
    int   fileOutSize;          // bytes of compressed streams written (CompressStream, EC.cpp:3703)
    void SetCompressor(yk_compress_fn fn, void* user, int paletteMode = YK_PALETTE_BUG_COMPATIBLE);
    void SetAsyncTails(bool on);
    void FinishTails();
    void WriteFileHeader();                                         // EC.cpp:9007-9016
    void WriteEndTag();                                             // EC.cpp:9779-9782
    void GenerateDynamicTileChunk(u8* stream, int sizeStream);      // EncoderContext.h:235, EC.cpp:8524-8576 (type stream: streamType .. pType)

private:
    yk_compress_fn compressFn; void* compressUser;
    yk_palette* palette;
    bool asyncTails, workerStop;
    std::thread worker;
    std::mutex tailMutex;
    std::condition_variable tailCv, tailIdleCv;
    std::deque<std::function<void()> > tailQueue;
    int tailsInFlight;
    void postTail(std::function<void()> job);
    void writeChunk(const std::vector<u8>& chunk);
    void workerLoop();
    yk_ctx* ctx;
    int device, capW, capH;
    Image* lastTestOutput;
    bool prepared, useYCoCgPlanes;
    int planeIndexOf(Plane* p, Image* img);
    void ensureContext(int w, int h, int planes);
};
