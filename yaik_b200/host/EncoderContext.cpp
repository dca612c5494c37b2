// See EncoderContext.h.  Every stage member is a thin call into the C ABI (include/yaik_b200.h).
#include "EncoderContext.h"
#include <stdio.h>

u8* streamType = new u8[1];     // re-pointed by the caller exactly as with the reference (its 100000-byte buffer overflows
u8* pType = streamType;         // beyond 33 333 tile records); DynamicTileCompressor appends at pType

EncoderContext::EncoderContext(int cudaDevice)
    : original(NULL), mipmapMask(NULL), smoothMap(NULL), mapSmoothTile(NULL), mappedRGB(NULL),
      boundX0(0), boundY0(0), boundX1(0), boundY1(0), remainingPixels(0), mipMapTileSize(16),
      colorCompressionQuad(250), colorCompression1D(255), rangeCompression1D(15),
      halfCoW(true), halfCoH(false), halfCgW(true), halfCgH(false), downSampleCo(AVERAGE_BOX), downSampleCg(AVERAGE_BOX),
      YCoCgImg(NULL), workCo(NULL), workCg(NULL), lastError(YK_OK),
      outFile(NULL), fileOutSize(0), compressFn(NULL), compressUser(NULL), palette(NULL), asyncTails(false), workerStop(false), tailsInFlight(0),
      ctx(NULL), device(cudaDevice), capW(0), capH(0), lastTestOutput(NULL), prepared(false), useYCoCgPlanes(false) {}

EncoderContext::~EncoderContext() {
    FinishTails();
    if (worker.joinable()) {
        { std::lock_guard<std::mutex> lk(tailMutex); workerStop = true; }
        tailCv.notify_all();
        worker.join();
    }
    if (palette) yk_palette_destroy(palette);
    if (ctx) yk_destroy(ctx);
    delete mipmapMask; delete smoothMap; delete mapSmoothTile; delete mappedRGB;
    delete YCoCgImg; delete workCo; delete workCg;
}

void EncoderContext::ensureContext(int w, int h, int planes) {
    if (ctx && w <= capW && h <= capH) return;
    if (ctx) yk_destroy(ctx);
    ctx = NULL;
    lastError = yk_create(&ctx, device, w, h, 4, 1);
    if (lastError) { fprintf(stderr, "yaik_b200: yk_create failed: %s %s\n", yk_error_string(lastError), yk_last_cuda_error()); return; }
    capW = w; capH = h; (void)planes;
}

void EncoderContext::SetImageToEncode(Image* img) {
    original = img;
    const int w = img->GetWidth(), h = img->GetHeight(), n = img->HasAlpha() ? 4 : 3;
    ensureContext(w, h, n);
    if (!ctx) return;
    const int32_t* planes[4] = { 0, 0, 0, 0 };
    for (int i = 0; i < n; i++) planes[i] = img->GetPlane(i)->GetPixels();
    lastError = yk_set_image(ctx, 0, planes, n, w, h);
    delete mipmapMask; delete smoothMap; delete mapSmoothTile; delete mappedRGB;
    mipmapMask = NULL; smoothMap = NULL; mapSmoothTile = NULL; mappedRGB = NULL;
    delete YCoCgImg; delete workCo; delete workCg;
    YCoCgImg = NULL; workCo = NULL; workCg = NULL;
    boundX0 = 0; boundY0 = 0; boundX1 = w; boundY1 = h; remainingPixels = w * h; mipMapTileSize = 16;
    prepared = false; lastTestOutput = NULL;
}

void EncoderContext::CheckMipmapMask() {            // EC.cpp:2784-2794: all-255 mask, bound = full image
    if (!mipmapMask && original) {
        mipmapMask = new Plane(original->GetWidth(), original->GetHeight(), false);
        BoundingBox bb = mipmapMask->GetRect();
        mipmapMask->Fill(bb, 255);
    }
}

void EncoderContext::MipPrefilter(bool active) {
    (void)active;                                   // the reference ignores it as well (EC.cpp:1257)
    if (!ctx || !original) { lastError = YK_ERR_STATE; return; }
    const int w = original->GetWidth(), h = original->GetHeight();
    if (!mipmapMask) mipmapMask = new Plane(w, h, false);
    if (!original->HasAlpha()) {                    // EC.cpp:1419-1426
        boundX0 = 0; boundY0 = 0; boundX1 = w; boundY1 = h; remainingPixels = w * h;
        BoundingBox bb = mipmapMask->GetRect(); mipmapMask->Fill(bb, 255);
        return;
    }
    lastAlpha.bitmap.assign((size_t)((w + 15) / 16) * ((h + 15) / 16) / 8 + 8, 0);
    int nb = 0, wrote = 0;
    lastError = yk_alpha_reject(ctx, 0, lastAlpha.bitmap.data(), (int)lastAlpha.bitmap.size(), &nb, lastAlpha.bbox,
                                &lastAlpha.remainingPixels, &wrote, lastAlpha.chunkBBoxTiles);
    if (lastError) return;
    lastAlpha.bitmap.resize(nb); lastAlpha.wroteChunk = wrote != 0;
    boundX0 = lastAlpha.bbox[0]; boundY0 = lastAlpha.bbox[1]; boundX1 = lastAlpha.bbox[2]; boundY1 = lastAlpha.bbox[3];
    remainingPixels = lastAlpha.remainingPixels; mipMapTileSize = 16;       // EC.cpp:1287-1291
    // host tail (EC.cpp:1367-1396): 'MIPM' HeaderBase + MipmapHeader + bitmap
    if (outFile && lastAlpha.wroteChunk) {
        const std::vector<u8> bitmap = lastAlpha.bitmap;
        int bb[4]; for (int k = 0; k < 4; k++) bb[k] = lastAlpha.chunkBBoxTiles[k];
        postTail([this, bitmap, bb]() {
            std::vector<u8> chunk(bitmap.size() + 64);
            size_t n = 0;
            if (yk_chunk_mipm(chunk.data(), chunk.size(), &n, bb, bitmap.data(), (int)bitmap.size()) == YK_OK) { chunk.resize(n); writeChunk(chunk); }
        });
    }
}

void EncoderContext::PrepareQuadSmooth() {
    if (!ctx || !original) { lastError = YK_ERR_STATE; return; }
    lastError = yk_prepare_quad_smooth(ctx, 0, 3);  // rejectFactor of Convert(), EC.cpp:9042
    prepared = lastError == YK_OK;
}

int EncoderContext::FittingQuadSmooth(int rejectFactor, Plane* a, Plane* b, Plane* c, Image* testOutput, bool useYCoCg,
                                      int tileBitSizeX, int tileBitSizeY) {
    CheckMipmapMask();
    if (!ctx || !original) { lastError = YK_ERR_STATE; return 0; }
    if (useYCoCg || a != original->GetPlane(0) || b != original->GetPlane(1) || c != original->GetPlane(2)) {
        lastError = YK_ERR_UNSUPPORTED;             // only the RGB form Convert() uses (PlaneBit 7, EC.cpp:9043)
        fprintf(stderr, "yaik_b200: FittingQuadSmooth is provided for the three RGB planes of the image only\n");
        return 0;
    }
    const int w = original->GetWidth(), h = original->GetHeight();
    const int tsx = 1 << tileBitSizeX, tsy = 1 << tileBitSizeY;
    if (!smoothMap) {                               // EC.cpp:3739-3749: state planes appear on the first call (filled by SyncStatePlanes)
        smoothMap = new Plane(w, h, false); smoothMap->Clear();
        mapSmoothTile = Image::CreateImage(w, h, 3, true, false);
        mappedRGB = Image::CreateImage(w + 1, h + 1, 3, true, false);
    }
    lastGradient.bitmap.assign((size_t)((w + 63) / 32) * ((h + 63) / 32) * 64 / 8 + 16, 0);
    lastGradient.rgbStream.assign((size_t)3 * (w / tsx + 1) * (h / tsy + 1) + 16, 0);   // EC.cpp:3780-3785
    int nb = 0, nr = 0, bbox[4] = { 0, 0, 0, 0 }, done = 0;
    lastError = yk_gradient_pass(ctx, 0, rejectFactor, tileBitSizeX, tileBitSizeY, lastGradient.bitmap.data(), (int)lastGradient.bitmap.size(), &nb,
                                 lastGradient.rgbStream.data(), (int)lastGradient.rgbStream.size(), &nr, bbox, &done);
    if (lastError) { fprintf(stderr, "yaik_b200: FittingQuadSmooth: %s %s\n", yk_error_string(lastError), yk_last_cuda_error()); return 0; }
    lastGradient.bitmap.resize(nb); lastGradient.rgbStream.resize(nr);
    lastGradient.minX = bbox[0]; lastGradient.minY = bbox[1]; lastGradient.maxX = bbox[2]; lastGradient.maxY = bbox[3];
    lastGradient.tileDone = done; lastGradient.shX = tileBitSizeX; lastGradient.shY = tileBitSizeY;
    lastTestOutput = testOutput;
    // host tail (EC.cpp:4239-4350): if (maxX > minX && maxY > minY && rgbStream.size() > 0)
    //   CompressStream(bitmap), PaletteCompressor(rgbStream) -> CompressStream, fwrite 'GTIL' + HeaderGradientTile + streams
    if (outFile && compressFn) {
        const std::vector<u8> bitmap = lastGradient.bitmap, rgb = lastGradient.rgbStream;
        const int shX = tileBitSizeX, shY = tileBitSizeY, cc = colorCompressionQuad;
        int bb[4] = { bbox[0], bbox[1], bbox[2], bbox[3] };
        postTail([this, bitmap, rgb, shX, shY, cc, bb]() {
            std::vector<u8> chunk(bitmap.size() * 2 + rgb.size() * 6 + 4096);
            size_t n = 0;
            const int rc = yk_chunk_gtil(chunk.data(), chunk.size(), &n, palette, compressFn, compressUser, shX, shY, 7, bb,
                                         bitmap.data(), (int)bitmap.size(), rgb.data(), (int)rgb.size(), cc);
            if (rc == YK_OK && n) { chunk.resize(n); fileOutSize += (int)n - 36; writeChunk(chunk); }
            else if (rc) lastError = rc;
        });
    }
    return done;
}

int EncoderContext::planeIndexOf(Plane* p, Image* img) {
    if (!img) return -1;
    for (int n = 0; n < 3; n++) if (img->GetPlane(n) == p) return n;
    return -1;
}

u8* EncoderContext::DynamicTileCompressor(u8* stream, Plane* src, Plane* map, Plane* debug) {
    (void)debug; (void)map;                         // map is mapSmoothTile[plane]: the device holds it in compact form
    if (!ctx || !original) { lastError = YK_ERR_STATE; return stream; }
    const int n = planeIndexOf(src, original);
    if (n < 0) { lastError = YK_ERR_UNSUPPORTED; return stream; }
    const int w = original->GetWidth(), h = original->GetHeight();
    int ni = 0, nt = 0;
    lastError = yk_range1d(ctx, 0, n, stream, w * h, &ni, pType, 3 * (w / 8 + 1) * (h / 8 + 1), &nt);
    if (lastError) { fprintf(stderr, "yaik_b200: DynamicTileCompressor: %s\n", yk_error_string(lastError)); return stream; }
    pType += nt;                                    // EC.cpp:8503-8505
    return stream + ni;
}

void EncoderContext::convRGB2YCoCg(bool notUseRGBAsIs) {
    useYCoCgPlanes = notUseRGBAsIs;                 // the conversion itself runs fused with chromaReduction()
}

void EncoderContext::chromaReduction() {
    if (!ctx || !original) { lastError = YK_ERR_STATE; return; }
    if (!useYCoCgPlanes) { lastError = YK_ERR_UNSUPPORTED; return; }       // "RGB as is" feeds DynamicTileEncode the image planes directly
    const int half[4] = { halfCoW, halfCoH, halfCgW, halfCgH }, mode[2] = { (int)downSampleCo, (int)downSampleCg };
    lastError = yk_chroma_prepare(ctx, 0, half, mode);
    if (lastError) { fprintf(stderr, "yaik_b200: chromaReduction: %s\n", yk_error_string(lastError)); return; }
    delete YCoCgImg; delete workCo; delete workCg;
    const int w = original->GetWidth(), h = original->GetHeight();
    YCoCgImg = Image::CreateImage(w, h, 1, false, false);
    workCo = new Plane(halfCoW ? w / 2 : w, halfCoH ? h / 2 : h, false);
    workCg = new Plane(halfCgW ? w / 2 : w, halfCgH ? h / 2 : h, false);
    Plane* out[3] = { YCoCgImg->GetPlane(0), workCo, workCg };
    for (int k = 0; k < 3 && !lastError; k++) lastError = yk_chroma_plane(ctx, 0, k, out[k]->GetPixels(), NULL, NULL);
}

int EncoderContext::DynamicTileEncode(bool mode3BitOnly, Plane* plane, Plane* dst, bool isCo, bool isCg, bool isHalfX, bool isHalfY) {
    CheckMipmapMask();
    if (!ctx || !original) { lastError = YK_ERR_STATE; return 0; }
    // which device plane: one of the image's colour planes, or Y / workCo / workCg of the chroma front-end
    int which = -1;
    if (YCoCgImg && plane == YCoCgImg->GetPlane(0)) which = 0;
    else if (plane && plane == workCo) which = 1;
    else if (plane && plane == workCg) which = 2;
    const int n = which < 0 ? planeIndexOf(plane, original) : -1;
    if (which < 0 && (n < 0 || isCo || isCg || isHalfX || isHalfY)) { lastError = YK_ERR_UNSUPPORTED; return 0; }
    if (which >= 0) {                               // the flags must be the ones the plane was made with
        const bool hx = which == 1 ? halfCoW : which == 2 ? halfCgW : false, hy = which == 1 ? halfCoH : which == 2 ? halfCgH : false;
        if (isHalfX != hx || isHalfY != hy || (isCo || isCg) != (which != 0)) { lastError = YK_ERR_ARG; return 0; }
    }
    const int pw = plane->GetWidth(), ph = plane->GetHeight(), nt = (pw / 8) * (ph / 8);
    lastDynamic.nibbles.assign((size_t)nt * 32 + 8, 0);
    lastDynamic.tileDefs.assign((size_t)nt + 8, 0);
    int nn = 0, nd = 0, cons[4] = { 0, 0, 0, 0 };
    if (which < 0)
        lastError = yk_range_dyn(ctx, 0, n, mode3BitOnly ? 1 : 0, lastDynamic.nibbles.data(), (int)lastDynamic.nibbles.size(), &nn,
                                 lastDynamic.tileDefs.data(), (int)lastDynamic.tileDefs.size(), &nd, cons, dst ? dst->GetPixels() : NULL);
    else
        lastError = yk_range_dyn_chroma(ctx, 0, which, mode3BitOnly ? 1 : 0, lastDynamic.nibbles.data(), (int)lastDynamic.nibbles.size(), &nn,
                                        lastDynamic.tileDefs.data(), (int)lastDynamic.tileDefs.size(), &nd, cons, dst ? dst->GetPixels() : NULL);
    if (lastError) { fprintf(stderr, "yaik_b200: DynamicTileEncode: %s\n", yk_error_string(lastError)); return 0; }
    lastDynamic.nibbles.resize((nn + 1) / 2); lastDynamic.tileDefs.resize(nd); lastDynamic.nNibbles = nn;
    lastDynamic.constraint.x = (s16)cons[0]; lastDynamic.constraint.y = (s16)cons[1];
    lastDynamic.constraint.w = (s16)cons[2]; lastDynamic.constraint.h = (s16)cons[3];
    // host tail (EC.cpp:4515-4589): ZSTD-21 of tileDefs and nibbles, fwrite 'PLNT' + PlaneTile
    if (outFile && compressFn) {
        const std::vector<u8> nib = lastDynamic.nibbles; const std::vector<u16> defs = lastDynamic.tileDefs;
        const int nNib = 2 * (int)nib.size();           // the reference closes a half byte (EC.cpp:4524-4526)
        const int type = isCo ? 1 : (isCg ? 2 : 0), hx = isHalfX, hy = isHalfY;
        int cb[4] = { cons[0], cons[1], cons[2], cons[3] };
        postTail([this, nib, defs, nNib, type, hx, hy, cb]() {
            std::vector<u8> chunk(nib.size() * 2 + defs.size() * 4 + 4096);
            size_t n = 0;
            const int rc = yk_chunk_plnt(chunk.data(), chunk.size(), &n, compressFn, compressUser, cb, defs.data(), (int)defs.size(), nib.data(), nNib, type, hx, hy);
            if (rc == YK_OK) { chunk.resize(n); writeChunk(chunk); } else lastError = rc;
        });
    }
    return 0;                                       // the reference returns layerSize, which it never updates (EC.cpp:4407, 4601)
}

void EncoderContext::SyncStatePlanes() {
    if (!ctx || !original) { lastError = YK_ERR_STATE; return; }
    const int w = original->GetWidth(), h = original->GetHeight();
    if (!smoothMap) smoothMap = new Plane(w, h, false);
    if (!mipmapMask) mipmapMask = new Plane(w, h, false);
    if (!mapSmoothTile) mapSmoothTile = Image::CreateImage(w, h, 3, true, false);
    if (!mappedRGB) mappedRGB = Image::CreateImage(w + 1, h + 1, 3, true, false);
    int32_t* mst[3] = { mapSmoothTile->GetPlane(0)->GetPixels(), mapSmoothTile->GetPlane(1)->GetPixels(), mapSmoothTile->GetPlane(2)->GetPixels() };
    int32_t* mrgb[3] = { mappedRGB->GetPlane(0)->GetPixels(), mappedRGB->GetPlane(1)->GetPixels(), mappedRGB->GetPlane(2)->GetPixels() };
    int32_t* rec[3] = { 0, 0, 0 };
    if (lastTestOutput) for (int n = 0; n < 3; n++) rec[n] = lastTestOutput->GetPlane(n)->GetPixels();
    lastError = yk_download_state(ctx, 0, smoothMap->GetPixels(), mst, mrgb, mipmapMask->GetPixels(), lastTestOutput ? rec : NULL);
}

// ---- host tails ------------------------------------------------------------------------------------------------------
void EncoderContext::SetCompressor(yk_compress_fn fn, void* user, int paletteMode) {
    FinishTails();
    compressFn = fn; compressUser = user;
    if (palette) yk_palette_destroy(palette);
    palette = yk_palette_create(paletteMode);
}

void EncoderContext::SetAsyncTails(bool on) {
    FinishTails();
    asyncTails = on;
    if (on && !worker.joinable()) worker = std::thread(&EncoderContext::workerLoop, this);
}

void EncoderContext::workerLoop() {
    for (;;) {
        std::function<void()> job;
        {
            std::unique_lock<std::mutex> lk(tailMutex);
            tailCv.wait(lk, [this] { return workerStop || !tailQueue.empty(); });
            if (tailQueue.empty()) return;              // stop requested and nothing left
            job = std::move(tailQueue.front());
            tailQueue.pop_front();
        }
        job();
        { std::lock_guard<std::mutex> lk(tailMutex); tailsInFlight--; }
        tailIdleCv.notify_all();
    }
}

void EncoderContext::postTail(std::function<void()> job) {
    if (!asyncTails) { job(); return; }
    { std::lock_guard<std::mutex> lk(tailMutex); tailQueue.push_back(std::move(job)); tailsInFlight++; }
    tailCv.notify_one();
}

void EncoderContext::FinishTails() {
    if (!worker.joinable()) return;
    std::unique_lock<std::mutex> lk(tailMutex);
    tailIdleCv.wait(lk, [this] { return tailsInFlight == 0; });
}

void EncoderContext::writeChunk(const std::vector<u8>& chunk) {        // one writer at a time: the worker, or the caller in the synchronous form
    if (outFile && !chunk.empty()) fwrite(chunk.data(), 1, chunk.size(), outFile);
}

void EncoderContext::WriteFileHeader() {
    if (!outFile || !original) return;
    const int w = original->GetWidth(), h = original->GetHeight(), a = original->HasAlpha() ? 1 : 0;
    postTail([this, w, h, a]() {
        std::vector<u8> chunk(16); size_t n = 0;
        if (yk_chunk_file_header(chunk.data(), chunk.size(), &n, w, h, a) == YK_OK) { chunk.resize(n); writeChunk(chunk); }
    });
}

void EncoderContext::WriteEndTag() {
    if (!outFile) return;
    postTail([this]() {
        std::vector<u8> chunk(8); size_t n = 0;
        if (yk_chunk_end(chunk.data(), chunk.size(), &n) == YK_OK) { chunk.resize(n); writeChunk(chunk); }
    });
}

void EncoderContext::GenerateDynamicTileChunk(u8* stream, int sizeStream) {
    if (!outFile || !compressFn || sizeStream <= 0) return;
    const std::vector<u8> idx(stream, stream + sizeStream), type(streamType, pType);
    const int cc = colorCompression1D, cr = rangeCompression1D;
    postTail([this, idx, type, cc, cr]() {
        std::vector<u8> chunk(idx.size() * 3 + type.size() * 2 + 4096);
        size_t n = 0;
        const int rc = yk_chunk_1dtl(chunk.data(), chunk.size(), &n, compressFn, compressUser, idx.data(), (int)idx.size(), type.data(), (int)type.size(), cc, cr);
        if (rc == YK_OK && n) { chunk.resize(n); writeChunk(chunk); } else if (rc) lastError = rc;
    });
}
