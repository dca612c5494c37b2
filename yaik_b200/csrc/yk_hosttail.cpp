// yaik_b200 — the host tails behind the analysis stage, plain C++ (no CUDA, no zstd):
//
//   yk_palette_*   PaletteCompressor (KLab/YAIK encoder/EncoderContext.cpp = "EC.cpp" 3209-3502): the delta code-book
//                  coder every rgbStream of a gradient pass goes through.  The reference keeps its code book in a
//                  100 000-entry global that it never clears, and FindCodeBook always scans 64 entries (EC.cpp:3216-3217,
//                  3248-3255): a call can emit an index left behind by an earlier, larger call.  Here
//                  the table is an object (yk_palette), grown as needed, with an explicit switch: YK_PALETTE_BUG_COMPATIBLE
//                  reproduces the reference's bytes call for call (stale entries included, from a fresh object = a
//                  fresh process), YK_PALETTE_DECODABLE only ever emits indices of the code book it writes.
//                  Work per colour is bounded: a 64-entry search window and hash look-ups instead of the reference's
//                  linear scans over the whole code book.
//   yk_chunk_*     the chunk serialisers of Convert(): file header, MIPM (EC.cpp:1367-1396), GTIL (4239-4350), 1DTL
//                  (8524-8576), PLNT (4515-4589), end tag (9779-9782); layouts from include/YAIK_private.h:96-118,
//                  172-197, 290-300, 347-356.  Entropy coding goes through a caller-supplied callback (the reference
//                  links zstd 1.3.4; this library links none).  Reference quirks are kept: bbox.h = maxY - minX
//                  (EC.cpp:4258), a GTIL chunk only when the pass emitted colours and its box is not empty (4239).
//                  Bytes the reference leaves uninitialised (HeaderGradientTile::version, MipmapHeader::streamSize,
//                  struct padding) are written as zero.
#include "../../include/yaik_b200.h"

#include <algorithm>
#include <stdint.h>
#include <string.h>
#include <unordered_map>
#include <vector>

namespace {

struct Delta { int dr, dg, db; };
static inline uint32_t key_of(int dr, int dg, int db) { return (uint32_t)(dr + 256) | ((uint32_t)(dg + 256) << 10) | ((uint32_t)(db + 256) << 20); }

}  // namespace

struct yk_palette {
    int mode;
    std::vector<Delta> table;       // the reference's CodeRGB[]: entries past the current call's code book keep older contents
};

extern "C" yk_palette* yk_palette_create(int mode) {
    if (mode != YK_PALETTE_BUG_COMPATIBLE && mode != YK_PALETTE_DECODABLE) return nullptr;
    yk_palette* p = new yk_palette();
    p->mode = mode;
    p->table.assign(64, Delta{ 0, 0, 0 });      // a zero-initialised global
    return p;
}
extern "C" void yk_palette_destroy(yk_palette* p) { delete p; }
extern "C" void yk_palette_reset(yk_palette* p) { if (p) p->table.assign(64, Delta{ 0, 0, 0 }); }

extern "C" int yk_palette_compress(yk_palette* p, const uint8_t* in, int size, uint8_t* out, int outCap, int* outBytes) {
    if (!p || !in || !out || !outBytes || size < 3 || outCap < 0) return YK_ERR_ARG;
    *outBytes = 0;
    const int n = size / 3;
    int w = 0;
    auto put = [&](int v) -> bool { if (w < outCap) { out[w++] = (uint8_t)v; return true; } return false; };

    // ---- phase 1 (EC.cpp:3285-3311): per colour the delta to the closest of the 64 colours before it (first minimum in
    // stream order); deltas are registered in order of first appearance with a reference count
    struct Entry { Delta d; int ref; };
    std::vector<Entry> book;
    std::unordered_map<uint32_t, int> where;
    book.push_back(Entry{ Delta{ 0, 0, 0 }, 0 });                  // "special null code, better be at top"
    where.emplace(key_of(0, 0, 0), 0);
    for (int i = 1; i < n; i++) {
        const uint8_t* pix = in + 3 * i;
        const int start = i - 64 < 0 ? 0 : i - 64;
        int best = 999999999, bR = 0, bG = 0, bB = 0;
        for (int prev = start; prev < i; prev++) {
            const int dR = pix[0] - in[3 * prev], dG = pix[1] - in[3 * prev + 1], dB = pix[2] - in[3 * prev + 2];
            const int dist = dR * dR + dG * dG + dB * dB;
            if (dist < best) { best = dist; bR = dR; bG = dG; bB = dB; }
        }
        const uint32_t k = key_of(bR, bG, bB);
        auto it = where.find(k);
        if (it != where.end()) book[it->second].ref++;
        else { where.emplace(k, (int)book.size()); book.push_back(Entry{ Delta{ bR, bG, bB }, 0 }); }
    }
    // EC.cpp:3317: qsort by descending count, entry 0 stays.  The comparator calls equal counts equal, so their order is
    // the C library's; glibc's qsort is a merge sort (stable), which is what the compiled reference shows: first
    // appearance breaks ties.
    std::stable_sort(book.begin() + 1, book.end(), [](const Entry& a, const Entry& b) { return a.ref > b.ref; });
    const int count = (int)book.size();
    if ((int)p->table.size() < count) p->table.resize(count, Delta{ 0, 0, 0 });
    for (int i = 0; i < count; i++) p->table[i] = book[i].d;
    const int finalCount = count > 128 ? 128 : count;

    // FindCodeBook (EC.cpp:3248-3255): the first of the 64 leading table entries that matches - stale ones included in
    // the reference; only the code book just written in decodable mode
    const int visible = p->mode == YK_PALETTE_BUG_COMPATIBLE ? 64 : (finalCount < 64 ? finalCount : 64);
    std::unordered_map<uint32_t, int> first;
    first.reserve(128);
    for (int i = visible - 1; i >= 0; i--) first[key_of(p->table[i].dr, p->table[i].dg, p->table[i].db)] = i;
    auto find = [&](int dR, int dG, int dB) -> int { auto it = first.find(key_of(dR, dG, dB)); return it == first.end() ? -1 : it->second; };

    // ---- header (EC.cpp:3324-3339): code book, first colour
    if (!put(finalCount)) return YK_ERR_CAPACITY;
    for (int i = 0; i < finalCount; i++)
        if (!put(p->table[i].dr) || !put(p->table[i].dg) || !put(p->table[i].db)) return YK_ERR_CAPACITY;
    if (!put(in[0]) || !put(in[1]) || !put(in[2])) return YK_ERR_CAPACITY;

    // ---- phase 2 (EC.cpp:3348-3486)
    for (int i = 1; i < n; i++) {
        const uint8_t* pix = in + 3 * i;
        const int start = i - 65 < 0 ? 0 : i - 65;
        int bestIndex = 999, bestDistance = 0;
        bool done = false;
        for (int prev = i - 1; prev >= start; prev--) {
            const int index = find(pix[0] - in[3 * prev], pix[1] - in[3 * prev + 1], pix[2] - in[3 * prev + 2]);
            if (index < 0) continue;
            if (prev == i - 1) {                                   // code book delta from the previous colour
                if (!put(index & 0x7F)) return YK_ERR_CAPACITY;
                done = true;
                break;
            }
            const int distance = (i - prev) - 2;                   // smallest index among the reachable colours
            if (distance < 64 && index < bestIndex) { bestIndex = index; bestDistance = distance; done = true; }
        }
        if (bestIndex != 999) {
            if (!put(0xC0 | (bestDistance & 0x3F)) || !put(bestIndex & 0x7F)) return YK_ERR_CAPACITY;
        }
        if (!done) {
            const int dR = pix[0] - pix[-3], dG = pix[1] - pix[-2], dB = pix[2] - pix[-1];
            const int mask = (dR ? 1 : 0) | (dG ? 2 : 0) | (dB ? 4 : 0);
            if (dR >= -128 && dR <= 127 && dG >= -128 && dG <= 127 && dB >= -128 && dB <= 127) {
                if (!put(0x80 | mask)) return YK_ERR_CAPACITY;     // relative to the previous colour, component mask
                if (dR && !put(dR)) return YK_ERR_CAPACITY;
                if (dG && !put(dG)) return YK_ERR_CAPACITY;
                if (dB && !put(dB)) return YK_ERR_CAPACITY;
            } else {
                if (!put(0x88 | mask)) return YK_ERR_CAPACITY;     // absolute components
                if (dR && !put(pix[0])) return YK_ERR_CAPACITY;
                if (dG && !put(pix[1])) return YK_ERR_CAPACITY;
                if (dB && !put(pix[2])) return YK_ERR_CAPACITY;
            }
        }
    }
    *outBytes = w;
    return YK_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// chunk serialisers
namespace {

struct Writer {
    uint8_t* dst; size_t cap, n;
    bool ok;
    Writer(uint8_t* d, size_t c) : dst(d), cap(c), n(0), ok(true) {}
    void bytes(const void* p, size_t k) { if (n + k > cap) { ok = false; return; } if (k) memcpy(dst + n, p, k); n += k; }
    void zeros(size_t k) { if (n + k > cap) { ok = false; return; } memset(dst + n, 0, k); n += k; }
    void u8v(unsigned v) { uint8_t b = (uint8_t)v; bytes(&b, 1); }
    void u16(unsigned v) { uint8_t b[2] = { (uint8_t)v, (uint8_t)(v >> 8) }; bytes(b, 2); }
    void s16(int v) { u16((unsigned)v & 0xFFFFu); }
    void u32(uint32_t v) { uint8_t b[4] = { (uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24) }; bytes(b, 4); }
    void tag(const char* t) { bytes(t, 4); }
};

// ZSTD_compress through the caller's callback into a scratch buffer sized like the reference's
static bool squeeze(yk_compress_fn fn, void* user, const void* src, size_t srcBytes, size_t dstCap, int level, std::vector<uint8_t>& out) {
    out.assign(dstCap ? dstCap : 1, 0);
    const size_t r = fn(user, out.data(), out.size(), src, srcBytes, level);
    if (r == 0 || r > out.size()) return false;
    out.resize(r);
    return true;
}

static int finish(Writer& w, size_t* n) {
    if (!w.ok) return YK_ERR_CAPACITY;
    if (n) *n = w.n;
    return YK_OK;
}

}  // namespace

// FileHeader (YAIK_private.h:96-105, EC.cpp:9007-9016)
extern "C" int yk_chunk_file_header(uint8_t* dst, size_t cap, size_t* n, int width, int height, int hasAlpha) {
    if (!dst || width < 0 || height < 0 || width > 65535 || height > 65535) return YK_ERR_ARG;
    Writer w(dst, cap);
    w.tag("YAIK"); w.u16(1); w.u16((unsigned)width); w.u16((unsigned)height); w.u16(hasAlpha ? 1 : 0);
    return finish(w, n);
}

// end tag (EC.cpp:9779-9782)
extern "C" int yk_chunk_end(uint8_t* dst, size_t cap, size_t* n) {
    if (!dst) return YK_ERR_ARG;
    Writer w(dst, cap);
    w.u32(0xDEADBEEFu);
    return finish(w, n);
}

// 'MIPM' (EC.cpp:1367-1396): HeaderBase, MipmapHeader { bbox in 16x16 tiles, streamSize (not set by the reference), version 1,
// mipmapLevel 4 }, the 1-bit tile bitmap as it is, zero padding to 4 bytes
extern "C" int yk_chunk_mipm(uint8_t* dst, size_t cap, size_t* n, const int bboxTiles[4], const uint8_t* bitmap, int bitmapBytes) {
    if (!dst || !bboxTiles || bitmapBytes < 0 || (bitmapBytes && !bitmap)) return YK_ERR_ARG;
    Writer w(dst, cap);
    const uint32_t base = 16u + (uint32_t)bitmapBytes, length = (base + 3u) & ~3u;
    w.tag("MIPM"); w.u32(length);
    for (int k = 0; k < 4; k++) w.s16(bboxTiles[k]);
    w.u32(0); w.u8v(1); w.u8v(4); w.zeros(2);
    w.bytes(bitmap, (size_t)bitmapBytes);
    w.zeros(length - base);
    return finish(w, n);
}

// 'GTIL' (EC.cpp:4239-4350): written only when the pass emitted colours and its box is not empty.  Streams: ZSTD-18 of the
// whole swizzled accept bitmap, ZSTD-18 of PaletteCompressor(rgbStream).
extern "C" int yk_chunk_gtil(uint8_t* dst, size_t cap, size_t* n, yk_palette* palette, yk_compress_fn compress, void* user,
                             int shX, int shY, int planeBits, const int bbox[4], const uint8_t* bitmap, int bitmapBytes,
                             const uint8_t* rgb, int rgbBytes, int colorCompression) {
    if (!dst || !palette || !compress || !bbox || bitmapBytes < 0 || rgbBytes < 0 || (bitmapBytes && !bitmap) || (rgbBytes && !rgb)) return YK_ERR_ARG;
    if (n) *n = 0;
    const int minX = bbox[0], minY = bbox[1], maxX = bbox[2], maxY = bbox[3];
    if (!(maxX > minX && maxY > minY && rgbBytes > 0)) return YK_OK;           // EC.cpp:4239: no chunk
    std::vector<uint8_t> zBitmap, zRgb, pal((size_t)rgbBytes * 3);
    const size_t capB = (size_t)bitmapBytes * 2 < 1000 ? 1000 : (size_t)bitmapBytes * 2;          // CompressStream, EC.cpp:3692-3708
    if (!squeeze(compress, user, bitmap, (size_t)bitmapBytes, capB, 18, zBitmap)) return YK_ERR_STATE;
    int palBytes = 0;
    const int rc = yk_palette_compress(palette, rgb, rgbBytes, pal.data(), (int)pal.size(), &palBytes);
    if (rc) return rc;
    const size_t capR = (size_t)palBytes * 2 < 1000 ? 1000 : (size_t)palBytes * 2;
    if (!squeeze(compress, user, pal.data(), (size_t)palBytes, capR, 18, zRgb)) return YK_ERR_STATE;
    Writer w(dst, cap);
    const uint32_t base = 28u + (uint32_t)zBitmap.size() + (uint32_t)zRgb.size(), length = (base + 3u) & ~3u;
    w.tag("GTIL"); w.u32(length);
    w.s16(minX); w.s16(minY); w.s16(maxX - minX); w.s16(maxY - minX);          // bbox.h uses minX: EC.cpp:4258
    w.u32((uint32_t)zBitmap.size()); w.u32((uint32_t)zRgb.size()); w.u32((uint32_t)palBytes); w.u32((uint32_t)rgbBytes);
    w.u8v((unsigned)colorCompression); w.u8v(0); w.u8v((unsigned)(shX | (shY << 3))); w.u8v((unsigned)planeBits);
    w.bytes(zBitmap.data(), zBitmap.size()); w.bytes(zRgb.data(), zRgb.size());
    w.zeros(length - base);
    return finish(w, n);
}

// '1DTL' (GenerateDynamicTileChunk, EC.cpp:8524-8576): Header1D, ZSTD-18 of the type stream, ZSTD-18 of the index stream
// (R, G, B concatenated by the caller, EC.cpp:9451-9465); nothing when the index stream is empty
extern "C" int yk_chunk_1dtl(uint8_t* dst, size_t cap, size_t* n, yk_compress_fn compress, void* user, const uint8_t* idx, int idxBytes,
                             const uint8_t* type, int typeBytes, int compressionColor, int compressionRange) {
    if (!dst || !compress || idxBytes < 0 || typeBytes < 0 || (idxBytes && !idx) || (typeBytes && !type)) return YK_ERR_ARG;
    if (n) *n = 0;
    if (idxBytes <= 0) return YK_OK;
    std::vector<uint8_t> zIdx, zType;
    if (!squeeze(compress, user, idx, (size_t)idxBytes, (size_t)idxBytes * 2, 18, zIdx)) return YK_ERR_STATE;
    if (!squeeze(compress, user, type, (size_t)typeBytes, (size_t)idxBytes, 18, zType)) return YK_ERR_STATE;
    Writer w(dst, cap);
    const uint32_t base = 20u + (uint32_t)zIdx.size() + (uint32_t)zType.size(), length = (base + 3u) & ~3u;
    w.tag("1DTL"); w.u32(length);
    w.u32((uint32_t)zIdx.size()); w.u32((uint32_t)idxBytes); w.u32((uint32_t)zType.size()); w.u32((uint32_t)typeBytes);
    w.u8v((unsigned)compressionColor); w.u8v((unsigned)compressionRange); w.u8v(0); w.zeros(1);
    w.bytes(zType.data(), zType.size()); w.bytes(zIdx.data(), zIdx.size());
    w.zeros(length - base);
    return finish(w, n);
}

// 'PLNT' (EC.cpp:4515-4589): PlaneTile { constraint box, sizes, version 1, format }, ZSTD-21 of the u16 tile definitions,
// ZSTD-21 of the nibble bytes (an odd nibble count is closed with a zero nibble)
extern "C" int yk_chunk_plnt(uint8_t* dst, size_t cap, size_t* n, yk_compress_fn compress, void* user, const int constraint[4],
                             const uint16_t* defs, int nDefs, const uint8_t* nibbles, int nNibbles, int planeType, int halfX, int halfY) {
    if (!dst || !compress || !constraint || nDefs < 0 || nNibbles < 0 || (nDefs && !defs) || (nNibbles && !nibbles) || planeType < 0 || planeType > 2) return YK_ERR_ARG;
    const size_t nibBytes = ((size_t)nNibbles + 1) >> 1;
    const size_t dw = (size_t)(constraint[2] + 7) / 8 + 1, dh = (size_t)(constraint[3] + 7) / 8 + 1;
    std::vector<uint8_t> zDefs, zNib;
    if (!squeeze(compress, user, defs, (size_t)nDefs * 2, dw * dh * 3 + 1024, 21, zDefs)) return YK_ERR_STATE;
    if (!squeeze(compress, user, nibbles, nibBytes, dw * dh * 64 + 1024, 21, zNib)) return YK_ERR_STATE;
    Writer w(dst, cap);
    const uint32_t base = 24u + (uint32_t)zDefs.size() + (uint32_t)zNib.size(), length = (base + 3u) & ~3u;
    w.tag("PLNT"); w.u32(length);
    for (int k = 0; k < 4; k++) w.s16(constraint[k]);
    w.u32((uint32_t)zDefs.size()); w.u32((uint32_t)zNib.size()); w.u32((uint32_t)nibBytes);
    w.u8v(1); w.u8v((unsigned)((planeType << 2) | (halfX ? 1 : 0) | (halfY ? 2 : 0))); w.zeros(2);
    w.bytes(zDefs.data(), zDefs.size()); w.bytes(zNib.data(), zNib.size());
    w.zeros(length - base);
    return finish(w, n);
}
