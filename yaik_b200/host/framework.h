// Host-side mirror of the reference's Plane / Image containers (KLab/YAIK encoder/framework.h:74-225) — same names,
// same int32 row-major storage, same clamping accessor — so code written against the reference's framework keeps
// compiling against this path.  Only what the encoder-analysis stage touches is provided (no PNG I/O, no resampling).
// Plane storage is pinned host memory (yk_host_alloc) so EncoderContext::SetImageToEncode uploads at full PCIe rate.
#pragma once
#include <stdint.h>
#include <string.h>
#include "../../include/yaik_b200.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef int16_t s16;

struct BoundingBox { s16 x, y, w, h; };             // include/YAIK_private.h:15-20
enum EDownSample { NEAREST_TL, NEAREST_BR, AVERAGE_BOX, MAX_BOX, MIN_BOX };     // encoder/framework.h:60-66

class Plane {                                       // encoder/framework.h:74-127
public:
    // pinned = true: page-locked storage (image planes that get uploaded); false: ordinary memory (state planes)
    Plane(int w_, int h_, bool pinned_ = true) : w(w_), h(h_), pinned(pinned_) {
        pixels = pinned ? (int*)yk_host_alloc((size_t)w * h * sizeof(int)) : new int[(size_t)w * h];
    }
    ~Plane() { if (pinned) yk_host_free(pixels); else delete[] pixels; }
    inline int  GetWidth() { return w; }
    inline int  GetHeight() { return h; }
    inline int* GetPixels() { return pixels; }
    inline int  GetIndex(int x, int y) { return x + y * w; }
    inline BoundingBox GetRect() { BoundingBox r; r.x = 0; r.y = 0; r.w = (s16)w; r.h = (s16)h; return r; }
    void SetPixel(int x, int y, int v) { pixels[x + y * w] = v; }
    void Fill(BoundingBox& rect, int value) {       // encoder/Plane.cpp:371-377
        for (int y = rect.y; y < rect.y + rect.h; y++)
            for (int x = rect.x; x < rect.x + rect.w; x++) pixels[x + y * w] = value;
    }
    void Clear() { memset(pixels, 0, (size_t)w * h * sizeof(int)); }
    int GetPixelValue(int x, int y, bool& isOutside) {      // clamping accessor, encoder/framework.h:116-121
        isOutside = false;
        if (x < 0 || x >= w) { isOutside = true; if (x >= w) x = w - 1; if (x < 0) x = 0; }
        if (y < 0 || y >= h) { isOutside = true; if (y >= h) y = h - 1; if (y < 0) y = 0; }
        return pixels[x + y * w];
    }
private:
    int* pixels;
    int  w, h;
    bool pinned;
};
typedef Plane* TPlane;

class Image {                                       // encoder/framework.h:137-225
    Plane* planes[4];
    int planeCount, w, h;
    Image() : planeCount(0), w(0), h(0) { planes[0] = planes[1] = planes[2] = planes[3] = NULL; }
public:
    ~Image() { for (int n = 0; n < 4; n++) delete planes[n]; }
    static Image* CreateImage(int w, int h, int channelCount, bool fill, bool pinned = true) {     // encoder/Image.cpp:10-22
        Image* i = new Image();
        i->w = w; i->h = h; i->planeCount = channelCount;
        for (int n = 0; n < channelCount; n++) { i->planes[n] = new Plane(w, h, pinned); if (fill) i->planes[n]->Clear(); }
        return i;
    }
    inline int GetWidth() { return w; }
    inline int GetHeight() { return h; }
    inline TPlane GetPlane(int index) { return planes[index]; }
    bool HasAlpha() { return planeCount == 4; }
    void Clear() { for (int n = 0; n < 4; n++) if (planes[n]) planes[n]->Clear(); }
    void SetPixel(int x, int y, int r, int g, int b) { planes[0]->SetPixel(x, y, r); planes[1]->SetPixel(x, y, g); planes[2]->SetPixel(x, y, b); }
};
