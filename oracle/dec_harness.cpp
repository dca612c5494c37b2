// TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
//
// Drives the UNMODIFIED reference decoder library (KLab/YAIK decoder/*.cpp, compiled by oracle/Makefile target `dec`
// from the sources where they lie under /root/reference into oracle/_ref/) through its public API (include/YAIK.h:
// YAIK_Init, YAIK_DecodeImagePre, YAIK_DecodeImage) on one .yaik stream and writes the decoded image.
// Used by tests/ for the decoder-side round trip: the streams of the CUDA path, serialised by the
// product's chunk writers, must decode — with the reference's own decoder — to the same image as the reference
// encoder's file.
//
// usage: yaik_dec in.yaik out.raw      out: int32 w, h, channels, then h*w*channels bytes (interleaved RGB[A])
#include "../include/YAIK.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <vector>

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s in.yaik out.raw\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    void* data = NULL;
    if (posix_memalign(&data, 64, (size_t)n + 64)) return 1;          // "the stream must be aligned to 4 bytes"
    memset(data, 0, (size_t)n + 64);
    if (fread(data, 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read\n"); return 1; }
    fclose(f);
    FILE* out = fopen(argv[2], "wb");
    if (!out) { perror(argv[2]); return 1; }
    // the DEVEL build of the decoder writes debug PNGs into the cwd and printf()s
    char tmpl[] = "/tmp/yaik_dec_XXXXXX";
    char* dir = mkdtemp(tmpl);
    if (!dir || chdir(dir) != 0) { perror("mkdtemp/chdir"); return 1; }
    if (!freopen("/dev/null", "w", stdout)) return 1;

    YAIK_LIB lib = YAIK_Init(1, NULL);
    if (!lib) { fprintf(stderr, "YAIK_Init failed\n"); return 1; }
    YAIK_SDecodedImage info;
    memset(&info, 0, sizeof info);
    if (!YAIK_DecodeImagePre(lib, data, (u32)n, &info)) { fprintf(stderr, "YAIK_DecodeImagePre failed: error %d\n", (int)YAIK_GetErrorCode()); return 3; }
    const int ch = info.hasAlpha ? 4 : 3;
    std::vector<u8> img((size_t)info.width * info.height * ch, 0);
    info.outputImage = img.data();
    info.outputImageStride = info.width * ch;
    const bool ok = YAIK_DecodeImage(data, (u32)n, &info);
    if (!ok) { fprintf(stderr, "YAIK_DecodeImage failed: error %d\n", (int)YAIK_GetErrorCode()); return 4; }
    int hdr[3] = { (int)info.width, (int)info.height, ch };
    fwrite(hdr, 4, 3, out);
    fwrite(img.data(), 1, img.size(), out);
    fclose(out);
    YAIK_Release(lib);
    char cmd[256]; snprintf(cmd, sizeof cmd, "rm -rf '%s'", dir);
    if (system(cmd) != 0) {}
    return 0;
}
