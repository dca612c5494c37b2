// yaik_b200 — host side of the packed upload: the reference hands over `Plane::GetPixels()` buffers (int32 per sample,
// encoder/framework.h:74-127) whose samples are 0..255 on this path, so yk_set_image packs them to bytes on the host
// (a quarter of the PCIe traffic), checking the range while it does (a sample outside 0..255 is YK_ERR_RANGE, as in the
// device path).  Plain C++ with an AVX2 body selected at run time; a small persistent thread pool splits the rows.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

unsigned pack_row_scalar(const int32_t* src, uint8_t* dst, int n) {
    unsigned bad = 0;
    for (int i = 0; i < n; i++) { bad |= (unsigned)src[i]; dst[i] = (uint8_t)src[i]; }
    return bad;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) unsigned pack_row_avx2(const int32_t* src, uint8_t* dst, int n) {
    __m256i acc = _mm256_setzero_si256();
    const __m256i perm = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    // the packed bytes are only read back by the DMA engine: write them around the cache when the row is 32-byte aligned
    const bool stream = (((uintptr_t)dst) & 31u) == 0;
    int i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 8));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 16)), d = _mm256_loadu_si256((const __m256i*)(src + i + 24));
        acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(c, d)));
        // low bytes only (the range check catches anything else): 32 -> 16 -> 8 bits, lanes put back in order
        const __m256i ab = _mm256_packus_epi32(_mm256_and_si256(a, _mm256_set1_epi32(255)), _mm256_and_si256(b, _mm256_set1_epi32(255)));
        const __m256i cd = _mm256_packus_epi32(_mm256_and_si256(c, _mm256_set1_epi32(255)), _mm256_and_si256(d, _mm256_set1_epi32(255)));
        const __m256i abcd = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(ab, cd), perm);
        if (stream) _mm256_stream_si256((__m256i*)(dst + i), abcd); else _mm256_storeu_si256((__m256i*)(dst + i), abcd);
    }
    if (stream) _mm_sfence();
    unsigned bad = 0;
    alignas(32) unsigned tmp[8];
    _mm256_store_si256((__m256i*)tmp, acc);
    for (int k = 0; k < 8; k++) bad |= tmp[k];
    for (; i < n; i++) { bad |= (unsigned)src[i]; dst[i] = (uint8_t)src[i]; }
    return bad;
}
#endif

typedef unsigned (*PackRowFn)(const int32_t*, uint8_t*, int);
PackRowFn pick_pack_row() {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2")) return pack_row_avx2;
#endif
    return pack_row_scalar;
}

// persistent workers: run(f) calls f(part, nParts) on every worker and the caller, returns when all are done
class Pool {
public:
    explicit Pool(int n) : n_(n < 1 ? 1 : n) {
        for (int i = 1; i < n_; i++) th_.emplace_back([this, i] { loop(i); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return n_; }
    void run(const std::function<void(int, int)>& f) {
        { std::lock_guard<std::mutex> l(m_); job_ = &f; pending_ = n_ - 1; gen_++; }
        cv_.notify_all();
        f(0, n_);
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return pending_ == 0; });
        job_ = nullptr;
    }
private:
    void loop(int id) {
        unsigned long seen = 0;
        for (;;) {
            const std::function<void(int, int)>* f;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                f = job_;
            }
            if (f) (*f)(id, n_);
            { std::lock_guard<std::mutex> l(m_); if (--pending_ == 0) done_.notify_all(); }
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)>* job_ = nullptr;
    int pending_ = 0;
    unsigned long gen_ = 0;
    bool stop_ = false;
};

}  // namespace

struct YkHostPacker {
    Pool pool;
    PackRowFn row;
    explicit YkHostPacker(int threads) : pool(threads), row(pick_pack_row()) {}
};

YkHostPacker* yk_hostpack_create(int threads) {
    if (threads <= 0) {
        unsigned hc = std::thread::hardware_concurrency();
        threads = (int)(hc ? hc : 4);
        if (threads > 8) threads = 8;
    }
    return new YkHostPacker(threads);
}
void yk_hostpack_destroy(YkHostPacker* p) { delete p; }
int yk_hostpack_threads(const YkHostPacker* p) { return p->pool.size(); }

// int32 plane (pitch == w) -> bytes with row pitch `pitch` (>= w; the padding is zeroed).  Returns the OR of all samples.
unsigned yk_hostpack_plane(YkHostPacker* p, const int32_t* src, uint8_t* dst, int w, int h, size_t pitch) {
    std::atomic<unsigned> bad(0);
    p->pool.run([&](int part, int parts) {
        const int y0 = (int)((long long)h * part / parts), y1 = (int)((long long)h * (part + 1) / parts);
        unsigned b = 0;
        for (int y = y0; y < y1; y++) {
            b |= p->row(src + (size_t)y * w, dst + (size_t)y * pitch, w);
            if (pitch > (size_t)w) memset(dst + (size_t)y * pitch + w, 0, pitch - w);
        }
        bad.fetch_or(b);
    });
    return bad.load();
}
