// Developer microbenchmark: per-SM throughput of 2-D TMA box loads (cp.async.bulk.tensor) of int32 planes by box shape,
// one issuing thread per CTA, NBUF boxes in flight.  nvcc -arch=sm_100a -o tma_box tma_box.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int NBUF>
__global__ void k(const __grid_constant__ CUtensorMap tm, int boxBytes, int nIter, int nbx, int nby, int bw, int bh, unsigned long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar[NBUF];
    if (threadIdx.x == 0) {
        for (int i = 0; i < NBUF; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int stageBytes = (boxBytes + 127) / 128 * 128;
        long long t0 = clock64();
        for (int it = 0; it < nIter + NBUF; it++) {
            const int i = it % NBUF;
            if (it >= NBUF) {   // wait for the box issued NBUF iterations ago
                unsigned ok = 0; const unsigned par = ((it / NBUF) - 1) & 1;
                do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(&bar[i])), "r"(par) : "memory"); } while (!ok);
            }
            if (it < nIter) {
                const int id = (blockIdx.x * 977 + it * 31) % (nbx * nby);
                const int x = (id % nbx) * bw, y = (id / nbx) * bh;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bar[i])), "r"(boxBytes) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             :: "r"(s32(smem + i * stageBytes)), "l"(&tm), "r"(x), "r"(y), "r"(s32(&bar[i])) : "memory");
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}
int main() {
    const int W = 8192, H = 8192;
    int32_t* d; CK(cudaMalloc(&d, (size_t)W * H * 4)); CK(cudaMemset(d, 1, (size_t)W * H * 4));
    unsigned long long* cyc; CK(cudaMalloc(&cyc, 148 * 8));
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn; cudaDriverEntryPointQueryResult q; CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    const int shapes[][2] = { {68, 17}, {132, 17}, {20, 17} };
    for (auto& sh : shapes) {
        CUtensorMap tm; cuuint64_t dims[2] = { W, H }, str[1] = { (cuuint64_t)W * 4 }; cuuint32_t box[2] = { (cuuint32_t)sh[0], (cuuint32_t)sh[1] }, es[2] = { 1, 1 };
        CUresult r = ((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode %dx%d failed %d\n", sh[0], sh[1], (int)r); continue; }
        const int boxBytes = sh[0] * sh[1] * 4, nIter = 400;
        const int bwAl = (sh[0] + 63) / 64 * 64;
        auto run = [&](auto kern, int NB) -> int {
            const int smem = NB * ((boxBytes + 127) / 128 * 128);
            if (smem > 220 * 1024) { printf("%dx%d NB=%d: too big\n", sh[0], sh[1], NB); return 0; }
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            for (int rep = 0; rep < 2; rep++) {
                kern<<<148, 32, smem>>>(tm, boxBytes, nIter, W / bwAl - 1, H / (sh[1] + 15) - 1, bwAl, sh[1] + 15 & ~15, cyc);
                CK(cudaDeviceSynchronize());
            }
            std::vector<unsigned long long> h(148); CK(cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost));
            double avg = 0; for (auto v : h) avg += v; avg /= 148;
            printf("box %3d x %2d int32 (%6d B, rows of %4d B) NB=%2d: %7.0f cycles per box, %5.2f B/cycle/SM, %6.0f GB/s chip at 1.9 GHz\n",
                   sh[0], sh[1], boxBytes, sh[0] * 4, NB, avg / nIter, boxBytes / (avg / nIter), boxBytes / (avg / nIter) * 148 * 1.9);
            return 0;
        };
        run(k<2>, 2); run(k<4>, 4); run(k<8>, 8); run(k<16>, 16); run(k<32>, 32);
    }
    return 0;
}
