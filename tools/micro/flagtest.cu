// Developer probe: can a one-thread kernel spinning on a flag in stream A be released by work enqueued later in stream B
// (memset / device-to-device copy / a one-thread store kernel)?  usage: flagtest <variant 0..3> [preload]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k_wait(const volatile unsigned* f, unsigned v) { while ((int)(*f - v) < 0) __nanosleep(20); __threadfence_system(); }
__global__ void k_set(volatile unsigned* f, unsigned v) { __threadfence_system(); *f = v; }
__global__ void k_big(int* p) { p[blockIdx.x * blockDim.x + threadIdx.x] += 1; }
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    if (argc > 2) { cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_set); cudaFuncGetAttributes(&fa, k_wait); cudaFuncGetAttributes(&fa, k_big); }   // force-load (lazy module loading)
    cudaStream_t a, b; cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    unsigned* flag; char *x, *y; int* big;
    cudaMalloc(&flag, 64); cudaMalloc(&x, 1 << 20); cudaMalloc(&y, 1 << 20); cudaMalloc(&big, 148 * 768 * 4);
    cudaMemset(flag, 0, 64); cudaDeviceSynchronize();
    if (variant >= 2) cudaMemsetAsync(x, 0, 1 << 20, a);
    k_wait<<<1, 1, 0, a>>>(flag, 1);
    if (variant >= 3) k_big<<<148, 768, 196 * 1024 > 48 * 1024 ? 0 : 0, a>>>(big);
    if (variant >= 1) { cudaMemsetAsync(y, 0, 1 << 20, b); cudaMemcpyAsync(x + 4096, y, 1024, cudaMemcpyDefault, b); }
    k_set<<<1, 1, 0, b>>>(flag, 1);
    cudaError_t e = cudaStreamSynchronize(a);
    printf("variant %d: %s\n", variant, cudaGetErrorString(e));
    return 0;
}
