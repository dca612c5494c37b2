// TEST INFRASTRUCTURE ONLY — see cuda_emu.h.
#include "cuda_emu.h"

namespace yk_emu {
Cta* g_cta = nullptr;
dim3 g_blockDim, g_gridDim;
thread_local uint3 t_threadIdx, t_blockIdx;
thread_local int t_lane, t_warp;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    unsigned nthreads = block.x * block.y * block.z;
    if (nthreads % 32 != 0) { fprintf(stderr, "yk_emu: block size %u not a multiple of 32\n", nthreads); abort(); }
    Cta cta;
    cta.bar.reset(new std::barrier<>(nthreads));
    cta.warps.resize(nthreads / 32);
    for (auto& w : cta.warps) w.bar.reset(new std::barrier<>(32));
    cta.dynSmem.resize(smem + 16);
    g_cta = &cta; g_blockDim = block; g_gridDim = grid;
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; t++)
        pool.emplace_back([&, t]() {
            t_threadIdx.x = t % block.x; t_threadIdx.y = (t / block.x) % block.y; t_threadIdx.z = t / (block.x * block.y);
            t_lane = t % 32; t_warp = t / 32;
            for (unsigned bz = 0; bz < grid.z; bz++)
                for (unsigned by = 0; by < grid.y; by++)
                    for (unsigned bx = 0; bx < grid.x; bx++) {
                        t_blockIdx.x = bx; t_blockIdx.y = by; t_blockIdx.z = bz;
                        body();
                        cta.bar->arrive_and_wait();     // next CTA reuses the function-static "shared memory"
                    }
        });
    for (auto& th : pool) th.join();
    g_cta = nullptr;
}
}  // namespace yk_emu
