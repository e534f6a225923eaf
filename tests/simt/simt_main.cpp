/* simt_main.cpp -- TEST INFRASTRUCTURE: the product's megakernel source compiled for the host over cuda_shim.h.
 * Built per (driver, knobs, SDF unit) by tests/test_simt_emulation.py:
 *   g++ -std=c++20 -O1 -ffp-contract=off -fno-fast-math -mfma -DPT_SCHED=.. [-DPT_HAS_SDF=1 ..] simt_main.cpp
 *       ../../pathtracer_b200/csrc/pt_prepare.cpp [sdf_unit.o] -shared -o simt_<tag>.so
 * simt_dispatch() = pt_dispatch / pt_dispatch_sum of libpt_cuda on the emulated device: the same host-side preparation
 * (pt_prepare.cpp), the same grid, the kernel entry the JIT would define (PT_DEFINE_RENDER_KERNEL). */
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "cuda_shim.h"

thread_local SimtDim3 threadIdx, blockIdx;
SimtDim3 gridDim, blockDim;
thread_local SimtWarp* simt_warp = nullptr;
thread_local SimtBlock* simt_block = nullptr;
thread_local unsigned simt_phase = 0;

#if defined(PT_HAS_SDF) && PT_HAS_SDF
/* the generated SDF unit is a separate object with C linkage on the host (PT_SDF_ENTRY) */
extern "C" float pt_sdf_dispatch(float px, float py, float pz, unsigned set1, unsigned set2, unsigned set3, unsigned set4);
extern "C" float pt_sdfmaterial_dispatch(float px, float py, float pz, unsigned set1, unsigned set2, unsigned set3, unsigned set4);
#endif

#ifdef PT_STATS
extern "C" { unsigned long long pt_stats[16] = {0ull}; } /* the drivers' scheduling statistics (PT_STAT) */
#endif

#define PT_KERNEL_NS ptk_emu
#include "pt_internal.h"
#include "pt_kernel.cuh"

PT_DEFINE_RENDER_KERNEL(pt_render_emu)
#if PT_PREGEN
PT_DEFINE_GEN_KERNEL(pt_gen_emu)
#endif
#if PT_RESOLVE
PT_DEFINE_RESOLVE_KERNEL(pt_resolve_emu)
#endif

extern "C" int simt_sched(void) { return PT_SCHED; }

/* accum_mode 0: pt_dispatch (params carry frame / currentSamples); 2: pt_dispatch_sum(first, n).  image: W*H float4.
 * persistent_ctas: grid of the tile-streaming driver (PT_SCHED=6); ignored otherwise. */
/* only blocks bx0 <= bx < bx0 + nbx, by0 <= by < by0 + nby of the grid run when nbx > 0 (scheduling statistics of a
 * full-size frame from a sample of its tiles, tests/simt_stats.py) */
static int g_bx0 = 0, g_by0 = 0, g_nbx = 0, g_nby = 0;
extern "C" void simt_select_blocks(int bx0, int by0, int nbx, int nby) { g_bx0 = bx0; g_by0 = by0; g_nbx = nbx; g_nby = nby; }
extern "C" void simt_stats(unsigned long long* out16, int reset) {
#ifdef PT_STATS
    for (int i = 0; i < 16; i++) { out16[i] = pt_stats[i]; if (reset) pt_stats[i] = 0ull; }
#else
    for (int i = 0; i < 16; i++) out16[i] = 0ull;
    (void)reset;
#endif
}

/* pt_set_surface_ext of the emulated device (surface extensions; the kernel must be built with -DPT_EXT_BSDF=1) */
static std::vector<pt_surface_ext> g_surface_ext;
extern "C" void simt_set_surface_ext(const pt_surface_ext* table, int n) { g_surface_ext.assign(table, table + (n > 0 ? n : 0)); }

extern "C" int simt_dispatch(const pt_ubo* ubo, const pt_params* params, int accum_mode, int first, int n, float* image,
                             int persistent_ctas) {
    PtDevScene sc;
    PtDevParams dp;
    std::string err;
    if (pt_prepare_scene(ubo, &sc, &err) != 0) { fprintf(stderr, "simt: %s\n", err.c_str()); return -1; }
    if (pt_prepare_params(params, accum_mode, first, n, &dp, &err) != 0) { fprintf(stderr, "simt: %s\n", err.c_str()); return -2; }
    if (pt_prepare_surface_ext(g_surface_ext.data(), (int)g_surface_ext.size(), &sc) && !PT_EXT_BSDF) {
        fprintf(stderr, "simt: surface extensions set but the kernel was built without PT_EXT_BSDF\n");
        return -3;
    }
    gridDim = {(unsigned)((dp.width + 15) / 16), (unsigned)((dp.height + 7) / 8), 1u};
    blockDim = {PT_BLOCK_THREADS, 1u, 1u};
#if PT_PREGEN
    /* option "pregen" as pt_lib.cpp's launch() runs it: the generation kernel's records, then the render kernel, band by
     * band of CTA rows (persistent_ctas = rows per band here; 0 = the whole frame at once).  Every block runs its
     * generation part and then its render part: a block only reads the records of its own warps' tiles. */
    const unsigned gy_all = gridDim.y;
    const unsigned rows = (persistent_ctas > 0 && (unsigned)persistent_ctas < gy_all) ? (unsigned)persistent_ctas : gy_all;
    const size_t per_row = (size_t)gridDim.x * 4u * 32u * (size_t)dp.samplesPerFrame;
    std::vector<float4> gen(3 * rows * per_row); /* two planes of records + the radiance bundles of PT_RESOLVE */
    for (unsigned y0 = 0; y0 < gy_all; y0 += rows) {
    const unsigned ny = (gy_all - y0 < rows) ? gy_all - y0 : rows;
    gridDim.y = ny;
    dp.gen = gen.data();
    dp.genCount = (unsigned long long)ny * per_row;
    dp.rad = gen.data() + 2 * dp.genCount;
    dp.blockY0 = (int)y0;
    for (auto& g : gen) g = make_float4(NAN, NAN, NAN, NAN); /* a record nobody wrote would poison the image */
#else
    (void)persistent_ctas;
#endif
    const float* ubo_f = reinterpret_cast<const float*>(ubo);
    float4* img = reinterpret_cast<float4*>(image);
    const int nwarps = PT_BLOCK_THREADS / 32;
    /* One block at a time: `__shared__` arrays are function-local statics here, i.e. one copy for whoever runs.  For the
     * persistent grid of PT_SCHED=6 this means the first CTA drains the tile counter and the later ones find it
     * exhausted -- legal for a dynamic scheduler, and the counter's hand-over / self-reset is still exercised. */
    const unsigned concurrent = 1;
    for (unsigned b0 = 0; b0 < gridDim.x * gridDim.y; b0 += concurrent) {
        std::vector<std::thread> threads;
        std::vector<std::unique_ptr<SimtBlock>> blocks;
        std::vector<std::unique_ptr<SimtWarp>> warps;
        for (unsigned b = b0; b < b0 + concurrent && b < gridDim.x * gridDim.y; b++) {
            if (g_nbx > 0) {
                const int bx = (int)(b % gridDim.x), by = (int)(b / gridDim.x);
                if (bx < g_bx0 || bx >= g_bx0 + g_nbx || by < g_by0 || by >= g_by0 + g_nby) continue;
            }
            blocks.emplace_back(new SimtBlock(PT_BLOCK_THREADS));
            for (int w = 0; w < nwarps; w++) warps.emplace_back(new SimtWarp());
            SimtBlock* blk = blocks.back().get();
            for (int t = 0; t < PT_BLOCK_THREADS; t++) {
                SimtWarp* wp = warps[warps.size() - nwarps + t / 32].get();
                threads.emplace_back([=, &sc, &dp]() {
                    threadIdx = {(unsigned)t, 0u, 0u};
                    blockIdx = {b % gridDim.x, b / gridDim.x, 0u};
                    simt_warp = wp;
                    simt_block = blk;
                    simt_phase = 0;
#if PT_PREGEN
                    pt_gen_emu(dp, dp.gen);
                    __syncthreads();
#endif
                    pt_render_emu(sc, dp, ubo_f, img);
#if PT_RESOLVE
                    __syncthreads();
                    pt_resolve_emu(dp, ubo_f, img);
#endif
                });
            }
        }
        for (auto& th : threads) th.join();
    }
#if PT_PREGEN
    }
#endif
    return 0;
}
