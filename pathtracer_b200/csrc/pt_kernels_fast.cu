/* pt_kernels_fast.cu -- the statically compiled FAST instance of the megakernel (scenes without SDFs).
 * Compile flags (csrc/Makefile): -fmad=true -prec-div=false -prec-sqrt=false -ftz=true, MUFU transcendentals. */
#define PT_FAST 1
#define PT_MIN_BLOCKS 6 /* 85 registers: +12% on scene1 over 4 CTAs/SM (profiles/r01_sched_ab.md) */
#define PT_KERNEL_NS ptk_fast
#ifndef PT_SCHED
#define PT_SCHED 0 /* v1 driver (the scene-specialised JIT kernels are what bench.py and the CLI use) */
#endif
#include "pt_kernel.cuh"

PT_DEFINE_RENDER_KERNEL(pt_render_fast)

extern "C" void pt_launch_fast(const PtDevScene* sc, const PtDevParams* pr, const float* ubo, void* image,
                               void* stream) {
    dim3 grid((pr->width + 15) / 16, (pr->height + 7) / 8);
    pt_render_fast<<<grid, PT_BLOCK_THREADS, 0, (cudaStream_t)stream>>>(*sc, *pr, ubo, (float4*)image);
}
extern "C" const void* pt_static_kernel_fast(void) { return (const void*)pt_render_fast; }

/* FP32 peak: 16 independent FFMA chains per thread, nothing else in the loop (pt_fp32_peak) */
__global__ void __launch_bounds__(256) pt_fma_peak_kernel(float* out, int iters) {
    float a[16];
    const float x = 1.0f + 1e-7f * (float)threadIdx.x, y = 1e-9f * (float)blockIdx.x;
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = (float)i + x;
#pragma unroll 8 /* 128 FFMA per trip: the loop counter and branch cost 2 % of the issue slots, not 16 % */
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, y);
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
extern "C" void pt_launch_fma_peak(float* out, int blocks, int threads, int iters, void* stream) {
    pt_fma_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(out, iters);
}
