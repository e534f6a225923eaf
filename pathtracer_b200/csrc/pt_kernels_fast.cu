/* pt_kernels_fast.cu -- the statically compiled FAST instance of the megakernel (scenes without SDFs).
 * Compile flags (csrc/Makefile): -fmad=true -prec-div=false -prec-sqrt=false -ftz=true, MUFU transcendentals. */
#define PT_FAST 1
#define PT_MIN_BLOCKS 6 /* 85 registers: +12% on scene1 over 4 CTAs/SM (profiles/r01_sched_ab.md) */
#define PT_KERNEL_NS ptk_fast
#ifndef PT_SCHED
#define PT_SCHED 0 /* v1 driver: faster than v2 on scenes without SDFs (profiles/r01_v2_sched) */
#endif
#include "pt_kernel.cuh"

PT_DEFINE_RENDER_KERNEL(pt_render_fast)

extern "C" void pt_launch_fast(const PtDevScene* sc, const PtDevParams* pr, const float* ubo, void* image,
                               void* stream) {
    dim3 grid((pr->width + 15) / 16, (pr->height + 7) / 8);
    pt_render_fast<<<grid, PT_BLOCK_THREADS, 0, (cudaStream_t)stream>>>(*sc, *pr, ubo, (float4*)image);
}
extern "C" const void* pt_static_kernel_fast(void) { return (const void*)pt_render_fast; }
