/* pt_kernels_strict.cu -- the statically compiled STRICT instance of the megakernel (scenes without SDFs).
 * Compile flags (csrc/Makefile): -fmad=false -prec-div=true -prec-sqrt=true -ftz=false.  Bit-exact vs oracle/. */
#define PT_KERNEL_NS ptk_strict
#ifndef PT_SCHED
#define PT_SCHED 0
#endif
#include "pt_kernel.cuh"

PT_DEFINE_RENDER_KERNEL(pt_render_strict)

extern "C" void pt_launch_strict(const PtDevScene* sc, const PtDevParams* pr, const float* ubo, void* image,
                                 void* stream) {
    dim3 grid((pr->width + 15) / 16, (pr->height + 7) / 8);
    pt_render_strict<<<grid, PT_BLOCK_THREADS, 0, (cudaStream_t)stream>>>(*sc, *pr, ubo, (float4*)image);
}

/* pt_math.h on the device, for the CPU<->GPU bit-equality tests (pt_math_eval) */
__global__ void pt_math_eval_kernel(int fn, const float* x, const float* y, float* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r = 0.0f;
    switch (fn) {
        case 0: r = pt_sin(x[i]); break;
        case 1: r = pt_cos(x[i]); break;
        case 2: r = pt_acos(x[i]); break;
        case 3: r = pt_exp2(x[i]); break;
        case 4: r = pt_log2(x[i]); break;
        case 5: r = pt_exp(x[i]); break;
        case 6: r = pt_log(x[i]); break;
        case 7: r = pt_pow(x[i], y[i]); break;
        case 8: { unsigned s = __float_as_uint(x[i]); ptk_strict::PCG32(s); r = __uint_as_float(s); break; }
        case 9: { unsigned s = __float_as_uint(x[i]); r = ptk_strict::RandomFloatPCG32(s); break; }
    }
    out[i] = r;
}
extern "C" void pt_launch_math_eval(int fn, const float* x, const float* y, float* out, size_t n, void* stream) {
    if (n == 0) return;
    pt_math_eval_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fn, x, y, out, n);
}

/* pt_finalize: sum over all samples -> the reference's units (sum / total * apertureSize^2 * ISO, w = 1) */
__global__ void pt_finalize_kernel(float4* image, int n, float invTotal, float exposure) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = image[i];
    v.x = v.x * invTotal * exposure; v.y = v.y * invTotal * exposure; v.z = v.z * invTotal * exposure; v.w = 1.0f;
    image[i] = v;
}
extern "C" void pt_launch_finalize(void* image, int n_texels, float invTotal, float exposure, void* stream) {
    pt_finalize_kernel<<<(n_texels + 255) / 256, 256, 0, (cudaStream_t)stream>>>((float4*)image, n_texels, invTotal, exposure);
}
extern "C" const void* pt_static_kernel_strict(void) { return (const void*)pt_render_strict; }
