/* cuda_shim.h -- TEST INFRASTRUCTURE: just enough of the CUDA execution model on the host to run the product's kernel
 * source (pathtracer_b200/csrc/pt_kernel.cuh, unmodified) without a GPU.
 *
 * One host thread per CUDA thread, one block at a time.  `__shared__` arrays become function-local statics (shared by the
 * 128 threads of the running block), `__syncthreads()` a barrier over the block, and the warp collectives
 * (`__ballot_sync`, `__shfl_sync`, `__syncwarp`) a barrier over the 32 threads of the warp plus a double-buffered
 * exchange area -- so the warp-synchronous DRIVERS (phase votes, sample claims, parking, tile bookkeeping) execute the
 * way they do on the device: every lane runs its own divergent code between two collectives, and the collectives see
 * all 32 lanes.  Only full-mask collectives are supported (the drivers use nothing else); anything else aborts.
 * Strict-mode arithmetic on the host is the oracle's (pt_math.h, -ffp-contract=off), so an emulated strict render must
 * equal the oracle bit for bit -- tests/test_simt_emulation.py -- which checks driver logic before a GPU minute is
 * spent.  This is a checker: nothing under pathtracer_b200/ includes it, and it is no rendering path of the product. */
#ifndef PT_CUDA_SHIM_H
#define PT_CUDA_SHIM_H

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __constant__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __grid_constant__
#define __launch_bounds__(...)

struct SimtDim3 { unsigned x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(8) uint2 { unsigned x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r = {x, y}; return r; }

struct SimtWarp {
    std::barrier<> bar{32};
    uint32_t xchg[2][32];
};
struct SimtBlock {
    std::barrier<> bar;
    explicit SimtBlock(int n) : bar(n) {}
};
extern thread_local SimtDim3 threadIdx, blockIdx;
extern SimtDim3 gridDim, blockDim;
extern thread_local SimtWarp* simt_warp;
extern thread_local SimtBlock* simt_block;
extern thread_local unsigned simt_phase; /* parity of the warp's next collective (same on all its lanes) */

static inline void simt_require_full(unsigned mask) {
    if (mask != 0xffffffffu) { fprintf(stderr, "simt: partial-mask collective (%08x) is not supported\n", mask); abort(); }
}
static inline void __syncthreads() { simt_block->bar.arrive_and_wait(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt_require_full(mask); simt_warp->bar.arrive_and_wait(); }
static inline uint32_t simt_exchange(uint32_t mine, int src) { /* src < 0: gather a ballot */
    SimtWarp& w = *simt_warp;
    const unsigned ph = simt_phase & 1u;
    simt_phase++;
    w.xchg[ph][threadIdx.x & 31u] = mine;
    w.bar.arrive_and_wait();
    if (src >= 0) return w.xchg[ph][src & 31];
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= (w.xchg[ph][i] ? 1u : 0u) << i;
    return r;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { /* redux.sync.or.b32 */
    simt_require_full(mask);
    SimtWarp& w = *simt_warp;
    const unsigned ph = simt_phase & 1u;
    simt_phase++;
    w.xchg[ph][threadIdx.x & 31u] = v;
    w.bar.arrive_and_wait();
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= w.xchg[ph][i];
    return r;
}
static inline unsigned __ballot_sync(unsigned mask, bool pred) { simt_require_full(mask); return simt_exchange(pred ? 1u : 0u, -1); }
static inline float __shfl_sync(unsigned mask, float v, int src) {
    simt_require_full(mask);
    uint32_t b; memcpy(&b, &v, 4);
    b = simt_exchange(b, src);
    float r; memcpy(&r, &b, 4);
    return r;
}
static inline int __shfl_sync(unsigned mask, int v, int src) { simt_require_full(mask); return (int)simt_exchange((uint32_t)v, src); }
static inline unsigned __shfl_sync(unsigned mask, unsigned v, int src) { simt_require_full(mask); return simt_exchange(v, src); }

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __fns(unsigned mask, unsigned base, int offset) { /* offset-th set bit at or above base (offset >= 1) */
    for (unsigned i = base; i < 32; i++) if ((mask >> i) & 1u) { if (--offset == 0) return i; }
    return 0xffffffffu;
}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float2int_rz(float f) { /* cvt.rzi.s32.f32: saturating, NaN -> 0 */
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
}
static inline float __uint2float_rn(unsigned u) { return (float)u; }
/* glibc's <cmath> already declares __sinf / __cosf / __expf / __powf (its internal float entry points); the fast-math
 * drivers are not emulated anyway (PT_FAST is never defined here) */
static inline float __fdividef(float a, float b) { return a / b; }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __nanosleep(unsigned) {}

template <class T> static inline T simt_atomic_add(T* p, T v) { return std::atomic_ref<T>(*p).fetch_add(v); }
static inline float atomicAdd(float* p, float v) { return simt_atomic_add(p, v); }
static inline int atomicAdd(int* p, int v) { return simt_atomic_add(p, v); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return simt_atomic_add(p, v); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return simt_atomic_add(p, v); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_or(v); }
static inline unsigned atomicAnd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_and(v); }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }

#endif
