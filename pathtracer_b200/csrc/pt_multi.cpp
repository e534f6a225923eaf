/* pt_multi.cpp -- sample-split rendering over the GPUs of one box from a single host thread.
 *
 * The reference drives one GPU.  The path shards embarrassingly -- the seed of a sample depends only on the pixel and
 * the sample index (GenerateSeed, shader.comp:948-958) -- so device g renders its own contiguous slice of the sample
 * index range for every pixel into a private fp32 SUM image (pt_dispatch_sum), and the path's one exchange step is a
 * single ncclReduce of those images to the first device over NVLink, followed by pt_finalize there.  The union of the
 * samples equals a one-GPU run; images agree up to fp32 summation order.
 *
 * Built on the public ABI only (one pt_ctx per device).  Kernel launches are asynchronous, so one thread keeps every
 * device busy by issuing the dispatches round-robin.  libnccl is dlopen'ed (like libnvrtc): a one-GPU box does not
 * need it, and a process that already carries torch's NCCL gets that copy.  This is the C++ twin of what bench.py does
 * with one process per GPU over torch.distributed.
 */
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "pt_abi.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
enum { kNcclFloat32 = 7, kNcclSum = 0 }; /* nccl.h: ncclFloat32 = 7, ncclSum = 0 */

struct Nccl {
    void* h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    std::string error;
};
Nccl g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
    const char* env = getenv("PT_NCCL_LIB");
    const char* names[] = {env ? env : "libnccl.so.2", "libnccl.so.2", "/usr/lib/x86_64-linux-gnu/libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !g_nccl.h; i++) g_nccl.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!g_nccl.h) { g_nccl.error = "cannot load libnccl.so.2 (set PT_NCCL_LIB)"; return; }
#define PT_SYM(field, name)                                                    \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name);                           \
    if (!g_nccl.field) { g_nccl.error = std::string("libnccl lacks ") + name; g_nccl.h = nullptr; return; }
    PT_SYM(CommInitAll, "ncclCommInitAll")
    PT_SYM(CommDestroy, "ncclCommDestroy")
    PT_SYM(Reduce, "ncclReduce")
    PT_SYM(GroupStart, "ncclGroupStart")
    PT_SYM(GroupEnd, "ncclGroupEnd")
    PT_SYM(GetErrorString, "ncclGetErrorString")
#undef PT_SYM
}

}  // namespace

struct pt_multi {
    std::vector<int> devices;
    std::vector<pt_ctx*> ctx;
    std::vector<ncclComm_t> comms; /* empty for one device */
    int width = 0, height = 0;
    double reduce_seconds = 0.0;
    std::string error;
};

namespace {

thread_local std::string g_create_error; /* pt_multi_last_error(NULL): why the last pt_multi_create failed */

int mfail(pt_multi* m, int code, const std::string& msg) {
    if (m) m->error = msg; else g_create_error = msg;
    return code;
}
int from_ctx(pt_multi* m, int g, int rc) {
    if (rc != PT_OK) m->error = "device " + std::to_string(m->devices[g]) + ": " + pt_last_error(m->ctx[g]);
    return rc;
}

}  // namespace

extern "C" {

const char* pt_multi_last_error(const pt_multi* m) { return m ? m->error.c_str() : g_create_error.c_str(); }

void pt_multi_destroy(pt_multi* m) {
    if (!m) return;
    for (size_t g = 0; g < m->comms.size(); g++)
        if (m->comms[g]) g_nccl.CommDestroy(m->comms[g]);
    for (pt_ctx* c : m->ctx) pt_destroy(c);
    delete m;
}

int pt_multi_create(const int* devices, int n_devices, int mode, pt_multi** out) {
    if (!out || !devices || n_devices < 1) return mfail(nullptr, PT_ERR_ARG, "pt_multi_create: bad argument");
    *out = nullptr;
    pt_multi* m = new pt_multi();
    for (int g = 0; g < n_devices; g++) {
        for (int k = 0; k < g; k++)
            if (devices[k] == devices[g]) { /* NCCL wants distinct devices */
                delete m;
                return mfail(nullptr, PT_ERR_ARG, "pt_multi_create: the same device listed twice");
            }
        m->devices.push_back(devices[g]);
    }
    for (int g = 0; g < n_devices; g++) {
        pt_ctx* c = nullptr;
        int rc = pt_create(devices[g], &c);
        if (rc == PT_OK) rc = pt_set_mode(c, mode);
        if (rc != PT_OK) {
            g_create_error = pt_last_error(c);
            if (c) pt_destroy(c);
            pt_multi_destroy(m);
            return rc;
        }
        m->ctx.push_back(c);
    }
    if (n_devices > 1) {
        std::call_once(g_nccl_once, load_nccl);
        if (!g_nccl.h) { pt_multi_destroy(m); return mfail(nullptr, PT_ERR_CUDA, g_nccl.error); }
        m->comms.assign(n_devices, nullptr);
        ncclResult_t r = g_nccl.CommInitAll(m->comms.data(), n_devices, m->devices.data());
        if (r != 0) {
            m->comms.clear();
            pt_multi_destroy(m);
            return mfail(nullptr, PT_ERR_CUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r));
        }
        /* NCCL sets its transports up lazily, inside the first collective (0.4 s on a 2-GPU box): do that here, on a
         * scratch buffer, not inside the first pt_multi_render */
        std::vector<void*> scratch(n_devices, nullptr);
        for (int g = 0; g < n_devices; g++) {
            cudaSetDevice(m->devices[g]);
            cudaMalloc(&scratch[g], 1024);
            cudaMemsetAsync(scratch[g], 0, 1024, (cudaStream_t)pt_stream_handle(m->ctx[g]));
        }
        r = g_nccl.GroupStart();
        for (int g = 0; g < n_devices && r == 0; g++) {
            cudaSetDevice(m->devices[g]);
            r = g_nccl.Reduce(scratch[g], scratch[g], 256, kNcclFloat32, kNcclSum, 0, m->comms[g], (cudaStream_t)pt_stream_handle(m->ctx[g]));
        }
        const ncclResult_t r2 = g_nccl.GroupEnd();
        for (int g = 0; g < n_devices; g++) {
            pt_sync(m->ctx[g]);
            cudaSetDevice(m->devices[g]);
            cudaFree(scratch[g]);
        }
        if (r != 0 || r2 != 0) {
            pt_multi_destroy(m);
            return mfail(nullptr, PT_ERR_CUDA, std::string("ncclReduce (warm-up): ") + g_nccl.GetErrorString(r != 0 ? r : r2));
        }
    }
    *out = m;
    return PT_OK;
}

int pt_multi_num_devices(const pt_multi* m) { return m ? (int)m->ctx.size() : 0; }
pt_ctx* pt_multi_ctx(pt_multi* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

int pt_multi_set_surface_ext(pt_multi* m, const pt_surface_ext* table, int n) {
    if (!m) return PT_ERR_ARG;
    for (size_t g = 0; g < m->ctx.size(); g++) {
        int rc = from_ctx(m, (int)g, pt_set_surface_ext(m->ctx[g], table, n));
        if (rc != PT_OK) return rc;
    }
    return PT_OK;
}

/* per-context settings (pt_set_jit, pt_set_pipeline, pt_set_bvh) go through pt_multi_ctx before this call */
int pt_multi_set_scene(pt_multi* m, const pt_ubo* ubo, const char* const* sdf_glsl, int n_sdf) {
    if (!m || !ubo) return mfail(m, PT_ERR_ARG, "pt_multi_set_scene: null argument");
    for (size_t g = 0; g < m->ctx.size(); g++) { /* the kernel is compiled once: cubins are cached process-wide */
        int rc = from_ctx(m, (int)g, pt_set_scene(m->ctx[g], ubo, sdf_glsl, n_sdf));
        if (rc != PT_OK) return rc;
    }
    return PT_OK;
}

int pt_multi_resize(pt_multi* m, int width, int height) {
    if (!m) return PT_ERR_ARG;
    for (size_t g = 0; g < m->ctx.size(); g++) {
        int rc = from_ctx(m, (int)g, pt_resize(m->ctx[g], width, height));
        if (rc != PT_OK) return rc;
    }
    m->width = width;
    m->height = height;
    return PT_OK;
}

/* Renders sample indices first_sample .. first_sample + total_samples - 1 of every pixel: device g takes the g-th
 * contiguous slice, samples_per_dispatch at a time; then one ncclReduce to the first device and pt_finalize there.
 * Blocks until the image on the first device is final.  seconds (optional) = host wall time of dispatches + reduce +
 * finalize; pt_multi_reduce_seconds() = the reduce alone. */
int pt_multi_render(pt_multi* m, const pt_params* base, int first_sample, int total_samples, int samples_per_dispatch,
                    double* seconds) {
    if (!m || !base) return mfail(m, PT_ERR_ARG, "pt_multi_render: null argument");
    const int G = (int)m->ctx.size();
    if (total_samples < G || samples_per_dispatch <= 0)
        return mfail(m, PT_ERR_ARG, "pt_multi_render: need total_samples >= number of devices and samples_per_dispatch > 0");
    if (m->width <= 0) return mfail(m, PT_ERR_ARG, "pt_multi_render: call pt_multi_resize first");
    for (int g = 0; g < G; g++) {
        int rc = from_ctx(m, g, pt_clear(m->ctx[g]));
        if (rc != PT_OK) return rc;
    }
    for (int g = 0; g < G; g++) pt_sync(m->ctx[g]);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<int> next(G), end(G);
    for (int g = 0; g < G; g++) {
        next[g] = first_sample + (int)((long long)total_samples * g / G);
        end[g] = first_sample + (int)((long long)total_samples * (g + 1) / G);
    }
    for (bool more = true; more;) { /* round-robin: launches are asynchronous, every device stays busy */
        more = false;
        for (int g = 0; g < G; g++) {
            if (next[g] >= end[g]) continue;
            const int n = (end[g] - next[g] < samples_per_dispatch) ? end[g] - next[g] : samples_per_dispatch;
            int rc = from_ctx(m, g, pt_dispatch_sum(m->ctx[g], base, next[g], n));
            if (rc != PT_OK) return rc;
            next[g] += n;
            more = true;
        }
    }
    m->reduce_seconds = 0.0;
    if (G > 1) {
        for (int g = 0; g < G; g++) pt_sync(m->ctx[g]);
        const auto r0 = std::chrono::steady_clock::now();
        const size_t count = (size_t)m->width * (size_t)m->height * 4;
        ncclResult_t r = g_nccl.GroupStart();
        for (int g = 0; g < G && r == 0; g++) {
            cudaSetDevice(m->devices[g]);
            void* img = pt_image_ptr(m->ctx[g]);
            r = g_nccl.Reduce(img, img, count, kNcclFloat32, kNcclSum, 0, m->comms[g], (cudaStream_t)pt_stream_handle(m->ctx[g]));
        }
        const ncclResult_t r2 = g_nccl.GroupEnd();
        if (r == 0) r = r2;
        if (r != 0) return mfail(m, PT_ERR_CUDA, std::string("ncclReduce: ") + g_nccl.GetErrorString(r));
        for (int g = 0; g < G; g++) pt_sync(m->ctx[g]);
        m->reduce_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - r0).count();
    }
    int rc = from_ctx(m, 0, pt_finalize(m->ctx[0], base, total_samples));
    if (rc == PT_OK) rc = from_ctx(m, 0, pt_sync(m->ctx[0]));
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

double pt_multi_reduce_seconds(const pt_multi* m) { return m ? m->reduce_seconds : 0.0; }

int pt_multi_read_xyz(pt_multi* m, float* rgba, size_t n_floats) {
    if (!m) return PT_ERR_ARG;
    return from_ctx(m, 0, pt_read_xyz(m->ctx[0], rgba, n_floats));
}

} /* extern "C" */
