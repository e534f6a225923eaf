/* pt_jit.cpp -- run-time compilation of the megakernel with the scene's SDF snippets (NVRTC -> sm_100a cubin).
 *
 * Stands in for the reference's GLSLToSPIRV + vkCreateComputePipelines on every scene load
 * (host:1811-1868, 2056-2092, RecompileComputeShaders host:3836-3841).  The translation unit handed to NVRTC is
 *     #define PT_HAS_SDF 1 [+ baked primitive counts]
 *     #include "pt_kernel.cuh"        (embedded in the library, see tools/embed_headers.py)
 *     <output of pt_sdf_generate()>   (snippets + SDF()/SDFMATERIAL() dispatchers)
 *     PT_DEFINE_RENDER_KERNEL(pt_render_jit)
 * libnvrtc is dlopen'ed so that libpt_cuda.so loads (and the static kernels run) on a box without it; compiling
 * needs no GPU, which is how the CPU-side tests exercise this path.
 */
#include <dlfcn.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "pt_internal.h"

namespace {

typedef struct _nvrtcProgram* nvrtcProgram;
typedef int nvrtcResult;

struct Nvrtc {
    void* h = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*DestroyProgram)(nvrtcProgram*);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
    const char* (*GetErrorString)(nvrtcResult);
    nvrtcResult (*Version)(int*, int*);
    std::string where;
};

Nvrtc g_nvrtc;
std::once_flag g_once;
std::string g_load_error;

/* Process-wide cache of compiled kernels, keyed by the complete translation unit (which spells out the mode, the baked
 * counts, every tuning knob and the generated SDF code): a cubin is independent of the device it will be loaded on, so
 * the contexts of a multi-GPU render (pt_multi.cpp) compile each scene once instead of once per GPU. */
std::mutex g_cache_mu;
std::map<std::string, std::pair<std::vector<char>, std::string>> g_cubin_cache;

void load_nvrtc() {
    const char* env = getenv("PT_NVRTC_LIB");
    const char* names[] = {env ? env : "libnvrtc.so.12", "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so", nullptr};
    for (int i = 0; names[i] && !g_nvrtc.h; i++) {
        g_nvrtc.h = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
        if (g_nvrtc.h) g_nvrtc.where = names[i];
    }
    if (!g_nvrtc.h) { g_load_error = "cannot load libnvrtc.so.12 (set PT_NVRTC_LIB)"; return; }
#define PT_SYM(field, name)                                                       \
    *(void**)(&g_nvrtc.field) = dlsym(g_nvrtc.h, name);                            \
    if (!g_nvrtc.field) { g_load_error = std::string("libnvrtc lacks ") + name; g_nvrtc.h = nullptr; return; }
    PT_SYM(CreateProgram, "nvrtcCreateProgram")
    PT_SYM(DestroyProgram, "nvrtcDestroyProgram")
    PT_SYM(CompileProgram, "nvrtcCompileProgram")
    PT_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    PT_SYM(GetCUBIN, "nvrtcGetCUBIN")
    PT_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    PT_SYM(GetProgramLog, "nvrtcGetProgramLog")
    PT_SYM(GetErrorString, "nvrtcGetErrorString")
    PT_SYM(Version, "nvrtcVersion")
#undef PT_SYM
}

}  // namespace

int pt_jit_compile(const std::string& sdf_unit, const PtJitOptions& opt, std::vector<char>* cubin, std::string* log) {
    std::call_once(g_once, load_nvrtc);
    if (!g_nvrtc.h) { *log = g_load_error; return PT_ERR_COMPILE; }

    std::string src;
    src += opt.mode == PT_MODE_FAST ? "#define PT_FAST 1\n#define PT_KERNEL_NS ptk_jit_fast\n"
                                    : "#define PT_KERNEL_NS ptk_jit_strict\n";
    if (!sdf_unit.empty()) src += "#define PT_HAS_SDF 1\n";
    if (opt.bvh) src += "#define PT_BVH 1\n";
    if (opt.bake_counts) {
        const char* names[6] = {"PT_N_SPHERES_CONST", "PT_N_PLANES_CONST", "PT_N_BOXES_CONST", "PT_N_LENSES_CONST",
                                "PT_N_CYCLIDES_CONST", "PT_N_SDF_CONST"};
        for (int i = 0; i < 6; i++) src += std::string("#define ") + names[i] + " " + std::to_string(opt.counts[i]) + "\n";
    }
    /* Driver selection and tuning knobs (env overrides are for A/B measurements: bench.py, profiles/).
     * Measured on B200 (profiles/r01_sched_ab.md): scenes without SDFs are fastest with the v1 driver (nested
     * loops), scenes with SDFs with the v2 in-warp scheduler and rolled primitive loops (the kernel is instruction-
     * cache bound); the fast build wants 6 CTAs/SM, the strict one 4. */
    struct Knob { const char* name; int dflt; };
    /* BVH scenes (profiles/r01_bvh): with boxes / lenses / cyclides among the primitives the in-warp scheduler wins
     * (2.09 vs 1.61 Gsamples/s on the 74-primitive mix), with spheres only v1 does (3.76 vs 3.58 on 169 spheres).
     * Scan with many primitives: unrolled loops over baked counts spill (169 spheres: 0.10 vs 0.88 Gsamples/s rolled). */
    const bool heavy = (opt.counts[2] + opt.counts[3] + opt.counts[4]) > 0;
    const int n_scanned = opt.counts[0] + opt.counts[1] + opt.counts[2] + opt.counts[3] + opt.counts[4];
    /* v2s, the in-warp scheduler with sample stealing (profiles/r01_steal): +16..40 % over v2 on the SDF workloads, +9 % on
     * the 74-primitive mix; scenes the reference's scan handles without SDFs stay on v1 (10.85 vs 9.85 Gsamples/s, cfg2). */
    const int sched = (!sdf_unit.empty() || (opt.bvh && heavy)) ? 5 : 0;
    const char* sched_env = getenv("PT_SCHED");
    int sched_eff = (sched_env && sched_env[0]) ? atoi(sched_env) : sched;
    if (sched_eff == 6 && opt.mode != PT_MODE_FAST) sched_eff = 5; /* v2sp sums in schedule order: strict builds keep v2s' table */
    const int no_unroll = (!sdf_unit.empty() || (!opt.bvh && n_scanned > 16)) ? 1 : 0;
    /* v2s: strict mode needs the per-sample table (sums in sample order); fast mode pools the whole dispatch (0) */
    const char* steal_env = getenv("PT_STEAL_S");
    const char* park_env = getenv("PT_MPARK");
    const int park = (park_env && park_env[0]) ? atoi(park_env) : 0;
    /* the strict build's table shares the 48 KB of static shared memory with the parking stack */
    const int steal_dflt = opt.mode == PT_MODE_FAST ? 0 : ((park && !sdf_unit.empty()) ? 8 : 16);
    int steal_s = (steal_env && steal_env[0]) ? atoi(steal_env) : steal_dflt;
    if (opt.mode != PT_MODE_FAST && steal_s == 0) steal_s = steal_dflt;
    if (opt.mode != PT_MODE_FAST && park && !sdf_unit.empty() && steal_s > 8) steal_s = 8;
    /* shared memory per CTA: v2d 45 KB, v2s 16 KB + 1.5 KB x PT_STEAL_S -> 5 CTAs/SM fit */
    const int min_blocks_fast = (sched_eff == 4 || ((sched_eff == 5 || sched_eff == 7) && steal_s > 8)) ? 5 : 6;
    const Knob knobs[] = {{"PT_SCHED", sched}, {"PT_SDF_REPS", 16}, {"PT_FEED_T", 8}, {"PT_STEAL_S", -1},
                          {"PT_MIN_BLOCKS", opt.mode == PT_MODE_FAST ? min_blocks_fast : 4},
                          {"PT_NO_UNROLL", no_unroll}, {"PT_STATS", 0},
                          {"PT_WF_REFILL", 8}, {"PT_POOL_MIN", 24}, {"PT_COOP_NORMALS", 0}, {"PT_BVH_WHILE_WHILE", 0}, {"PT_REGEN_T", 16}, {"PT_SDF_MIN", 0}, {"PT_SDF_EXIT", 0}, {"PT_SWAP_MIN", 8}, {"PT_TILE_SLOTS", 4}, {"PT_MPARK", 0}, {"PT_MPARK_CAP", 20}, {"PT_MPARK_MIN", 12}};
    for (const Knob& k : knobs) {
        const char* v = getenv(k.name);
        const int val = (v && v[0]) ? atoi(v) : k.dflt;
        if (std::string(k.name) == "PT_STATS" && val == 0) continue;
        if (std::string(k.name) == "PT_SCHED") { src += "#define PT_SCHED " + std::to_string(sched_eff) + "\n"; continue; }
        if (std::string(k.name) == "PT_STEAL_S") { src += "#define PT_STEAL_S " + std::to_string(steal_s) + "\n"; continue; }
        src += std::string("#define ") + k.name + " " + std::to_string(val) + "\n";
    }
    src += opt.wavefront ? "#include \"pt_wavefront.cuh\"\n" : "#include \"pt_kernel.cuh\"\n";
    src += sdf_unit;
    src += "\nPT_DEFINE_RENDER_KERNEL(pt_render_jit)\n";
    if (!sdf_unit.empty()) src += "PT_DEFINE_SDF_EVAL_KERNEL(pt_sdf_eval_jit)\n";
    if (opt.wavefront) src += "PT_DEFINE_WAVEFRONT_KERNELS\n";

    if (!getenv("PT_JIT_DUMP")) {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        auto it = g_cubin_cache.find(src);
        if (it != g_cubin_cache.end()) {
            *cubin = it->second.first;
            *log = it->second.second;
            return PT_OK;
        }
    }
    std::vector<const char*> hdr_names, hdr_texts;
    for (int i = 0; i < pt_embedded_header_count; i++) {
        hdr_names.push_back(pt_embedded_headers[i].name);
        hdr_texts.push_back(pt_embedded_headers[i].text);
    }
    nvrtcProgram prog = nullptr;
    nvrtcResult r = g_nvrtc.CreateProgram(&prog, src.c_str(), "pt_render_jit.cu", (int)hdr_names.size(), hdr_texts.data(),
                                          hdr_names.data());
    if (r != 0) { *log = std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(r); return PT_ERR_COMPILE; }

    std::vector<const char*> o = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device", "-lineinfo", "--ptxas-options=-v"};
    if (opt.mode == PT_MODE_FAST) {
        o.insert(o.end(), {"--fmad=true", "--prec-div=false", "--prec-sqrt=false", "--ftz=true"});
    } else {
        o.insert(o.end(), {"--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false"});
    }
    r = g_nvrtc.CompileProgram(prog, (int)o.size(), o.data());
    size_t ls = 0;
    g_nvrtc.GetProgramLogSize(prog, &ls);
    std::string plog;
    if (ls > 1) {
        plog.resize(ls);
        g_nvrtc.GetProgramLog(prog, &plog[0]);
        while (!plog.empty() && plog.back() == '\0') plog.pop_back();
    }
    if (r != 0) {
        *log = std::string("NVRTC: ") + g_nvrtc.GetErrorString(r) + "\n" + plog;
        g_nvrtc.DestroyProgram(&prog);
        return PT_ERR_COMPILE;
    }
    size_t cs = 0;
    r = g_nvrtc.GetCUBINSize(prog, &cs);
    if (r != 0 || cs == 0) {
        *log = "NVRTC produced no cubin";
        g_nvrtc.DestroyProgram(&prog);
        return PT_ERR_COMPILE;
    }
    cubin->resize(cs);
    g_nvrtc.GetCUBIN(prog, cubin->data());
    if (const char* dir = getenv("PT_JIT_DUMP")) { /* keep the generated source and the cubin (profiling: nvdisasm -g) */
        static int serial = 0;
        const std::string base = std::string(dir) + "/pt_render_jit_" + std::to_string(serial++);
        if (FILE* f = fopen((base + ".cu").c_str(), "w")) { fwrite(src.data(), 1, src.size(), f); fclose(f); }
        if (FILE* f = fopen((base + ".cubin").c_str(), "wb")) { fwrite(cubin->data(), 1, cubin->size(), f); fclose(f); }
    }
    g_nvrtc.DestroyProgram(&prog);
    *log = plog;
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        if (g_cubin_cache.size() >= 64) g_cubin_cache.clear(); /* bound the memory of a long-lived host process */
        g_cubin_cache[src] = std::make_pair(*cubin, plog);
    }
    return PT_OK;
}
