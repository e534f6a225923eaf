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

/* the drivers the "auto" option resolves to (DESIGN.md section 7 has the A/B numbers) */
#ifndef PT_DEFAULT_SCHED_SDF
#define PT_DEFAULT_SCHED_SDF 5
#endif
#ifndef PT_DEFAULT_HEAVY_MIN
#define PT_DEFAULT_HEAVY_MIN 0
#endif
#ifndef PT_DEFAULT_PATHCOLOR_UNROLL_WITH_PREGEN
#define PT_DEFAULT_PATHCOLOR_UNROLL_WITH_PREGEN 1 /* cfg5 4.53 -> 4.58 Gsamples/s, cfg3 unchanged (profiles/r02_alu) */
#endif
#ifndef PT_DEFAULT_RESOLVE
#define PT_DEFAULT_RESOLVE 0 /* measured speed-neutral (cfg5 4.26 -> 4.30, cfg1 9.10 -> 8.94: profiles/r02_pregen); an option for schedule-independent sums */
#endif
#ifndef PT_DEFAULT_SCHED_ANALYTIC
#define PT_DEFAULT_SCHED_ANALYTIC 7 /* v3s: 10.73 vs v1's 9.96 Gsamples/s on cfg2, 8.10 vs 7.62 on cfg1 (profiles/r02_gpu1) */
#endif

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

int pt_knob_set(PtKnobs* k, const char* key, long long value) {
    if (!k || !key) return -1;
    const std::string s(key);
    const int v = (int)value;
    if (s == "sched") { if (v != -1 && v != 0 && v != 5 && v != 7 && v != 8) return -1; k->sched = v; }
    else if (s == "sdf_reps") { if (v < 1 || v > 512) return -1; k->sdf_reps = v; }
    else if (s == "feed_t") { if (v < 0 || v > 33) return -1; k->feed_t = v; }
    else if (s == "regen_t") { if (v < 1 || v > 32) return -1; k->regen_t = v; }
    else if (s == "steal_s") { if (v < -1 || v > 20) return -1; k->steal_s = v; }
    else if (s == "min_blocks") { if (v < -1 || v == 0 || v > 16) return -1; k->min_blocks = v; }
    else if (s == "no_unroll") { if (v < -1 || v > 1) return -1; k->no_unroll = v; }
    else if (s == "pool_cap") { if (v < 1 || v > 32) return -1; k->pool_cap = v; }
    else if (s == "pool_min") { if (v < 1 || v > 64) return -1; k->pool_min = v; }
    else if (s == "stats") { if (v < 0 || v > 1) return -1; k->stats = v; }
    else if (s == "wf_refill") { if (v < 1 || v > 512) return -1; k->wf_refill = v; }
    else if (s == "bvh_while_while") { if (v < 0 || v > 1) return -1; k->bvh_while_while = v; }
    else if (s == "sin_poly_every") { if (v < 0 || v > 64) return -1; k->sin_poly_every = v; }
    else if (s == "heavy_min") { if (v < -1 || v > 32) return -1; k->heavy_min = v; }
    else if (s == "pregen") { if (v < -1 || v > 1) return -1; k->pregen = v; }
    else if (s == "resolve") { if (v < -1 || v > 1) return -1; k->resolve = v; }
    else if (s == "pathcolor_unroll") { if (v < -1 || v > 1) return -1; k->pathcolor_unroll = v; }
    else return -1;
    return 0;
}
int pt_knob_get(const PtKnobs* k, const char* key, long long* value) {
    if (!k || !key || !value) return -1;
    const std::string s(key);
    if (s == "sched") *value = k->sched; else if (s == "sdf_reps") *value = k->sdf_reps;
    else if (s == "feed_t") *value = k->feed_t; else if (s == "regen_t") *value = k->regen_t;
    else if (s == "steal_s") *value = k->steal_s; else if (s == "min_blocks") *value = k->min_blocks;
    else if (s == "no_unroll") *value = k->no_unroll; else if (s == "pool_cap") *value = k->pool_cap;
    else if (s == "pool_min") *value = k->pool_min; else if (s == "stats") *value = k->stats;
    else if (s == "wf_refill") *value = k->wf_refill; else if (s == "bvh_while_while") *value = k->bvh_while_while;
    else if (s == "sin_poly_every") *value = k->sin_poly_every; else if (s == "heavy_min") *value = k->heavy_min;
    else if (s == "pregen") *value = k->pregen;
    else if (s == "resolve") *value = k->resolve;
    else if (s == "pathcolor_unroll") *value = k->pathcolor_unroll;
    else return -1;
    return 0;
}
int pt_knobs_parse(PtKnobs* k, const char* text, std::string* err) {
    if (!text) return 0;
    std::string s(text);
    size_t i = 0;
    while (i < s.size()) {
        size_t j = s.find(',', i);
        if (j == std::string::npos) j = s.size();
        const std::string item = s.substr(i, j - i);
        i = j + 1;
        if (item.empty()) continue;
        const size_t eq = item.find('=');
        if (eq == std::string::npos || pt_knob_set(k, item.substr(0, eq).c_str(), atoll(item.c_str() + eq + 1)) != 0) {
            if (err) *err = "bad option '" + item + "'";
            return -1;
        }
    }
    return 0;
}

/* Driver selection and the defaults of the "auto" options, as measured on B200:
 *   scenes the reference's scan handles without SDFs run the flat loop with the sample pool (v3s; round 1: v1);
 *   scenes with SDFs, and BVH scenes with boxes / lenses / cyclides, run the phase machine (v2m / v2s);
 *   primitive loops stay rolled when the kernel would otherwise outgrow the instruction cache (SDF scenes) or spill
 *   (scan over more than 16 primitives: 169 spheres unrolled 0.10 vs 0.88 Gsamples/s rolled);
 *   strict builds sum a round's samples in index order from a shared-memory table, fast builds pool the whole dispatch. */
std::string pt_jit_source(const std::string& sdf_unit, const PtJitOptions& opt) {
    const PtKnobs& k = opt.knobs;
    std::string src;
    src += opt.mode == PT_MODE_FAST ? "#define PT_FAST 1\n#define PT_KERNEL_NS ptk_jit_fast\n"
                                    : "#define PT_KERNEL_NS ptk_jit_strict\n";
    if (!sdf_unit.empty()) src += "#define PT_HAS_SDF 1\n";
    if (opt.bvh) src += "#define PT_BVH 1\n";
    if (opt.surface_ext) src += "#define PT_EXT_BSDF 1\n";
    if (opt.bake_counts) {
        const char* names[6] = {"PT_N_SPHERES_CONST", "PT_N_PLANES_CONST", "PT_N_BOXES_CONST", "PT_N_LENSES_CONST",
                                "PT_N_CYCLIDES_CONST", "PT_N_SDF_CONST"};
        for (int i = 0; i < 6; i++) src += std::string("#define ") + names[i] + " " + std::to_string(opt.counts[i]) + "\n";
    }
    const bool has_sdf = !sdf_unit.empty();
    const bool heavy = (opt.counts[2] + opt.counts[3] + opt.counts[4]) > 0;
    const int n_scanned = opt.counts[0] + opt.counts[1] + opt.counts[2] + opt.counts[3] + opt.counts[4];
    int sched = k.sched;
    if (sched < 0) sched = has_sdf ? PT_DEFAULT_SCHED_SDF : ((opt.bvh && heavy) ? 5 : PT_DEFAULT_SCHED_ANALYTIC);
    if (sched == 8 && !has_sdf) sched = 5; /* nothing marches */
    const int sdf_words = opt.counts[5] > 0 ? (opt.counts[5] + 31) / 32 : 1;
    if (sched == 8 && sdf_words > 1) sched = 5; /* the pool parks one mask word per path */
    const int no_unroll = k.no_unroll >= 0 ? k.no_unroll : ((has_sdf || (!opt.bvh && n_scanned > 16)) ? 1 : 0);
    /* camera rays from records (pt_gen_body): the pooled megakernel drivers only */
    /* auto: scenes with a cyclide -- their kernel is 39-41 KB of SASS with the camera inline and 35-36 KB without, i.e.
     * the generation kernel buys the instruction cache (cfg5 3.93 -> 4.26, cfg1 8.17 -> 8.86, cfg3 3.49 -> 3.61 Gsamples/s);
     * where the kernel fits anyway the records' 64 B per sample of memory traffic eat the gain (cfg2 11.49 -> 11.41, cfg4b
     * 1.058 -> 1.050): profiles/r02_pregen */
    const bool pregen = !opt.wavefront && (sched == 5 || sched == 7) && (k.pregen < 0 ? opt.counts[4] > 0 : k.pregen != 0);
    const bool resolve = pregen && (k.resolve < 0 ? PT_DEFAULT_RESOLVE != 0 : k.resolve != 0);
    const bool pooled = (sched == 5 || sched == 7 || sched == 8);
    int steal_s = k.steal_s;
    if (steal_s < 0) steal_s = opt.mode == PT_MODE_FAST ? 0 : (sched == 8 ? 4 : 16);
    if (opt.mode != PT_MODE_FAST && steal_s == 0) steal_s = (sched == 8 ? 4 : 16); /* strict builds sum in sample order */
    if (resolve) steal_s = 0; /* the resolve kernel sums in sample order in every mode: one round, no table, no shared sums */
    int pool_cap = k.pool_cap;
    /* static shared memory is capped at 48 KB: uniform block + per-warp table + per-warp pool of 50-word slots */
    while (sched == 8 && 16388 + 4 * 4 * ((steal_s == 0 ? 96 : 96 * steal_s) + 50 * pool_cap) > 49152 && pool_cap > 4) pool_cap--;
    int min_blocks = k.min_blocks;
    if (min_blocks < 0) {
        if (opt.mode != PT_MODE_FAST) min_blocks = 4;
        else if (sched == 8) min_blocks = 5;                    /* 16 KB + 26 KB of shared memory per CTA */
        else min_blocks = (pooled && steal_s > 8) ? 5 : 6;      /* 85 registers: +12 % over 4 CTAs/SM on scene1 */
    }
    src += "#define PT_SCHED " + std::to_string(sched) + "\n";
    if (sdf_words > 1) src += "#define PT_SDF_WORDS " + std::to_string(sdf_words) + "\n";
    src += "#define PT_SDF_REPS " + std::to_string(k.sdf_reps) + "\n";
    src += "#define PT_FEED_T " + std::to_string(k.feed_t) + "\n";
    src += "#define PT_REGEN_T " + std::to_string(k.regen_t) + "\n";
    src += "#define PT_STEAL_S " + std::to_string(steal_s) + "\n";
    src += "#define PT_MIN_BLOCKS " + std::to_string(min_blocks) + "\n";
    src += "#define PT_NO_UNROLL " + std::to_string(no_unroll) + "\n";
    src += "#define PT_POOL_CAP " + std::to_string(pool_cap) + "\n";
    src += "#define PT_POOL_MIN " + std::to_string(k.pool_min) + "\n";
    src += "#define PT_WF_REFILL " + std::to_string(k.wf_refill) + "\n";
    src += "#define PT_BVH_WHILE_WHILE " + std::to_string(k.bvh_while_while) + "\n";
    {   /* the heavy phase needs the scan (no BVH), at most 32 boxes + lenses + cyclides, and the phase machine */
        int heavy_min = k.heavy_min < 0 ? PT_DEFAULT_HEAVY_MIN : k.heavy_min;
        const int n_heavy = opt.counts[2] + opt.counts[3] + opt.counts[4];
        if (sched != 5 || opt.bvh || n_heavy == 0 || n_heavy > 32 || !opt.bake_counts) heavy_min = 0;
        src += "#define PT_HEAVY_MIN " + std::to_string(heavy_min) + "\n";
    }
    if (pregen) src += "#define PT_PREGEN 1\n";
    if (resolve) src += "#define PT_RESOLVE 1\n";
    if (k.pathcolor_unroll < 0 ? PT_DEFAULT_PATHCOLOR_UNROLL_WITH_PREGEN && pregen && !resolve : k.pathcolor_unroll != 0) src += "#define PT_PATHCOLOR_UNROLL 1\n";
    if (k.stats) src += "#define PT_STATS 1\n";
    if (k.sin_poly_every > 0 && opt.mode == PT_MODE_FAST) src += "#define PT_SIN_POLY_EVERY " + std::to_string(k.sin_poly_every) + "\n";
    src += opt.wavefront ? "#include \"pt_wavefront.cuh\"\n" : "#include \"pt_kernel.cuh\"\n";
    src += sdf_unit;
    src += "\nPT_DEFINE_RENDER_KERNEL(pt_render_jit)\n";
    if (pregen) src += "PT_DEFINE_GEN_KERNEL(pt_gen_jit)\n";
    if (resolve) src += "PT_DEFINE_RESOLVE_KERNEL(pt_resolve_jit)\n";
    if (has_sdf) src += "PT_DEFINE_SDF_EVAL_KERNEL(pt_sdf_eval_jit)\n";
    if (opt.wavefront) src += "PT_DEFINE_WAVEFRONT_KERNELS\n";
    return src;
}

bool pt_jit_uses_resolve(const std::string& source) { return source.find("#define PT_RESOLVE 1\n") != std::string::npos; }
bool pt_jit_uses_pregen(const std::string& source) { return source.find("#define PT_PREGEN 1\n") != std::string::npos; }

int pt_jit_compile(const std::string& sdf_unit, const PtJitOptions& opt, std::vector<char>* cubin, std::string* log) {
    std::call_once(g_once, load_nvrtc);
    if (!g_nvrtc.h) { *log = g_load_error; return PT_ERR_COMPILE; }
    const std::string src = pt_jit_source(sdf_unit, opt);

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
