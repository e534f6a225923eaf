/* pt_render -- headless example host for libpt_cuda.
 *
 * The reference's offscreen mode asks four questions on stdin -- samples, samples per frame, path length, camera
 * shot (host:3970-3978) -- then opens file dialogs for the scene and the output (host:3989,3492).  This CLI takes
 * the same parameters as flags and drives the library through the C ABI only (include/pt_abi.h).
 *
 *   pt_render --scene scenes/scene0.json --width 512 --height 512 --spp 64 --spf 8 --path-length 5 --shot 1 \
 *             [--fast] [--wavefront] [--jit 0|1|2] [--device 0] [--out render.pfm|render.exr|render.ppm] [--tonemap 3]
 *             [--resume checkpoint.pfm --done-samples N]      continue a run saved with --out checkpoint.pfm
 *             [--gpus N]      split the samples over the first N GPUs of the box (pt_multi: one NCCL reduce at the end)
 *             [--opt key=value]  a tuning option of the kernels (pt_set_option: sched, sdf_reps, ...); repeatable
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "pt_abi.h"

static int die(const char* what, pt_ctx* ctx) {
    fprintf(stderr, "pt_render: %s: %s\n", what, pt_last_error(ctx));
    return 1;
}

int main(int argc, char** argv) {
    std::string scene_path = "scenes/scene0.json", out_path;
    int width = 1280, height = 720, spp = 1000, spf = 1, path_length = 5, shot = 1, device = 0, tonemap = 3; /* host:30-31,1164-1174 */
    int fast = 0, jit = -1, wavefront = 0, done = 0, gpus = 1;
    std::string resume_path;
    std::vector<std::pair<std::string, long long>> opts;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { fprintf(stderr, "pt_render: %s needs a value\n", name); exit(2); }
            return argv[++i];
        };
        if (a == "--scene") scene_path = next("--scene");
        else if (a == "--width") width = atoi(next("--width"));
        else if (a == "--height") height = atoi(next("--height"));
        else if (a == "--spp") spp = atoi(next("--spp"));
        else if (a == "--spf") spf = atoi(next("--spf"));
        else if (a == "--path-length") path_length = atoi(next("--path-length"));
        else if (a == "--shot") shot = atoi(next("--shot"));
        else if (a == "--device") device = atoi(next("--device"));
        else if (a == "--tonemap") tonemap = atoi(next("--tonemap"));
        else if (a == "--out") out_path = next("--out");
        else if (a == "--jit") jit = atoi(next("--jit"));
        else if (a == "--gpus") gpus = atoi(next("--gpus"));
        else if (a == "--fast") fast = 1;
        else if (a == "--wavefront") wavefront = 1;
        else if (a == "--resume") resume_path = next("--resume");            /* PFM checkpoint written by --out x.pfm */
        else if (a == "--done-samples") done = atoi(next("--done-samples"));   /* samples per pixel already in it */
        else if (a == "--strict") fast = 0;
        else if (a == "--opt") {
            const std::string kv = next("--opt");
            const size_t eq = kv.find('=');
            if (eq == std::string::npos) { fprintf(stderr, "pt_render: --opt wants key=value\n"); return 2; }
            opts.emplace_back(kv.substr(0, eq), atoll(kv.c_str() + eq + 1));
        }
        else { fprintf(stderr, "pt_render: unknown flag %s\n", a.c_str()); return 2; }
    }
    pt_scene* scene = nullptr;
    if (pt_scene_load_json(scene_path.c_str(), &scene) != PT_OK) return die("load scene", nullptr);
    pt_ubo ubo;
    pt_params params;
    if (pt_scene_pack_ubo(scene, &ubo) != PT_OK) return die("pack ubo", nullptr);
    if (pt_scene_pack_params(scene, shot, width, height, spf, path_length, &params) != PT_OK) return die("pack params", nullptr);
    params.tonemap = tonemap;
    /* surface extensions of the scene file ("bsdf" keys on materials; none in the reference's scenes) */
    pt_surface_ext surface_ext[PT_MAX_SURFACE_EXT];
    const int n_surface_ext = pt_scene_surface_ext(scene, surface_ext, PT_MAX_SURFACE_EXT);
    if (n_surface_ext < 0) return die("surface extensions", nullptr);
    std::vector<const char*> sdf;
    for (int i = 0; i < pt_scene_num_sdf(scene); i++) sdf.push_back(pt_scene_sdf_glsl(scene, i));

    auto write_image = [&](const std::vector<float>& img) -> int {
        const bool ppm = out_path.size() > 4 && out_path.substr(out_path.size() - 4) == ".ppm";
        const bool exr = out_path.size() > 4 && out_path.substr(out_path.size() - 4) == ".exr";
        int rc = ppm ? pt_write_ppm(out_path.c_str(), img.data(), width, height, tonemap)
                     : (exr ? pt_write_exr(out_path.c_str(), img.data(), width, height, 1)
                            : pt_write_pfm(out_path.c_str(), img.data(), width, height, 0));
        if (rc != PT_OK) fprintf(stderr, "pt_render: cannot write %s\n", out_path.c_str());
        return rc;
    };

    if (gpus > 1) { /* sample-split over devices device .. device + gpus - 1 */
        std::vector<int> devs;
        for (int g = 0; g < gpus; g++) devs.push_back(device + g);
        pt_multi* m = nullptr;
        if (pt_multi_create(devs.data(), gpus, fast ? PT_MODE_FAST : PT_MODE_STRICT, &m) != PT_OK) {
            fprintf(stderr, "pt_render: pt_multi_create: %s\n", pt_multi_last_error(nullptr));
            return 1;
        }
        for (int g = 0; g < gpus; g++) {
            if (jit >= 0) pt_set_jit(pt_multi_ctx(m, g), jit);
            for (auto& o : opts)
                if (pt_set_option(pt_multi_ctx(m, g), o.first.c_str(), o.second) != PT_OK) return die("option", pt_multi_ctx(m, g));
            if (wavefront) pt_set_pipeline(pt_multi_ctx(m, g), PT_PIPE_WAVEFRONT);
        }
        auto t0 = std::chrono::steady_clock::now();
        double secs = 0.0;
        if (pt_multi_set_surface_ext(m, surface_ext, n_surface_ext) != PT_OK ||
            pt_multi_set_scene(m, &ubo, sdf.data(), (int)sdf.size()) != PT_OK || pt_multi_resize(m, width, height) != PT_OK) {
            fprintf(stderr, "pt_render: %s\n", pt_multi_last_error(m));
            return 1;
        }
        auto t1 = std::chrono::steady_clock::now();
        if (pt_multi_render(m, &params, 0, spp, spf, &secs) != PT_OK) {
            fprintf(stderr, "pt_render: %s\n", pt_multi_last_error(m));
            return 1;
        }
        printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, \"spf\": %d, \"path_length\": %d, \"mode\": \"%s\", "
               "\"gpus\": %d, \"compile_s\": %.3f, \"render_s\": %.4f, \"reduce_s\": %.5f, \"msamples_per_s\": %.2f}\n",
               scene_path.c_str(), width, height, spp, spf, path_length, fast ? "fast" : "strict", gpus,
               std::chrono::duration<double>(t1 - t0).count(), secs, pt_multi_reduce_seconds(m),
               (double)width * height * (double)spp / secs / 1e6);
        if (!out_path.empty()) {
            std::vector<float> img((size_t)width * height * 4);
            if (pt_multi_read_xyz(m, img.data(), img.size()) != PT_OK || write_image(img) != PT_OK) return 1;
        }
        pt_multi_destroy(m);
        pt_scene_free(scene);
        return 0;
    }

    pt_ctx* ctx = nullptr;
    if (pt_create(device, &ctx) != PT_OK) return die("create context", nullptr);
    pt_set_mode(ctx, fast ? PT_MODE_FAST : PT_MODE_STRICT);
    if (jit >= 0) pt_set_jit(ctx, jit);
    for (auto& o : opts)
        if (pt_set_option(ctx, o.first.c_str(), o.second) != PT_OK) return die("option", ctx);
    if (wavefront) pt_set_pipeline(ctx, PT_PIPE_WAVEFRONT);
    auto t0 = std::chrono::steady_clock::now();
    if (pt_set_surface_ext(ctx, surface_ext, n_surface_ext) != PT_OK) return die("surface extensions", ctx);
    if (pt_set_scene(ctx, &ubo, sdf.data(), (int)sdf.size()) != PT_OK) return die("set scene", ctx);
    auto t1 = std::chrono::steady_clock::now();
    if (pt_resize(ctx, width, height) != PT_OK) return die("resize", ctx);
    if (!resume_path.empty()) {
        std::vector<float> ck((size_t)width * height * 4);
        if (pt_read_pfm(resume_path.c_str(), ck.data(), width, height) != PT_OK) { fprintf(stderr, "pt_render: cannot read %s\n", resume_path.c_str()); return 1; }
        if (pt_write_xyz(ctx, ck.data(), ck.size()) != PT_OK) return die("upload checkpoint", ctx);
    }
    if (pt_render_resume(ctx, &params, resume_path.empty() ? 0 : done, spp, spf) != PT_OK) return die("render", ctx);
    auto t2 = std::chrono::steady_clock::now();
    float ms = 0.0f;
    long long launches = 0;
    pt_kernel_time(ctx, &ms, &launches);
    const double samples = (double)width * height * (double)(spp / spf * spf);
    printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, \"spf\": %d, \"path_length\": %d, \"mode\": \"%s\", "
           "\"compile_s\": %.3f, \"render_s\": %.3f, \"kernel_ms\": %.3f, \"launches\": %lld, \"msamples_per_s\": %.2f}\n",
           scene_path.c_str(), width, height, spp, spf, path_length, fast ? "fast" : "strict",
           std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count(), ms, launches,
           samples / (ms * 1e-3) / 1e6);
    if (!out_path.empty()) {
        std::vector<float> img((size_t)width * height * 4);
        if (pt_read_xyz(ctx, img.data(), img.size()) != PT_OK) return die("read back", ctx);
        if (write_image(img) != PT_OK) return 1;
    }
    pt_destroy(ctx);
    pt_scene_free(scene);
    return 0;
}
