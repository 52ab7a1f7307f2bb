// compat_shim.cpp -- the reference's `extern "C" gpu_render_scene` served by the B200 path.
//
// Replaces /root/reference/src/gpu_render.cu:1037-1108 at link time: same symbol, same
// argument meaning, same side effect (image_gpu.ppm in the CWD), same error behaviour
// (message on stderr, return).  The scene arrives as the reference's GPUScene: a host
// struct of device pointers.  The triangles are consumed where they lie on the device.
//
// Environment: DSRT_RNG=lcg selects the bit-compatible replay of the reference's per-pixel
// LCG; default is the counter-based Philox stream.  DSRT_DEVICE picks the GPU.
// DSRT_REF_FLATBOX=1 reproduces the reference's invisible zero-thickness leaves (dsrt_legacy.h).
#include "dsrt.h"
#include "dsrt_legacy.h"

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static_assert(sizeof(DsrtCamera) == 104, "GPUCamera layout");
static_assert(sizeof(DsrtRenderParams) == 48, "GPURenderParams layout");
static_assert(offsetof(DsrtLegacyScene, triangles) == 24, "GPUScene.triangles");
static_assert(offsetof(DsrtLegacyScene, materials) == 72, "GPUScene.materials");
static_assert(offsetof(DsrtLegacyScene, textures) == 96, "GPUScene.textures");
static_assert(offsetof(DsrtLegacyScene, texture_pool) == 120, "GPUScene.texture_pool");
static_assert(offsetof(DsrtLegacyScene, camera) == 140, "GPUScene.camera");
static_assert(offsetof(DsrtLegacyScene, params) == 296, "GPUScene.params");
static_assert(offsetof(DsrtLegacyScene, seed) == 344, "GPUScene.seed");
static_assert(offsetof(DsrtLegacyScene, sun_enabled) == 352, "GPUScene.sun_enabled");
static_assert(offsetof(DsrtLegacyScene, sun_dir) == 356, "GPUScene.sun_dir");
static_assert(offsetof(DsrtLegacyScene, sun_radiance) == 368, "GPUScene.sun_radiance");
static_assert(sizeof(DsrtLegacyScene) == 384, "GPUScene size");

namespace {

dsrt_ctx* shim_ctx() {
    static dsrt_ctx* ctx = nullptr;
    if (!ctx) {
        const char* dev = std::getenv("DSRT_DEVICE");
        if (dsrt_create(&ctx, dev ? std::atoi(dev) : 0) != DSRT_OK) {
            std::fprintf(stderr, "dsrt: %s\n", dsrt_last_error(nullptr));
            ctx = nullptr;
        }
    }
    return ctx;
}

}  // namespace

extern "C" void gpu_render_scene(const DsrtLegacyScene* scene, int width, int height) {
    if (!scene) { std::fprintf(stderr, "gpu_render_scene: null scene\n"); return; }
    dsrt_ctx* ctx = shim_ctx();
    if (!ctx) return;

    // DSRT_REF_FLATBOX=1: triangles the reference cannot see because its own BVH put them into a zero-thickness leaf
    // box stay invisible here too (include/dsrt_legacy.h)
    const char* flat = std::getenv("DSRT_REF_FLATBOX");
    int masked = 0;
    int rc = dsrt_scene_upload_legacy(ctx, scene, (flat && std::atoi(flat) != 0) ? 1 : 0, &masked);
    if (rc == DSRT_OK && masked > 0) std::fprintf(stderr, "gpu_render_scene: %d triangles in flat reference leaves masked (DSRT_REF_FLATBOX)\n", masked);
    if (rc != DSRT_OK) { std::fprintf(stderr, "gpu_render_scene: %s\n", dsrt_last_error(ctx)); return; }

    DsrtRenderParams params = scene->params;
    params.img_width = width;
    params.img_height = height;
    const char* rng = std::getenv("DSRT_RNG");
    params.rng_mode = (rng && std::strcmp(rng, "lcg") == 0) ? DSRT_RNG_REF_LCG : DSRT_RNG_PHILOX;
    const int spp = params.samples_per_pixel < 1 ? 1 : params.samples_per_pixel;
    params.samples_per_pixel = spp;

    const float sun_dir[3] = {scene->sun_dir.x, scene->sun_dir.y, scene->sun_dir.z};
    const float sun_rad[3] = {scene->sun_radiance.x, scene->sun_radiance.y, scene->sun_radiance.z};
    rc = dsrt_render_frame(ctx, &scene->camera, sun_dir, sun_rad, scene->sun_enabled ? 1 : 0,
                           &params, scene->seed, 0, spp, nullptr, 0);
    if (rc != DSRT_OK) { std::fprintf(stderr, "render_kernel failed: %s\n", dsrt_last_error(ctx)); return; }

    std::vector<unsigned char> fb((size_t)width * (size_t)height * 3);
    rc = dsrt_readback_rgb8(ctx, spp, params.gamma, fb.data(), fb.size());
    if (rc != DSRT_OK) { std::fprintf(stderr, "framebuffer read-back failed: %s\n", dsrt_last_error(ctx)); return; }
    if (dsrt_write_ppm("image_gpu.ppm", fb.data(), width, height) != DSRT_OK)
        std::fprintf(stderr, "Failed to open image_gpu.ppm for writing\n");
}
