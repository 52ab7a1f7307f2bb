/* dsrt_legacy.h -- the reference's own host->CUDA symbol, served by libdsrt.so.
 *
 *   extern "C" void gpu_render_scene(const GPUScene& scene, int width, int height);
 *       declared  /root/reference/src/main.cpp:24-25
 *       defined   /root/reference/src/gpu_render.cu:1037-1108
 *
 * A C++ reference parameter is passed as a pointer, so the symbol below has the same ABI.
 * DsrtLegacyScene mirrors the byte layout of the reference's GPUScene
 * (/root/reference/inc/gpu_scene.h:116-173, 384 bytes; offsets checked in compat_shim.cpp):
 * a host struct holding DEVICE pointers produced by the reference's build_gpu_scene().
 * Side effects are the reference's: the frame is written to "image_gpu.ppm" in the current
 * directory, errors go to stderr and the call returns (gpu_render.cu:1052-1107).
 *
 * Linking the reference's unmodified main.cpp + gpu_scene_builder.cpp against libdsrt.so
 * instead of its gpu_render.cu therefore swaps in the B200 renderer with no source change;
 * INTEGRATION.md shows that link line and the deeper binding that also retires the
 * per-frame CPU BVH build.
 */
#ifndef DSRT_LEGACY_H
#define DSRT_LEGACY_H

#include <stdint.h>
#include "dsrt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct DsrtLegacyScene {
    const DsrtSphere*   spheres;              /*   0 */
    int32_t             num_spheres;          /*   8 */
    int32_t             pad_sph0, pad_sph1;
    const DsrtTriangle* triangles;            /*  24 */
    const int32_t*      tri_indices;          /*  32  reference BVH leaf order, unused here */
    int32_t             num_triangles;        /*  40 */
    int32_t             pad_geo;
    void*               bvh_nodes;            /*  48  reference 40-byte binary nodes, unused here */
    int32_t             num_bvh_nodes;        /*  56 */
    int32_t*            bvh_tri_indices;      /*  64 */
    const DsrtMaterial* materials;            /*  72 */
    int32_t             num_materials;        /*  80 */
    int32_t             pad_mat0, pad_mat1;
    const DsrtTextureHeader* textures;        /*  96 */
    int32_t             num_textures;         /* 104 */
    int32_t             pad_tex0, pad_tex1;
    const float*        texture_pool;         /* 120 */
    int32_t             texture_pool_floats;  /* 128 */
    int32_t             pad_pool0, pad_pool1;
    DsrtCamera          camera;               /* 140 */
    int32_t             sky_type;             /* 244 */
    int32_t             env_tex_id;
    int32_t             pad_sky0, pad_sky1;
    DsrtVec3            sky_solid, sky_top, sky_bottom;   /* 260 */
    DsrtRenderParams    params;               /* 296 */
    uint64_t            seed;                 /* 344 */
    uint8_t             sun_enabled;          /* 352  (C++ bool) */
    DsrtVec3            sun_dir;              /* 356 */
    DsrtVec3            sun_radiance;         /* 368 */
} DsrtLegacyScene;

void gpu_render_scene(const DsrtLegacyScene* scene, int width, int height);

/* The scene half of gpu_render_scene on its own: builds the B200 scene from a reference GPUScene (device arrays,
 * consumed in place; see dsrt_scene_upload_device).  reference_flat_boxes != 0 also reproduces a quirk that depends on
 * the reference's OWN BVH, which the GPUScene carries (bvh_nodes, tri_indices): its box test rejects zero-thickness
 * boxes (src/gpu_render.cu:312: `if (t_max <= t_min) return false`), so the triangles of a reference leaf whose box is
 * flat along an axis -- axis-aligned coplanar panels grouped by its median split, src/gpu_scene_builder.cpp:343-442 --
 * are invisible in the reference.  With the switch those triangles are made unhittable here too (their number is
 * returned through masked_triangles, may be NULL); without it this path renders them.  gpu_render_scene turns the
 * switch on when the environment has DSRT_REF_FLATBOX=1. */
struct dsrt_ctx;
int dsrt_scene_upload_legacy(struct dsrt_ctx* ctx, const DsrtLegacyScene* scene, int reference_flat_boxes, int* masked_triangles);

#ifdef __cplusplus
}
#endif

#endif /* DSRT_LEGACY_H */
