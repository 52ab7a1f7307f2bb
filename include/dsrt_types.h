/* dsrt_types.h -- plain-old-data records that cross the host <-> CUDA boundary.
 *
 * These are the B200 path's own declarations of the records the reference passes
 * between its host code and its CUDA renderer.  They are byte-compatible with the
 * reference so a reference-side caller can hand its arrays over unchanged:
 *
 *   DsrtTriangle       <-> GPUTriangle       /root/reference/inc/gpu_scene.h:57-72   (116 B, align 4)
 *   DsrtMaterial       <-> GPUMaterial       /root/reference/inc/gpu_scene.h:34-45   ( 48 B)
 *   DsrtSphere         <-> GPUSphere         /root/reference/inc/gpu_scene.h:50-55   ( 24 B)
 *   DsrtTextureHeader  <-> GPUTextureHeader  /root/reference/inc/gpu_scene.h:28-32   ( 12 B)
 *   DsrtCamera         <-> GPUCamera         /root/reference/inc/camera.h:13-30      (104 B)
 *   DsrtRenderParams   <-> GPURenderParams   /root/reference/inc/gpu_scene.h:96-111  ( 48 B)
 *
 * Plain C (no CUDA vector types) so that cgo / ctypes / JNI stubs can mirror them.
 */
#ifndef DSRT_TYPES_H
#define DSRT_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct DsrtVec3 { float x, y, z; } DsrtVec3;

/* Material kinds: reference enum MaterialType, gpu_scene.h:21-26. */
enum {
    DSRT_MAT_LAMBERTIAN    = 0,
    DSRT_MAT_METAL         = 1,
    DSRT_MAT_DIELECTRIC    = 2,
    DSRT_MAT_DIFFUSE_LIGHT = 3
};

typedef struct DsrtTextureHeader {
    int32_t width;
    int32_t height;
    int32_t offset;      /* float index into the RGB texture pool */
} DsrtTextureHeader;

typedef struct DsrtMaterial {
    int32_t  type;       /* DSRT_MAT_* */
    int32_t  albedo_tex; /* unused by the renderer (the per-triangle id is used) */
    int32_t  pad0, pad1;
    DsrtVec3 albedo;
    DsrtVec3 emissive;
    float    fuzz;       /* metal, clamped to [0,1] at use */
    float    ref_idx;    /* dielectric; <=0 or non-finite means 1.5 */
} DsrtMaterial;

typedef struct DsrtSphere {
    DsrtVec3 center;
    float    radius;
    int32_t  material_id;
    int32_t  pad;
} DsrtSphere;

typedef struct DsrtTriangle {
    DsrtVec3 v0, v1, v2;
    DsrtVec3 n0, n1, n2;      /* the reference loader writes the flat face normal three times */
    DsrtVec3 uv0, uv1, uv2;   /* (u, v, 0) */
    int32_t  material_id;
    int32_t  albedo_tex;      /* texture id of this triangle, -1 if none */
} DsrtTriangle;

typedef struct DsrtCamera {
    DsrtVec3 origin;
    DsrtVec3 lower_left_corner;
    DsrtVec3 horizontal;
    DsrtVec3 vertical;
    DsrtVec3 u, v, w;          /* carried for layout compatibility; the renderer does not read them */
    float    lens_radius;      /* carried; aperture is 0 in the product */
    int32_t  image_width;
    int32_t  image_height;
    int32_t  samples_per_pixel;
    int32_t  max_depth;
} DsrtCamera;

/* rng_mode values.  The reference leaves rng_mode = 0 and always runs its per-pixel LCG
 * (gpu_render.cu:77-80, 990).  Here 0 selects the counter-based Philox4x32-10 stream that
 * the wavefront pipeline is built around, 1 selects a bit-compatible replay of the
 * reference LCG (one sample of every pixel in flight at a time). */
enum {
    DSRT_RNG_PHILOX  = 0,
    DSRT_RNG_REF_LCG = 1
};

typedef struct DsrtRenderParams {
    int32_t img_width;
    int32_t img_height;
    int32_t samples_per_pixel;   /* total spp of the frame: the jitter divides by it (gpu_render.cu:995) */
    int32_t max_depth;           /* <=0 means 12 (gpu_render.cu:723-725) */
    int32_t use_bvh;             /* ignored, as in the reference */
    int32_t rng_mode;            /* DSRT_RNG_* */
    int32_t tile_size;           /* ignored, as in the reference */
    int32_t pad0;
    float   gamma;               /* <=0 means 1 (gpu_render.cu:1043) */
    float   exposure;            /* ignored, as in the reference */
    float   env_rotation;        /* ignored */
    float   pad1;
} DsrtRenderParams;

/* What one dsrt_scene_upload produced. */
typedef struct DsrtSceneStats {
    int32_t num_triangles;
    int32_t num_wide_nodes;      /* 96-byte BVH8 nodes */
    int32_t num_levels;          /* levels of the wide tree */
    int32_t max_leaf_tris;
    float   build_ms;            /* device time of the whole build (CUDA events) */
    float   upload_ms;           /* host->device copy of the triangle array */
    float   sah_cost;            /* surface-area cost of the wide tree, root area = 1 */
    float   scene_lo[3];
    float   scene_hi[3];
    uint64_t node_bytes;
    uint64_t tri_bytes;
} DsrtSceneStats;

/* What the last dsrt_render_frame did.  Ray counts are exact queue sizes. */
typedef struct DsrtFrameStats {
    uint64_t primary_rays;
    uint64_t bounce_rays;
    uint64_t shadow_rays;
    uint64_t paths;              /* pixel samples started */
    uint32_t kernel_launches;    /* kernels this library launched for the frame */
    uint32_t batches;
    float    device_ms;          /* CUDA-event time of the whole frame on the ctx stream */
    float    extend_ms;          /* BVH traversal kernels: k_primary + k_trace, closest and any hit
                                    (0 unless stage timing is on) */
    float    primary_ms;         /* k_primary alone (included in extend_ms; the sun shadow rays are
                                    traced inside k_trace and have no time of their own) */
    float    shade_ms;           /* k_shade */
    uint64_t culled_primary_rays; /* of primary_rays: pixel samples whose tile (or pixel) provably cannot see a
                                    triangle and that ended as misses without a traversal (k_tile_entries) */
} DsrtFrameStats;

#ifdef __cplusplus
}
#endif

#endif /* DSRT_TYPES_H */
