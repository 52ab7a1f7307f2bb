/* dsrt.h -- C ABI of the B200 render path (libdsrt.so).
 *
 * This is the boundary a reference-side caller binds instead of the reference's three
 * host->CUDA calls:
 *
 *   GPUScene build_gpu_scene(const hittable_list&, const camera&, const vec3& sun_dir_model)
 *                                            /root/reference/inc/gpu_scene_builder.h:72
 *                                            /root/reference/src/gpu_scene_builder.cpp:464-601
 *   extern "C" void gpu_render_scene(const GPUScene&, int width, int height)
 *                                            /root/reference/src/main.cpp:24-25
 *                                            /root/reference/src/gpu_render.cu:1037-1108
 *   void free_gpu_scene(GPUScene&)           /root/reference/inc/gpu_scene_builder.h:73
 *                                            /root/reference/src/gpu_scene_builder.cpp:603-626
 *
 * Conventions: every call returns 0 on success and a negative DSRT_E* code on failure; no
 * exception crosses the boundary; dsrt_last_error() gives the text.  A ctx owns all of its
 * device memory and one non-default CUDA stream, lives on one GPU, and is used by one host
 * thread at a time.  All pointers are plain host pointers unless the name says "device".
 * There is no CPU fallback: without a usable CUDA device dsrt_create fails.
 */
#ifndef DSRT_H
#define DSRT_H

#include <stddef.h>
#include <stdint.h>
#include "dsrt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DSRT_OK            0
#define DSRT_EINVAL       -1   /* bad argument */
#define DSRT_ECUDA        -2   /* a CUDA runtime call failed */
#define DSRT_ENOSCENE     -3   /* render/trace before dsrt_scene_upload */
#define DSRT_EUNSUPPORTED -4   /* not available in this build / on this box (libnccl missing) */
#define DSRT_EIO          -5   /* file could not be read/written */
#define DSRT_ENOMEM       -6

typedef struct dsrt_ctx dsrt_ctx;

/* ---- context ------------------------------------------------------------------------ */

/* Creates a context on CUDA device `device`.  Replaces nothing in the reference (which uses
 * the current device and the default stream, gpu_render.cu:1052-1097). */
int  dsrt_create(dsrt_ctx** out, int device);
void dsrt_destroy(dsrt_ctx* ctx);
/* Text of the last error of `ctx`; with ctx == NULL the last dsrt_create failure. */
const char* dsrt_last_error(const dsrt_ctx* ctx);
/* The ctx stream as a cudaStream_t, for callers that order their own work after ours. */
void* dsrt_stream(dsrt_ctx* ctx);

/* ---- scene: replaces build_gpu_scene()'s flatten + CPU BVH + 7 uploads ----------------
 * Copies the reference-layout arrays to the device and builds, on the GPU, a Morton-sorted
 * LBVH that is collapsed by surface area into an 8-wide BVH with quantised child boxes
 * (96-byte nodes) and 48-byte triangle records in leaf order.  Done once per mesh, not per
 * frame.  Spheres (never instantiated by the reference's main(), but part of GPUScene) are
 * intersected by a brute-force loop after the triangles exactly as the reference's scene_hit
 * does (gpu_render.cu:530-546); emissive ones switch the diffuse bounce to the reference's
 * light/BRDF mixture sampling (gpu_render.cu:838-932).  The parity hooks report a sphere hit
 * as tri = -2 - sphere index. */
int dsrt_scene_upload(dsrt_ctx* ctx,
                      const DsrtTriangle* tris, int ntris,
                      const DsrtMaterial* mats, int nmats,
                      const DsrtSphere* spheres, int nspheres,
                      const DsrtTextureHeader* textures, int ntextures,
                      const float* texture_pool, int texture_pool_floats);
/* Same, from arrays that already live in the memory of ctx's GPU (what the reference's
 * GPUScene holds after its own build_gpu_scene, inc/gpu_scene.h:116-173): the triangle array is
 * consumed in place, no host round trip; material ids are range-checked by a kernel.  The few
 * spheres (and the material table, to find the emissive ones) are read back and packed on the
 * host as in dsrt_scene_upload. */
int dsrt_scene_upload_device(dsrt_ctx* ctx,
                             const DsrtTriangle* tris_device, int ntris,
                             const DsrtMaterial* mats_device, int nmats,
                             const DsrtSphere* spheres_device, int nspheres,
                             const DsrtTextureHeader* textures_device, int ntextures,
                             const float* texture_pool_device, int texture_pool_floats);
int dsrt_scene_stats(dsrt_ctx* ctx, DsrtSceneStats* out);
/* Debug/verification: copies the wide BVH back (96-byte nodes, 48-byte triangle records). */
int dsrt_scene_download(dsrt_ctx* ctx, void* nodes, size_t node_bytes, void* tri_records, size_t tri_bytes);

/* ---- render: replaces gpu_render_scene() ---------------------------------------------
 * Renders samples [sample_begin, sample_end) of every pixel of `rect` (x0,y0,x1,y1, half
 * open, y = 0 is the bottom row as in the reference; NULL = whole image) and ADDS the
 * per-sample clamped radiance into the ctx's float accumulation buffer (W*H*3, pixel-major,
 * row 0 = bottom).  The buffer is zeroed first unless `keep_accum` is non-zero (progressive
 * refinement: later sample ranges of the same frame pass keep_accum = 1).  params->samples_per_pixel is the frame's total spp
 * (the sample jitter divides by it); max_depth, gamma and rng_mode are honoured, the other
 * fields are carried but unused, exactly as in the reference kernel.
 *
 * sun_dir is the reference's scene.sun_dir (model origin -> Sun; the estimator lights from
 * -sun_dir, gpu_render.cu:802-806), sun_radiance its scene.sun_radiance. */
int dsrt_render_frame(dsrt_ctx* ctx,
                      const DsrtCamera* camera,
                      const float sun_dir[3], const float sun_radiance[3], int sun_enabled,
                      const DsrtRenderParams* params, uint64_t seed,
                      int sample_begin, int sample_end,
                      const int rect[4], int keep_accum);
int dsrt_frame_stats(dsrt_ctx* ctx, DsrtFrameStats* out);

/* The same frame call without the wait.  A ctx has DSRT_MAX_SLOTS frame slots, each with its own
 * stream, image buffers and work memory (allocated at first use); dsrt_render_frame is slot 0
 * followed by a wait.  A frame enqueued here runs asynchronously; the host never synchronises
 * inside a frame (queue sizes live in device memory and the launch sequence is fixed), so a caller
 * that alternates two slots keeps the GPU busy across frames: the latency-bound end of frame k
 * (few rays left, tone map, copy to the host) runs under the camera rays of frame k+1.  The host
 * arguments are consumed before the call returns.  dsrt_readback_rgb8_async enqueues the tone map
 * and the copy of the slot's frame to `dst` behind it (dst should be page-locked, dsrt_host_alloc,
 * for the copy to be asynchronous; it must stay valid until dsrt_frame_wait).  dsrt_frame_wait
 * blocks until everything enqueued on the slot is done and returns the frame's statistics (out
 * may be NULL).  A slot takes one frame at a time: enqueueing on a slot that has not been waited
 * for is DSRT_EINVAL.  The reference has no counterpart: it is synchronous and re-uploads the
 * scene per frame (src/main.cpp:404-428, src/gpu_render.cu:1052-1097). */
#define DSRT_MAX_SLOTS 4
int dsrt_render_frame_async(dsrt_ctx* ctx, int slot,
                            const DsrtCamera* camera,
                            const float sun_dir[3], const float sun_radiance[3], int sun_enabled,
                            const DsrtRenderParams* params, uint64_t seed,
                            int sample_begin, int sample_end,
                            const int rect[4], int keep_accum);
int dsrt_readback_rgb8_async(dsrt_ctx* ctx, int slot, int total_spp, float gamma, unsigned char* dst, size_t bytes);
int dsrt_frame_wait(dsrt_ctx* ctx, int slot, DsrtFrameStats* out);
/* 1 = time the traversal / shading kernels separately with CUDA events (adds launches of
 * event records only; default 0). */
int dsrt_set_stage_timing(dsrt_ctx* ctx, int enabled);
/* Samples of a pixel that are in flight together in Philox mode (0 = choose from the image
 * size).  The LCG replay always uses 1. */
int dsrt_set_samples_per_batch(dsrt_ctx* ctx, int samples);
/* Binary-tree stage of the GPU BVH build used by the NEXT dsrt_scene_upload: 0 = PLOC (parallel
 * locally-ordered clustering, the default: lowest SAH cost), 1 = binary radix tree over the Morton
 * order (fastest build).  Both feed the same SAH-optimal collapse into the 8-wide BVH, and both
 * replace build_bvh_recursive (/root/reference/src/gpu_scene_builder.cpp:343-442). */
#define DSRT_BVH_PLOC 0
#define DSRT_BVH_LBVH 1
int dsrt_set_bvh_builder(dsrt_ctx* ctx, int builder);

/* Host arithmetic, no device needed (exposed for tests): the inclusive pixel rectangle
 * out = {x0, y0, x1, y1} (y = 0 is the bottom row, as in the kernel) outside of which no camera ray
 * of the reference camera model (gpu_render.cu:952-967) can hit the axis-aligned box [lo, hi].
 * Returns 1 when the rectangle is a real bound, 0 when it is the whole frame because a box corner
 * is beside or behind the camera.  dsrt_render_frame uses it to end such paths before generating
 * their rays. */
int dsrt_box_screen_bounds(const DsrtCamera* camera, const float lo[3], const float hi[3],
                           int width, int height, int out[4]);

/* Page-locked host memory for frame read-backs (optional: every read-back entry accepts any host
 * pointer; a page-locked destination lets the copy run at full PCIe rate and overlap with the writer
 * thread's work).  The reference reads back into pageable memory (gpu_render.cu:1086-1092). */
int dsrt_host_alloc(void** ptr, size_t bytes);
void dsrt_host_free(void* ptr);

/* Tone-map the accumulation buffer exactly as the reference kernel's tail
 * (gpu_render.cu:1003-1030: mean over total_spp, max 0, min 10, powf(c, 1/gamma), clamp,
 * (uchar)(255.99f*c), row flip) and copy W*H*3 bytes, top row first, to `dst`. */
int dsrt_readback_rgb8(dsrt_ctx* ctx, int total_spp, float gamma, unsigned char* dst, size_t bytes);
/* Raw accumulation buffer (sum over samples of clamp01(L)), W*H*3 floats, row 0 = bottom. */
int dsrt_readback_accum(dsrt_ctx* ctx, float* dst, size_t bytes);
/* Device pointer / element count of the accumulation buffer, for an NCCL sum-reduce issued
 * by the caller (single-frame sample sharding). */
int dsrt_accum_device(dsrt_ctx* ctx, void** device_ptr, size_t* num_floats);
/* The exchange step of a sample-sharded frame: ONE ncclReduce(sum, float) of this
 * ctx's accumulation buffer to rank `root` of `nccl_comm` (an ncclComm_t the caller created for its
 * ranks, passed as void* so that this header needs no NCCL types), enqueued on the ctx stream and
 * waited for.  Every rank of the communicator calls it once per frame after dsrt_render_frame with
 * its own [sample_begin, sample_end); the root then tone-maps with dsrt_readback_rgb8.  libnccl is
 * looked up at the first call (dlopen), so the library itself does not depend on it;
 * DSRT_EUNSUPPORTED when it cannot be found. */
int dsrt_reduce_accum(dsrt_ctx* ctx, void* nccl_comm, int root);
/* Makes the accumulation buffer exist at width x height and zeroes it: what a rank whose
 * dsrt_render_frame failed does before it still takes part in dsrt_reduce_accum, so that the
 * collective neither dead-locks nor sums a partial frame. */
int dsrt_clear_accum(dsrt_ctx* ctx, int width, int height);
/* Replace the accumulation buffer contents from host memory (root of a host-side reduce). */
int dsrt_upload_accum(dsrt_ctx* ctx, const float* src, size_t bytes);

/* ---- parity hooks ----------------------------------------------------------------------
 * Closest hit of n rays: rays = n x {ox,oy,oz,tmin, dx,dy,dz,tmax} floats.  Outputs per ray:
 * tri (index into the uploaded triangle array, -1 = miss), t, u, v (barycentrics), bit-equal
 * to hit_triangle_index()'s arithmetic (gpu_render.cu:322-353). */
int dsrt_trace_closest(dsrt_ctx* ctx, const float* rays, int n,
                       int32_t* tri, float* t, float* u, float* v);
/* Any hit in [tmin,tmax] (the sun shadow-ray query, gpu_render.cu:812-816): occluded[i]=0/1. */
int dsrt_trace_occluded(dsrt_ctx* ctx, const float* rays, int n, uint8_t* occluded);
/* Primary rays of the reference camera with a fixed sub-pixel jitter (jx,jy) for every
 * pixel (gpu_render.cu:941-968), traced for the first hit.  Outputs are W*H, index
 * x + y*W with y = 0 the bottom row. */
int dsrt_first_hits(dsrt_ctx* ctx, const DsrtCamera* camera, int width, int height,
                    float jx, float jy, int32_t* tri, float* t, float* u, float* v);
/* The first hits of global sample `sample` of every pixel as the render pipeline itself finds
 * them: the per-tile start nodes and tile culling of k_tile_entries and the fused ray generation +
 * traversal of k_primary, with the estimator's own jitter for params->rng_mode (gpu_render.cu:
 * 990-998; the LCG replay can show sample 0 only).  Outputs are W*H (from params), index x + y*W
 * with y = 0 the bottom row; a pixel without a hit has tri = -1, t = 1e9.  leaf_node / leaf_record
 * (may be NULL) say where the hit was found: the index of the 8-wide BVH node whose leaf child held
 * the triangle, and the triangle's position in the leaf-ordered record array (see
 * dsrt_scene_download) -- north_star's "(triangle index, BVH leaf)". */
int dsrt_primary_hits(dsrt_ctx* ctx, const DsrtCamera* camera, const DsrtRenderParams* params, uint64_t seed,
                      int sample, int32_t* tri, float* t, float* u, float* v,
                      int32_t* leaf_node, int32_t* leaf_record);
/* Device-resident closest-hit / any-hit over rays already in device memory (bench of the
 * traversal kernel alone).  rays_device: n x 8 floats; hits_device: n x 4 floats
 * {t,u,v,bits(tri)}; occluded_device: n bytes. */
int dsrt_trace_closest_device(dsrt_ctx* ctx, const void* rays_device, int n, void* hits_device);
int dsrt_trace_occluded_device(dsrt_ctx* ctx, const void* rays_device, int n, void* occluded_device);

/* ---- host side of the path (no GPU needed) ---------------------------------------------
 * Pose file: reference read_pose_file(), /root/reference/src/main.cpp:139-173.  Returns the
 * number of poses found (>= 0) or DSRT_EIO; writes at most max_poses x 9 doubles
 * {cam xyz, model xyz, yaw, pitch, roll}.  yaw/pitch/roll pass through float as in the
 * reference (PoseEntry::model_euler_deg is a float vec3). */
int dsrt_read_pose_file(const char* path, double* poses9, int max_poses);
/* World -> model frame in double, then cast: main.cpp:334-357 with rotate_yaw_deg_d
 * (main.cpp:118-128).  sep_m is |cam - model|; frames with sep_m < 1 are skipped by the
 * reference (main.cpp:340-345). */
void dsrt_pose_to_model_frame(const double pose9[9], float cam_in_model[3],
                              float sun_dir_model[3], double* sep_m);
/* point_camera_at() + camera::initialize(): main.cpp:178-187, inc/camera.h:91-116
 * (vup = +y, lookat = origin, focus = |lookfrom|, aperture 0, float math). */
void dsrt_camera_look_at_origin(const float lookfrom[3], float vfov_deg, int width, int height,
                                int spp, int max_depth, DsrtCamera* out);
/* Deterministic synthetic "station" mesh standing in for the ISS OBJ.
 * Writes ntris triangles and up to max_mats materials; returns the number of materials.
 * diffuse_only != 0 gives lambertian-only materials (BASELINE config 5). */
int dsrt_synth_mesh(int ntris, uint32_t seed, int diffuse_only,
                    DsrtTriangle* tris, DsrtMaterial* mats, int max_mats);
/* OBJ + MTL (+ map_Kd textures) -> the arrays dsrt_scene_upload takes.  Replaces the reference's
 * scene ingestion: triangle_mesh (inc/triangle_mesh.h:75-255: OBJ/MTL parser, MTL -> material
 * mapping, fan triangulation, flat normals) followed by the flattening half of build_gpu_scene
 * (src/gpu_scene_builder.cpp:72-138, 199-282, 518-551: material table by first use, texture
 * registry, sRGB -> linear float pool).  `scale` multiplies the vertex positions in double, as
 * triangle_mesh's constructor argument does.  Returns DSRT_EIO when the OBJ cannot be opened
 * (the reference silently continues with an empty mesh), DSRT_ENOMEM / DSRT_EIO when the file is
 * too large or malformed (no exception crosses this boundary).  Texture files are decoded
 * in-process (PNG, JPEG -- baseline and progressive, with stb_image's integer arithmetic, so the
 * texels are the reference's --, binary PNM, BMP, TGA); a missing or undecodable file becomes a
 * 1x1 white texture, the reference's own failure path, and is counted in the view. */
typedef struct dsrt_mesh dsrt_mesh;
typedef struct DsrtMeshView {
    const DsrtTriangle*      triangles;      int32_t num_triangles;
    const DsrtMaterial*      materials;      int32_t num_materials;
    const DsrtTextureHeader* textures;       int32_t num_textures;
    const float*             texture_pool;   int32_t texture_pool_floats;
    int32_t num_vertices, num_uvs;           /* `v` and `vt` records read */
    int32_t skipped_faces;                   /* triangles dropped for an out-of-range vertex index */
    int32_t textures_failed;                 /* textures replaced by 1x1 white */
    int32_t textures_flipped;                /* 1 = pool rows are bottom-up (inc/texture.h:133 side effect) */
    int32_t textures_unsupported;            /* of textures_failed: the file exists but is in a format this loader does
                                                not decode (the reference's stb_image also reads GIF, PSD, HDR, PIC):
                                                the picture WOULD differ from the reference's; the CLI refuses to render */
} DsrtMeshView;
int  dsrt_mesh_load_obj(const char* obj_path, double scale, dsrt_mesh** out);
int  dsrt_mesh_view(const dsrt_mesh* mesh, DsrtMeshView* view);     /* pointers valid until dsrt_mesh_free */
void dsrt_mesh_free(dsrt_mesh* mesh);

/* Binary P6 exactly as gpu_render.cu:1100-1107 writes it; PNG (8-bit RGB, zlib) replaces
 * the ImageMagick subprocess of main.cpp:28-36. */
int dsrt_write_ppm(const char* path, const unsigned char* rgb, int width, int height);
int dsrt_write_png(const char* path, const unsigned char* rgb, int width, int height);

/* Library build identification ("dsrt-b200 <n> sm_100a"). */
const char* dsrt_version(void);

#ifdef __cplusplus
}
#endif

#endif /* DSRT_H */
