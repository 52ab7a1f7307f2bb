/* oracle.h -- CPU restatement of the reference's render path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library, and only as the checker -- never as the thing measured as the product
 * or shipped.  The product (deep-space-ray-tracer_b200/) does not link or call it.
 *
 * Parity status: pinned.  The reference ships no tests or golden vectors for this path
 * (SURVEY.md section 4), so the restatement is pinned against outputs of the reference
 * itself: (a) its host functions for pose/camera (tests/golden/pose_camera_golden.json,
 * produced by oracle/gen_golden.py from the compiled reference), and (b) the unmodified
 * reference kernel run on the GPU box (oracle/_ref/libref_gpu.so; first hits bit-exact,
 * tests/test_gpu_parity.py).  libm differs between glibc and CUDA (sinf/cosf/powf), so
 * full images from this oracle agree with the GPU statistically, first hits bit for bit.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrScene OrScene;

/* tris: n x 29 floats in the reference GPUTriangle layout (gpu_scene.h:57-72); mats: nmats x 48
 * bytes in the GPUMaterial layout (gpu_scene.h:34-45).  Builds the reference's median-split
 * binary BVH (gpu_scene_builder.cpp:343-459). */
OrScene* or_scene_create(const float* tris, int n, const void* mats, int nmats);
/* spheres: n x 24 bytes in the GPUSphere layout (gpu_scene.h:50-55); textures: n x 12-byte
 * GPUTextureHeader + the float RGB pool they index (gpu_scene.h:28-32).  With spheres the
 * hooks below report a sphere hit as tri = -2 - sphere index (the reference's HitRecord holds
 * tri_index = -1 for them). */
void     or_scene_set_spheres(OrScene* s, const void* spheres, int n);
void     or_scene_set_textures(OrScene* s, const void* headers, int n, const float* pool, int pool_floats);
void     or_scene_destroy(OrScene* s);
int      or_scene_num_nodes(const OrScene* s);
int      or_scene_max_leaf(const OrScene* s);

/* cam12 = origin, lower_left_corner, horizontal, vertical (GPUCamera fields the kernel reads). */
void or_first_hits(const OrScene* s, const float cam12[12], int W, int H, float jx, float jy,
                   int32_t* tri, float* t, float* u, float* v);
/* First hit of sample `sample` of every pixel under the estimator's own jitter (gpu_render.cu:990-998):
 * rng_mode 1 = reference LCG (sample 0 only), 0 = the product's Philox stream. */
void or_primary_hits(const OrScene* s, const float cam12[12], int W, int H, int spp_total, int sample,
                     uint64_t seed, int rng_mode, int32_t* tri, float* t, float* u, float* v);
/* rays: n x {ox,oy,oz,tmin,dx,dy,dz,tmax}; closest hit through scene_hit (gpu_render.cu:509-551). */
void or_trace(const OrScene* s, const float* rays, int n, int32_t* tri, float* t, float* u, float* v);
void or_occluded(const OrScene* s, const float* rays, int n, uint8_t* occluded);

/* render_kernel + ray_color (gpu_render.cu:973-1031, 715-936) for samples [s_begin, s_end) of a
 * frame of spp_total samples.  accum (W*H*3, row 0 = bottom) receives the SUM of clamped
 * samples (added to its current contents); rgb (may be NULL) the reference's tone-mapped
 * bytes for accum/spp_total, top row first.  rng_mode 1 = reference LCG (needs s_begin == 0
 * or the lcg_state array carried over), 0 = the product's Philox4x32-10 stream.
 * ray_counts[3] (may be NULL) += primary, bounce, shadow rays traced. */
void or_render(const OrScene* s, const float cam12[12], int W, int H, int spp_total, int max_depth,
               const float sun_dir[3], const float sun_radiance[3], int sun_enabled,
               uint64_t seed, int rng_mode, float gamma, int s_begin, int s_end,
               float* accum, unsigned char* rgb, uint32_t* lcg_state, uint64_t ray_counts[3]);
/* The kernel tail alone (gpu_render.cu:1003-1030). */
void or_tonemap(const float* accum, int W, int H, int spp_total, float gamma, unsigned char* rgb);

/* Host side: main.cpp:139-173, 334-357, 118-128; camera.h:91-116. */
int  or_read_pose_file(const char* path, double* poses9, int max_poses);
void or_pose_to_model_frame(const double pose9[9], float cam_in_model[3], float sun_dir_model[3], double* sep_m);
void or_camera_init(const float lookfrom[3], float vfov_deg, int W, int H, float cam12[12], float uvw9[9]);

/* Worker threads of the parallel loops (default: all online cores). */
void or_set_threads(int n);
int  or_get_threads(void);

/* Philox4x32-10 known-answer hook: writes the 4 outputs for (ctr, key). */
void or_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
