// render.cuh -- shared declarations of the wavefront path tracer (render.cu) and the C ABI
// implementation (api.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "dsrt_types.h"

namespace dsrt {

// Device counters (uint32) -- one cache line, reset per batch; statistics follow as uint64.
enum {
    C_NIN = 0,        // rays in the current queue
    C_NOUT = 1,       // rays pushed to the next queue
    C_NSHADOW0 = 2,   // shadow rays pushed by k_shade at even depth
    C_NSHADOW1 = 3,   // ... at odd depth (double buffered, see k_advance)
    C_FETCH = 4,      // dynamic work counter of the traversal kernels
    C_TAIL_DONE = 5,  // blocks of k_tail that have finished (reset by the last one)
    C_FETCH_LIST = 6, // dynamic work counter of k_primary_lists
    C_ZERO = 7,       // always 0 (an opaque zero for the traversal kernels, see trace_loop)
    C_COUNT = 8
};
enum { S_PRIMARY = 0, S_BOUNCE = 1, S_SHADOW = 2, S_PATHS = 3, S_CULLED = 4, S_COUNT = 5 };

// Everything a frame's kernels need, passed by value as a __grid_constant__ parameter so
// that every field is a constant-bank read (the reference dereferences a GPUScene* in
// global memory for each field, gpu_render.cu:981).
struct FrameConst {
    // scene
    const uint4* nodes;
    const float4* tri_records;
    const float4* shade_records;      // 3 words per triangle, or 1 when shade_compact (bvh_build.h)
    int shade_compact;
    const float4* uv_records;
    const float4* materials;          // DsrtMaterial as 3 x float4
    const DsrtTextureHeader* textures;
    const float* texture_pool;
    const float4* spheres;            // 2 x float4 per sphere: centre.xyz, radius | 1/radius, bits(material), 0, 0
    const int* sphere_lights;         // indices of the emissive spheres, in sphere order (gpu_render.cu:842-848)
    int num_nodes, num_textures, texture_pool_floats, num_materials;
    int shade_bins;                   // 1: k_shade collects hits per material class (more than one class in the table, and it fits)
    int num_spheres, num_sphere_lights;
    float scene_lo[3], scene_hi[3];   // scene bounds padded by 2^-16 of the scene size (root-box test)
    // camera (reference GPUCamera fields that the kernel reads)
    float cam_o[3], cam_llc[3], cam_h[3], cam_v[3];
    int W, H;
    int x0, y0, rw, rh;               // render rectangle
    const uint2* tile_entries;        // per 8x4-pixel tile: the node groups its camera rays start from (k_tile_entries),
                                      // y == 0: the tile cannot hit a triangle
    const uint32_t* active_tiles;     // the tiles whose camera rays are generated at all, compacted by k_tile_entries
    uint32_t* n_active;               // their number (one device word per frame)
    int tile_walk;                    // 1: k_tile_entries walks the BVH per tile; 0: every tile starts at the root
    uint2* tile_lists;                // per tile kTileListCap + 1 words {node, record}: the triangles its camera rays can hit, when
                                      // they are few ([0].x = how many, kTileListNone = too many: the rays traverse the BVH)
    uint32_t* list_tiles;             // the tiles whose camera rays take that list (k_primary_lists), compacted
    uint32_t* n_list;                 // their number (device word); nullptr: no lists this frame
    int cull_tiles;                   // 1: a tile with no entries is skipped outright (Philox stream, no spheres)
    int cull_pixels;                  // 1: the bounds below are in force (Philox stream, no spheres, scene box in front of the camera)
    int hit_x0, hit_y0, hit_x1, hit_y1;   // pixels outside [hit_x0, hit_x1] x [hit_y0, hit_y1] cannot see the triangles:
                                      // the screen-space bounds of the scene box, padded (api.cu: scene_screen_bounds)
    int tiles_x, slots_per_sample;    // 8x4-pixel tiles; slots = tiles*32
    float inv_tiles_x, inv_slots_per_sample;   // 1/x as floats: quotient estimates for path -> (sample, tile, pixel)
    // estimator
    int spp_total, max_depth, rng_mode;
    uint32_t seed_lo, seed_hi;
    int sun_enabled;
    float ldir[3];                    // normalise(-sun_dir), computed with the reference's ops
    float sun_rad[3];
    // batch
    int s0, spb;                      // first global sample index and samples in this batch
    // work memory
    float4* rq_o[2];                  // ray queue: origin.xyz, bits(path)
    float4* rq_d[2];                  // ray queue: dir.xyz, unused
    float4* hits;                     // t,u,v,bits(id) per queue entry; id = uploaded triangle index,
                                      // -1 = miss, -2 - i = sphere i
    uint32_t* pq;                     // first hits (what k_primary leaves for the depth-0 k_shade): path of entry i of
                                      // `hits`; the camera ray is regenerated from the path, not stored
    uint2* dbg_leaf;                  // parity hook only: (node index, record position) of first hit i; else nullptr
    float4* st_a;                     // per path: throughput.xyz, bits(rng state / draw index)
    float4* st_b;                     // per path: L.xyz (k_resolve applies the per-sample clamp), unused
    float4* sq_o;                     // shadow queue: origin.xyz, bits(path)
    float4* sq_l;                     // shadow queue: L if the sun is visible
    uint32_t* rng_pix;                // LCG replay: per-pixel generator state
    float* accum;                     // W*H*3 sum of clamped samples, row 0 = bottom
    uint32_t* counters;               // C_COUNT uint32
    unsigned long long* stats;        // S_COUNT uint64
};

struct LaunchDims {
    int primary_blocks = 0, trace_blocks = 0, shade_blocks = 0, tail_blocks = 0;
};
#ifndef DSRT_TILE_LIST_CAP
#define DSRT_TILE_LIST_CAP 64
#endif
constexpr int kTileListCap = DSRT_TILE_LIST_CAP;      // longest candidate list a tile may have
constexpr uint32_t kTileListNone = 0xFFFFFFFFu;
constexpr int kShadeClassTable = 1024;          // k_shade's shared-memory table of material classes
constexpr uint32_t kTailAll = 0xFFFFFFFFu;      // launch_tail threshold: finish whatever is left

// Host-callable launchers (all asynchronous on `s`).
void launch_dims(LaunchDims* d, int device);
void launch_batch_begin(const FrameConst& fc, cudaStream_t s);
// Fills `table` (kTileEntries uint2 per 8x4-pixel tile of the render rectangle) with the nodes the rays of each tile
// start from, and fc.active_tiles / fc.n_active with the tiles that need camera rays at all (n_active zeroed first).
void launch_tile_entries(const FrameConst& fc, uint2* table, uint32_t* active_tiles, cudaStream_t s);
// track_leaf: the parity-hook variant that also fills fc.dbg_leaf.
void launch_primary(const FrameConst& fc, const LaunchDims& d, bool track_leaf, cudaStream_t s);
// The camera rays of the tiles with a candidate list (no-op when fc.n_list is null or the list is empty).
void launch_primary_lists(const FrameConst& fc, const LaunchDims& d, bool track_leaf, cudaStream_t s);
void launch_trace(const FrameConst& fc, int cur, int shadow_slot, const LaunchDims& d, cudaStream_t s);
void launch_shade(const FrameConst& fc, int cur, int depth, const LaunchDims& d, cudaStream_t s);
void launch_advance(const FrameConst& fc, int depth, cudaStream_t s);
// All remaining bounces of the rays in queue `cur` (traced, not yet shaded at depth0) in one launch --
// a no-op unless at most `threshold` rays are left (kTailAll: never a no-op).
void launch_tail(const FrameConst& fc, int cur, int depth0, uint32_t threshold, const LaunchDims& d, cudaStream_t s);
void launch_resolve(const FrameConst& fc, const LaunchDims& d, cudaStream_t s);
void launch_tonemap(const float* accum, unsigned char* rgb, int W, int H, float inv_spp, float inv_gamma, cudaStream_t s);
void launch_lcg_seed(uint32_t* rng_pix, int W, int H, uint32_t seed32, cudaStream_t s);

// Generic traversal over an array of rays {ox,oy,oz,tmin | dx,dy,dz,tmax}.
struct SphereSet { const float4* records; int count; };
void launch_trace_closest(const uint4* nodes, const float4* tri_records, int num_nodes, SphereSet spheres,
                          const float4* rays, int n, float4* hits, uint32_t* fetch_counter,
                          int blocks, cudaStream_t s);
void launch_trace_occluded(const uint4* nodes, const float4* tri_records, int num_nodes, SphereSet spheres,
                           const float4* rays, int n, uint8_t* occluded, uint32_t* fetch_counter,
                           int blocks, cudaStream_t s);
void launch_camera_rays(const FrameConst& fc, float jx, float jy, float4* rays, cudaStream_t s);

}  // namespace dsrt
