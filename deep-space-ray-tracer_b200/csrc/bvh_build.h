// bvh_build.h -- host interface of the GPU BVH8 builder (bvh_build.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "dsrt_types.h"

namespace dsrt {

struct Bvh8 {
    uint4* nodes = nullptr;          // num_nodes x 6 words (96 B each), breadth-first
    float4* tri_records = nullptr;   // num_tris x 3 words (48 B each), leaf order
    float4* shade_records = nullptr; // by uploaded triangle index: 3 words per triangle (n0 + material, n1 + texture, n2), or ONE
                                     // word (n + material | (texture + 1) << 16) when shade_compact
    bool shade_compact = false;      // every triangle has n0 == n1 == n2 (the reference's loader only makes flat normals,
                                     // triangle.h:70-73) and ids below 65535: 16-byte records instead of 48
    float4* uv_records = nullptr;    // num_tris x 2 words, only when textures are present
    int num_nodes = 0;
    int num_tris = 0;
    int num_levels = 0;
    int build_rounds = 0;            // PLOC merge rounds (0 for the radix tree)
    float build_ms = 0.f;
    float sah_cost = 0.f;
    float scene_lo[3] = {0, 0, 0};
    float scene_hi[3] = {0, 0, 0};
};

// Builds the wide BVH for `n` triangles already resident on the device in the reference's
// 116-byte layout.  Returns 0, or -1 with a message in `err`.
enum { DSRT_BUILDER_PLOC = 0, DSRT_BUILDER_LBVH = 1 };
int build_bvh8(cudaStream_t stream, const DsrtTriangle* d_tris, int n, bool want_uv, int builder,
               Bvh8* out, char* err, size_t err_len);
void free_bvh8(Bvh8& b);

}  // namespace dsrt
