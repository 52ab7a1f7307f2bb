// ref_host_flatten.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the reference's own OBJ -> GPUScene flattening on the CPU so that golden arrays for the
// OBJ/MTL/texture loader can be generated in a container without a GPU: the reference's
// unmodified inc/triangle_mesh.h (OBJ + MTL parser, MTL -> material mapping) and
// src/gpu_scene_builder.cpp (build_gpu_scene: material table, texture registry through
// stb_image, triangle records) are compiled where they lie, and the five CUDA runtime calls
// build_gpu_scene makes are satisfied by the host-memory shim at the bottom of this file, so
// the "device" pointers of the returned GPUScene are plain host pointers.
// Built into oracle/_ref/libref_host.so by oracle/Makefile.
#include <cuda_runtime_api.h>

#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>

#include "gpu_scene_builder.h"
#include "hittable_list.h"
#include "triangle_mesh.h"
#include "material.h"
#include "camera.h"

extern "C" {

struct RefFlat { GPUScene scene; };

// triangle_mesh(obj_path, fallback = lambertian(0.73) as main.cpp:241-246, scale) -> build_gpu_scene
void* refflat_load_obj(const char* obj_path, double scale) {
    auto fallbackM = std::make_shared<lambertian>(vec3(0.73, 0.73, 0.73));
    auto mesh = std::make_shared<triangle_mesh>(std::string(obj_path), fallbackM, scale);
    hittable_list world;
    world.add(mesh);
    camera cam;
    cam.image_width = 64; cam.image_height = 36;
    cam.lookfrom = point3(0, 0, 100); cam.lookat = point3(0, 0, 0); cam.vup = vec3(0, 1, 0);
    cam.focus_dist = 100.0f;
    cam.initialize();
    RefFlat* R = new RefFlat();
    R->scene = build_gpu_scene(world, cam, vec3(0, 0, 1));
    return R;
}

// n[6] = {triangles, materials, spheres, textures, texture_pool_floats, bvh_nodes}
void refflat_counts(void* h, int n[6]) {
    const GPUScene& S = ((RefFlat*)h)->scene;
    n[0] = S.num_triangles; n[1] = S.num_materials; n[2] = S.num_spheres; n[3] = S.num_textures;
    n[4] = S.texture_pool_floats; n[5] = S.num_bvh_nodes;
}

void refflat_copy(void* h, void* tris, void* mats, void* tex_headers, float* pool) {
    const GPUScene& S = ((RefFlat*)h)->scene;
    if (tris && S.num_triangles) memcpy(tris, S.triangles, sizeof(GPUTriangle) * (size_t)S.num_triangles);
    if (mats && S.num_materials) memcpy(mats, S.materials, sizeof(GPUMaterial) * (size_t)S.num_materials);
    if (tex_headers && S.num_textures) memcpy(tex_headers, S.textures, sizeof(GPUTextureHeader) * (size_t)S.num_textures);
    if (pool && S.texture_pool_floats) memcpy(pool, S.texture_pool, sizeof(float) * (size_t)S.texture_pool_floats);
}

void refflat_free(void* h) {
    RefFlat* R = (RefFlat*)h;
    if (!R) return;
    free_gpu_scene(R->scene);
    delete R;
}

// ---- host-memory shim for the CUDA runtime calls of gpu_scene_builder.cpp ----------------------
cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "host shim"; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }

}  // extern "C"
