// ref_gpu_harness.cu -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
//
// Wraps the UNMODIFIED reference renderer so tests and bench can run it on the GPU box next to
// the B200 path.  This translation unit #includes the reference's own
// /root/reference/src/gpu_render.cu (so render_kernel, scene_hit, make_camera_ray_jittered ...
// are the reference's code, compiled with the reference's flags: -rdc=true --extended-lambda,
// no fast-math) and links the reference's own src/gpu_scene_builder.cpp (build_gpu_scene, the
// CPU median-split BVH).  Nothing of the reference is copied into this repository: the
// sources are compiled where they lie and only the resulting oracle/_ref/libref_gpu.so
// (git-ignored) travels to the GPU box.  See oracle/Makefile.
//
// Added here: a C entry layer, two probe kernels (first hit of a camera ray / of an arbitrary
// ray through the reference's scene_hit) and CUDA-event timing around render_kernel.
#include "gpu_render.cu"             // found through -I/root/reference/src

#include "gpu_scene_builder.h"
#include "hittable_list.h"
#include "triangle.h"
#include "triangle_mesh.h"
#include "sphere.h"
#include "material.h"
#include "camera.h"

#include <cstring>
#include <memory>
#include <vector>

struct RefMaterial {   // layout of the B200 path's DsrtMaterial / the reference's GPUMaterial
    int type, albedo_tex, pad0, pad1;
    float albedo[3], emissive[3];
    float fuzz, ref_idx;
};

struct RefScene {
    GPUScene scene;
    GPUScene* d_scene = nullptr;
    int W = 0, H = 0;
    int normal_mismatches = 0;
    float build_ms = 0.f;
};

__global__ void ref_first_hit_kernel(const GPUScene* dscene, int W, int H, float jx, float jy,
                                     int* tri, float* t, float* u, float* v) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const GPUScene& scene = *dscene;
    RayD ray = make_camera_ray_jittered(scene, x, y, W, H, jx, jy);
    HitRecord rec;
    bool hit = scene_hit(scene, ray, 0.001f, 1e9f, rec);
    int i = x + y * W;
    tri[i] = hit ? rec.tri_index : -1;
    t[i] = hit ? rec.t : 1e9f;
    u[i] = hit ? rec.u : 0.f;
    v[i] = hit ? rec.v : 0.f;
}

__global__ void ref_trace_kernel(const GPUScene* dscene, const float* rays, int n,
                                 int* tri, float* t, float* u, float* v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GPUScene& scene = *dscene;
    const float* r = rays + (size_t)i * 8;
    RayD ray(make_float3(r[0], r[1], r[2]), make_float3(r[4], r[5], r[6]));
    HitRecord rec;
    bool hit = scene_hit(scene, ray, r[3], r[7], rec);
    tri[i] = hit ? rec.tri_index : -1;
    t[i] = hit ? rec.t : r[7];
    u[i] = hit ? rec.u : 0.f;
    v[i] = hit ? rec.v : 0.f;
}

extern "C" {

static camera make_ref_camera(const float lookfrom[3], float vfov, int W, int H, int spp, int max_depth) {
    camera cam;
    cam.image_width = W; cam.image_height = H;
    cam.samples_per_pixel = spp; cam.max_depth = max_depth;
    cam.vfov = vfov; cam.aperture = 0.0f;
    cam.lookfrom = vec3(lookfrom[0], lookfrom[1], lookfrom[2]);
    cam.lookat = vec3(0, 0, 0);
    cam.vup = vec3(0, 1, 0);
    cam.focus_dist = (cam.lookfrom - cam.lookat).length();
    cam.initialize();
    return cam;
}

static RefScene* finish_ref_scene(RefScene* R, const hittable_list& world, const camera& cam, const float sun_dir[3], int W, int H) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    R->scene = build_gpu_scene(world, cam, vec3(sun_dir[0], sun_dir[1], sun_dir[2]));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&R->build_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    R->W = W; R->H = H;
    if (cudaMalloc(&R->d_scene, sizeof(GPUScene)) != cudaSuccess ||
        cudaMemcpy(R->d_scene, &R->scene, sizeof(GPUScene), cudaMemcpyHostToDevice) != cudaSuccess) {
        delete R;
        return nullptr;
    }
    return R;
}

// tris: n x 29 floats in the reference's GPUTriangle layout (only v0,v1,v2 and material_id are
// consumed; the reference recomputes its flat normals, which are compared with the ones given).
// spheres: nspheres x {cx,cy,cz,radius,bits(material_id),pad} become reference `sphere` objects.
void* ref_scene_create_ex(const float* tris, int n, const RefMaterial* mats, int nmats,
                          const float* spheres, int nspheres,
                          const float lookfrom[3], float vfov, int W, int H, int spp, int max_depth,
                          const float sun_dir[3]) {
    std::vector<std::shared_ptr<material>> M;
    for (int i = 0; i < nmats; ++i) {
        const RefMaterial& m = mats[i];
        vec3 a(m.albedo[0], m.albedo[1], m.albedo[2]);
        if (m.type == MAT_METAL) M.push_back(std::make_shared<metal>(a, (double)m.fuzz));
        else if (m.type == MAT_DIELECTRIC) M.push_back(std::make_shared<dielectric>((double)m.ref_idx));
        else if (m.type == MAT_DIFFUSE_LIGHT) M.push_back(std::make_shared<diffuse_light>(vec3(m.emissive[0], m.emissive[1], m.emissive[2])));
        else M.push_back(std::make_shared<lambertian>(a));
    }
    RefScene* R = new RefScene();
    hittable_list world;
    for (int i = 0; i < n; ++i) {
        const float* p = tris + (size_t)i * 29;
        int mid;
        std::memcpy(&mid, p + 27, 4);
        auto t = std::make_shared<triangle>(vec3(p[0], p[1], p[2]), vec3(p[3], p[4], p[5]), vec3(p[6], p[7], p[8]),
                                            M[(size_t)mid]);
        float nn[3] = {t->n0.x(), t->n0.y(), t->n0.z()};
        if (std::memcmp(nn, p + 9, 12) != 0) R->normal_mismatches++;
        world.add(t);
    }
    for (int i = 0; i < nspheres; ++i) {
        const float* p = spheres + (size_t)i * 6;
        int mid;
        std::memcpy(&mid, p + 4, 4);
        world.add(std::make_shared<sphere>(vec3(p[0], p[1], p[2]), (double)p[3], M[(size_t)mid]));
    }
    camera cam = make_ref_camera(lookfrom, vfov, W, H, spp, max_depth);
    return finish_ref_scene(R, world, cam, sun_dir, W, H);
}

void* ref_scene_create(const float* tris, int n, const RefMaterial* mats, int nmats,
                       const float lookfrom[3], float vfov, int W, int H, int spp, int max_depth,
                       const float sun_dir[3]) {
    return ref_scene_create_ex(tris, n, mats, nmats, nullptr, 0, lookfrom, vfov, W, H, spp, max_depth, sun_dir);
}

// The reference's own OBJ path: triangle_mesh (inc/triangle_mesh.h) with main.cpp's fallback
// material, flattened by build_gpu_scene (materials, per-triangle textures through stb_image).
void* ref_scene_create_obj(const char* obj_path, double scale,
                           const float lookfrom[3], float vfov, int W, int H, int spp, int max_depth,
                           const float sun_dir[3]) {
    RefScene* R = new RefScene();
    auto fallbackM = std::make_shared<lambertian>(vec3(0.73, 0.73, 0.73));      // main.cpp:241
    auto mesh = std::make_shared<triangle_mesh>(std::string(obj_path), fallbackM, scale);
    hittable_list world;
    world.add(mesh);
    camera cam = make_ref_camera(lookfrom, vfov, W, H, spp, max_depth);
    return finish_ref_scene(R, world, cam, sun_dir, W, H);
}

// Copies the flattened scene arrays back from the device (what build_gpu_scene uploaded):
// any pointer may be NULL; counts are returned through n[6] = {triangles, materials, spheres,
// textures, texture_pool_floats, bvh_nodes}.
int ref_scene_arrays(void* h, int n[6], void* tris, void* mats, void* spheres, void* tex_headers, float* pool) {
    RefScene* R = (RefScene*)h;
    if (!R) return -1;
    const GPUScene& S = R->scene;
    n[0] = S.num_triangles; n[1] = S.num_materials; n[2] = S.num_spheres; n[3] = S.num_textures;
    n[4] = S.texture_pool_floats; n[5] = S.num_bvh_nodes;
    if (tris && S.num_triangles) cudaMemcpy(tris, S.triangles, sizeof(GPUTriangle) * (size_t)S.num_triangles, cudaMemcpyDeviceToHost);
    if (mats && S.num_materials) cudaMemcpy(mats, S.materials, sizeof(GPUMaterial) * (size_t)S.num_materials, cudaMemcpyDeviceToHost);
    if (spheres && S.num_spheres) cudaMemcpy(spheres, S.spheres, sizeof(GPUSphere) * (size_t)S.num_spheres, cudaMemcpyDeviceToHost);
    if (tex_headers && S.num_textures) cudaMemcpy(tex_headers, S.textures, sizeof(GPUTextureHeader) * (size_t)S.num_textures, cudaMemcpyDeviceToHost);
    if (pool && S.texture_pool_floats) cudaMemcpy(pool, S.texture_pool, sizeof(float) * (size_t)S.texture_pool_floats, cudaMemcpyDeviceToHost);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// The reference's GPUScene itself (a host struct of device pointers, 384 bytes): what its main() hands to
// gpu_render_scene -- for tests of the legacy boundary (include/dsrt_legacy.h).
const void* ref_scene_legacy(void* h) {
    RefScene* R = (RefScene*)h;
    return R ? (const void*)&R->scene : nullptr;
}

int ref_scene_info(void* h, int* num_bvh_nodes, int* normal_mismatches, float* build_ms, float camera12[12]) {
    RefScene* R = (RefScene*)h;
    if (!R) return -1;
    if (num_bvh_nodes) *num_bvh_nodes = R->scene.num_bvh_nodes;
    if (normal_mismatches) *normal_mismatches = R->normal_mismatches;
    if (build_ms) *build_ms = R->build_ms;
    if (camera12) {
        const GPUCamera& c = R->scene.camera;
        const vec3* v[4] = {&c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical};
        for (int k = 0; k < 4; ++k) { camera12[3 * k] = v[k]->x(); camera12[3 * k + 1] = v[k]->y(); camera12[3 * k + 2] = v[k]->z(); }
    }
    return 0;
}

// Override the seed (reference fixes 1337) -- used only to measure the reference's own
// seed-to-seed noise floor for the image tolerance.
int ref_scene_set_seed(void* h, unsigned long long seed) {
    RefScene* R = (RefScene*)h;
    if (!R) return -1;
    R->scene.seed = seed;
    return cudaMemcpy(R->d_scene, &R->scene, sizeof(GPUScene), cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -2;
}

// Runs the reference render_kernel exactly as gpu_render_scene launches it (8x8 blocks,
// gpu_render.cu:1074-1078) and returns the RGB8 framebuffer and the kernel's device time.
int ref_render(void* h, unsigned char* rgb, float* kernel_ms) {
    RefScene* R = (RefScene*)h;
    if (!R) return -1;
    const int W = R->W, H = R->H;
    const size_t bytes = (size_t)W * H * 3;
    unsigned char* d_fb = nullptr;
    if (cudaMalloc(&d_fb, bytes) != cudaSuccess) return -2;
    const float gamma = (R->scene.params.gamma > 0.0f) ? R->scene.params.gamma : 1.0f;
    const float exposure = (R->scene.params.exposure > 0.0f) ? R->scene.params.exposure : 2.0f;
    dim3 block(8, 8), grid((W + 7) / 8, (H + 7) / 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    render_kernel<<<grid, block>>>(R->d_scene, d_fb, W, H, 1.0f / gamma, exposure);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (kernel_ms) *kernel_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (err == cudaSuccess) err = cudaMemcpy(rgb, d_fb, bytes, cudaMemcpyDeviceToHost);
    cudaFree(d_fb);
    return err == cudaSuccess ? 0 : -3;
}

int ref_first_hits(void* h, float jx, float jy, int* tri, float* t, float* u, float* v) {
    RefScene* R = (RefScene*)h;
    if (!R) return -1;
    const int W = R->W, H = R->H, n = W * H;
    int* d_tri; float *d_t, *d_u, *d_v;
    cudaMalloc(&d_tri, 4 * (size_t)n); cudaMalloc(&d_t, 4 * (size_t)n); cudaMalloc(&d_u, 4 * (size_t)n); cudaMalloc(&d_v, 4 * (size_t)n);
    dim3 block(8, 8), grid((W + 7) / 8, (H + 7) / 8);
    ref_first_hit_kernel<<<grid, block>>>(R->d_scene, W, H, jx, jy, d_tri, d_t, d_u, d_v);
    cudaError_t err = cudaDeviceSynchronize();
    if (err == cudaSuccess) {
        cudaMemcpy(tri, d_tri, 4 * (size_t)n, cudaMemcpyDeviceToHost);
        cudaMemcpy(t, d_t, 4 * (size_t)n, cudaMemcpyDeviceToHost);
        cudaMemcpy(u, d_u, 4 * (size_t)n, cudaMemcpyDeviceToHost);
        cudaMemcpy(v, d_v, 4 * (size_t)n, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_tri); cudaFree(d_t); cudaFree(d_u); cudaFree(d_v);
    return err == cudaSuccess ? 0 : -3;
}

int ref_trace(void* h, const float* rays, int n, int* tri, float* t, float* u, float* v) {
    RefScene* R = (RefScene*)h;
    if (!R) return -1;
    float* d_rays; int* d_tri; float *d_t, *d_u, *d_v;
    cudaMalloc(&d_rays, 32 * (size_t)n);
    cudaMalloc(&d_tri, 4 * (size_t)n); cudaMalloc(&d_t, 4 * (size_t)n); cudaMalloc(&d_u, 4 * (size_t)n); cudaMalloc(&d_v, 4 * (size_t)n);
    cudaMemcpy(d_rays, rays, 32 * (size_t)n, cudaMemcpyHostToDevice);
    ref_trace_kernel<<<(n + 63) / 64, 64>>>(R->d_scene, d_rays, n, d_tri, d_t, d_u, d_v);
    cudaError_t err = cudaDeviceSynchronize();
    if (err == cudaSuccess) {
        cudaMemcpy(tri, d_tri, 4 * (size_t)n, cudaMemcpyDeviceToHost);
        cudaMemcpy(t, d_t, 4 * (size_t)n, cudaMemcpyDeviceToHost);
        cudaMemcpy(u, d_u, 4 * (size_t)n, cudaMemcpyDeviceToHost);
        cudaMemcpy(v, d_v, 4 * (size_t)n, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_rays); cudaFree(d_tri); cudaFree(d_t); cudaFree(d_u); cudaFree(d_v);
    return err == cudaSuccess ? 0 : -3;
}

void ref_scene_destroy(void* h) {
    RefScene* R = (RefScene*)h;
    if (!R) return;
    free_gpu_scene(R->scene);
    cudaFree(R->d_scene);
    delete R;
}

}  // extern "C"
