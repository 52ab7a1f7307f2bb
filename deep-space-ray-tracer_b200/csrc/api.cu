// api.cu -- implementation of the C ABI declared in include/dsrt.h.
//
// This is the only layer a host program talks to.  It owns the context (one GPU, one
// stream), the scene (wide BVH + records + materials), the frame work memory (ray queues,
// path state, accumulation buffer) and orchestrates the wavefront launches of render.cu.
// There is deliberately no CPU code path here: every entry that computes needs the device.
#include "dsrt.h"
#include "dsrt_legacy.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "bvh8.cuh"
#include "bvh_build.h"
#include "render.cuh"

using namespace dsrt;

namespace {

char g_create_error[512] = "";

constexpr int kMaxEvents = 8192;

}  // namespace

// Everything one frame in flight needs: its own stream, image buffers, work memory, counters and timers.
// Slot 0 serves the synchronous calls; dsrt_render_frame_async can use any slot, and frames enqueued on
// different slots overlap on the GPU (the latency-bound end of one frame under the camera rays of the next).
struct FrameSlot {
    bool created = false;
    cudaStream_t stream = nullptr;
    int W = 0, H = 0;
    float* accum = nullptr;
    uint32_t* rng_pix = nullptr;
    unsigned char* d_rgb = nullptr;
    uint2* tile_table = nullptr;       // per-tile traversal entry groups of the current frame (render.cu: k_tile_entries)
    uint32_t* active_tiles = nullptr;  // tiles that need camera rays, compacted
    uint2* tile_lists = nullptr;       // per-tile candidate triangles (render.cu: tile_candidates)
    uint32_t* list_tiles = nullptr;    // tiles whose camera rays take their candidate list, compacted
    size_t tile_table_tiles = 0;
    int lcg_next_sample = -1;

    size_t cap_paths = 0;
    float4 *rq_o[2] = {nullptr, nullptr}, *rq_d[2] = {nullptr, nullptr};
    float4 *hits = nullptr, *st_a = nullptr, *st_b = nullptr, *sq_o = nullptr, *sq_l = nullptr;
    uint32_t* pq = nullptr;
    uint2* dbg_leaf = nullptr;         // parity hook only
    size_t cap_leaf = 0;
    uint32_t* counters = nullptr;      // C_COUNT words, then the active-tile count of the frame
    unsigned long long* stats = nullptr;
    unsigned long long* h_stats = nullptr;     // pinned: S_COUNT statistics of the frame + 8 counter words (hooks)

    cudaEvent_t ev_frame[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_kind;          // what the interval [k, k+1) measured: 0 k_trace, 1 k_shade, 2 k_primary, -1 other
    size_t nev = 0;
    bool pending = false;              // a frame was enqueued and not yet waited for
    uint32_t launches = 0, batches = 0;
    DsrtFrameStats fstats{};
};

constexpr int kMaxSlots = 4;

struct dsrt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;     // = slots[0].stream
    char err[512] = "";

    // scene
    bool have_scene = false;
    Bvh8 bvh;
    float4* d_mats = nullptr;
    int nmats = 0;
    float mean_tri_extent = 0.f;       // mean longest bounding-box side of the triangles
    int mat_classes = 1;               // how many of {lambertian/emitter, metal, dielectric} the material table holds
    DsrtTextureHeader* d_tex = nullptr;
    float* d_pool = nullptr;
    int ntex = 0, pool_floats = 0;
    float4* d_spheres = nullptr;       // 2 x float4 per sphere (render.cuh)
    int* d_sphere_lights = nullptr;
    int nspheres = 0, nsphere_lights = 0;
    DsrtSceneStats sstats{};

    FrameSlot slots[kMaxSlots];

    LaunchDims dims;
    int stage_timing = 0;
    int spb_override = 0;
    int use_tile_entries = 1;
    int use_tile_lists = 1;
    int builder = DSRT_BUILDER_PLOC;   // binary-tree stage of the BVH build (bvh_build.h)
    int work_gb = 40;                  // work-memory budget of a batch, per slot, in GB (DSRT_WORK_GB)
    int tail_threshold = 131072;       // rays left in a batch at which k_tail takes over (0 = never)
    int wave_depth = 16;               // bounces run as wavefront launches; deeper ones are finished by one closing k_tail

    // scratch for the trace hooks
    float4* t_rays = nullptr;
    float4* t_hits = nullptr;
    uint8_t* t_occ = nullptr;
    size_t t_cap = 0;
};

namespace {

int fail(dsrt_ctx* c, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c ? c->err : g_create_error, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return fail(ctx, DSRT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

void free_scene(dsrt_ctx* c) {
    free_bvh8(c->bvh);
    cudaFree(c->d_mats); c->d_mats = nullptr;
    cudaFree(c->d_tex); c->d_tex = nullptr;
    cudaFree(c->d_pool); c->d_pool = nullptr;
    cudaFree(c->d_spheres); c->d_spheres = nullptr;
    cudaFree(c->d_sphere_lights); c->d_sphere_lights = nullptr;
    c->nmats = c->ntex = c->pool_floats = c->nspheres = c->nsphere_lights = 0;
    c->mat_classes = 1;
    c->mean_tri_extent = 0.f;
    c->have_scene = false;
}

void free_work(FrameSlot& f) {
    for (int k = 0; k < 2; ++k) { cudaFree(f.rq_o[k]); cudaFree(f.rq_d[k]); f.rq_o[k] = f.rq_d[k] = nullptr; }
    cudaFree(f.hits); cudaFree(f.st_a); cudaFree(f.st_b); cudaFree(f.sq_o); cudaFree(f.sq_l); cudaFree(f.pq);
    f.hits = f.st_a = f.st_b = f.sq_o = f.sq_l = nullptr;
    f.pq = nullptr;
    f.cap_paths = 0;
}

void free_slot(FrameSlot& f) {
    if (!f.created) return;
    if (f.stream) cudaStreamSynchronize(f.stream);
    free_work(f);
    cudaFree(f.accum); cudaFree(f.rng_pix); cudaFree(f.d_rgb); cudaFree(f.tile_table); cudaFree(f.active_tiles);
    cudaFree(f.tile_lists); cudaFree(f.list_tiles);
    cudaFree(f.dbg_leaf);
    cudaFree(f.counters); cudaFree(f.stats);
    if (f.h_stats) cudaFreeHost(f.h_stats);
    for (cudaEvent_t e : f.ev_pool) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k) if (f.ev_frame[k]) cudaEventDestroy(f.ev_frame[k]);
    if (f.stream) cudaStreamDestroy(f.stream);
    f = FrameSlot();
}

int slot_init(dsrt_ctx* ctx, int k) {
    FrameSlot& f = ctx->slots[k];
    if (f.created) return DSRT_OK;
    f.created = true;
    CK(cudaStreamCreateWithFlags(&f.stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&f.counters, sizeof(uint32_t) * (C_COUNT + 8)));
    CK(cudaMemset(f.counters, 0, sizeof(uint32_t) * (C_COUNT + 8)));
    CK(cudaMalloc(&f.stats, sizeof(unsigned long long) * S_COUNT));
    CK(cudaMallocHost(&f.h_stats, sizeof(unsigned long long) * (S_COUNT + 8)));
    CK(cudaEventCreate(&f.ev_frame[0]));
    CK(cudaEventCreate(&f.ev_frame[1]));
    return DSRT_OK;
}

constexpr long long kWorkBytesPerPath = 9 * 16 + 4;      // what ensure_work allocates per path

int ensure_work(dsrt_ctx* ctx, FrameSlot& f, size_t paths) {
    if (paths <= f.cap_paths) return DSRT_OK;
    free_work(f);
    const size_t b = paths * sizeof(float4);
    for (int k = 0; k < 2; ++k) { CK(cudaMalloc(&f.rq_o[k], b)); CK(cudaMalloc(&f.rq_d[k], b)); }
    CK(cudaMalloc(&f.hits, b));
    CK(cudaMalloc(&f.st_a, b));
    CK(cudaMalloc(&f.st_b, b));
    CK(cudaMalloc(&f.sq_o, b));
    CK(cudaMalloc(&f.sq_l, b));
    CK(cudaMalloc(&f.pq, paths * sizeof(uint32_t)));
    f.cap_paths = paths;
    return DSRT_OK;
}

int ensure_image(dsrt_ctx* ctx, FrameSlot& f, int W, int H, bool* fresh) {
    *fresh = false;
    if (f.accum && f.W == W && f.H == H) return DSRT_OK;
    cudaFree(f.accum); cudaFree(f.rng_pix); cudaFree(f.d_rgb);
    f.accum = nullptr; f.rng_pix = nullptr; f.d_rgb = nullptr;
    const size_t px = (size_t)W * H;
    CK(cudaMalloc(&f.accum, px * 3 * sizeof(float)));
    CK(cudaMalloc(&f.rng_pix, px * sizeof(uint32_t)));
    CK(cudaMalloc(&f.d_rgb, px * 3));
    f.W = W; f.H = H;
    f.lcg_next_sample = -1;
    *fresh = true;
    return DSRT_OK;
}

int ensure_trace_scratch(dsrt_ctx* ctx, size_t n) {
    if (n <= ctx->t_cap) return DSRT_OK;
    cudaFree(ctx->t_rays); cudaFree(ctx->t_hits); cudaFree(ctx->t_occ);
    ctx->t_rays = nullptr; ctx->t_hits = nullptr; ctx->t_occ = nullptr; ctx->t_cap = 0;
    CK(cudaMalloc(&ctx->t_rays, n * 2 * sizeof(float4)));
    CK(cudaMalloc(&ctx->t_hits, n * sizeof(float4)));
    CK(cudaMalloc(&ctx->t_occ, n));
    ctx->t_cap = n;
    return DSRT_OK;
}

// normalise(-sun_dir) with the device's operation order (exact_math.cuh: vdot, vnorm_ref).
void host_ldir(const float s[3], float out[3]) {
    const float x = -s[0], y = -s[1], z = -s[2];
    const float len2 = std::fmaf(z, z, std::fmaf(x, x, y * y));
    const float L = std::sqrt(len2);
    if (L <= 0.0f) { out[0] = out[1] = out[2] = 0.0f; return; }
    const float inv = 1.0f / L;
    out[0] = x * inv; out[1] = y * inv; out[2] = z * inv;
}

cudaEvent_t pool_event(FrameSlot& f, size_t i) {
    while (f.ev_pool.size() <= i) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        f.ev_pool.push_back(e);
    }
    return f.ev_pool[i];
}

int sync_all_slots(dsrt_ctx* ctx) {
    for (int k = 0; k < kMaxSlots; ++k)
        if (ctx->slots[k].created) CK(cudaStreamSynchronize(ctx->slots[k].stream));
    return DSRT_OK;
}

}  // namespace

extern "C" const char* dsrt_version(void) { return "dsrt-b200 0.1 sm_100a"; }

extern "C" const char* dsrt_last_error(const dsrt_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

extern "C" int dsrt_create(dsrt_ctx** out, int device) {
    dsrt_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, DSRT_EINVAL, "dsrt_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(nullptr, DSRT_ECUDA, "no usable CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= count) return fail(nullptr, DSRT_EINVAL, "device %d out of range (0..%d)", device, count - 1);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, DSRT_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    ctx = new (std::nothrow) dsrt_ctx();
    if (!ctx) return fail(nullptr, DSRT_ENOMEM, "out of host memory");
    ctx->device = device;
    if (slot_init(ctx, 0) != DSRT_OK) {
        fail(nullptr, DSRT_ECUDA, "context set-up failed: %s", ctx->err);
        dsrt_destroy(ctx);
        return DSRT_ECUDA;
    }
    ctx->stream = ctx->slots[0].stream;
    launch_dims(&ctx->dims, device);
    // tuning / A-B switches for experiments; none is needed for normal use
    if (const char* e = getenv("DSRT_TAIL_THRESHOLD")) ctx->tail_threshold = atoi(e);
    if (const char* e = getenv("DSRT_TILE_ENTRIES")) ctx->use_tile_entries = atoi(e) != 0;
    if (const char* e = getenv("DSRT_TILE_LISTS")) ctx->use_tile_lists = std::max(0, std::min(2, atoi(e)));
    if (const char* e = getenv("DSRT_WORK_GB")) ctx->work_gb = std::max(1, std::min(120, atoi(e)));
    if (const char* e = getenv("DSRT_WAVE_DEPTH")) ctx->wave_depth = std::max(2, atoi(e));
    if (const char* e = getenv("DSRT_BVH_BUILDER")) ctx->builder = strcmp(e, "lbvh") == 0 ? DSRT_BUILDER_LBVH : DSRT_BUILDER_PLOC;
    *out = ctx;
    return DSRT_OK;
}

extern "C" void dsrt_destroy(dsrt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int k = 0; k < kMaxSlots; ++k)
        if (ctx->slots[k].created && ctx->slots[k].stream) cudaStreamSynchronize(ctx->slots[k].stream);
    free_scene(ctx);
    for (int k = 0; k < kMaxSlots; ++k) free_slot(ctx->slots[k]);
    cudaFree(ctx->t_rays); cudaFree(ctx->t_hits); cudaFree(ctx->t_occ);
    delete ctx;
}

extern "C" void* dsrt_stream(dsrt_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

namespace {

// Sum over the triangles of the longest side of their bounding boxes (one atomic per warp): the scene's typical
// triangle size, against which a frame decides whether per-tile candidate lists can pay (enqueue_frame).
__global__ void k_sum_extent(const DsrtTriangle* __restrict__ tris, int n, float* sum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float e = 0.f;
    if (i < n) {
        const DsrtTriangle& t = tris[i];
        const float ex = fmaxf(t.v0.x, fmaxf(t.v1.x, t.v2.x)) - fminf(t.v0.x, fminf(t.v1.x, t.v2.x));
        const float ey = fmaxf(t.v0.y, fmaxf(t.v1.y, t.v2.y)) - fminf(t.v0.y, fminf(t.v1.y, t.v2.y));
        const float ez = fmaxf(t.v0.z, fmaxf(t.v1.z, t.v2.z)) - fminf(t.v0.z, fminf(t.v1.z, t.v2.z));
        e = fmaxf(ex, fmaxf(ey, ez));
        if (!(e == e) || e > 3.0e38f) e = 0.f;
    }
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0 && e > 0.f) atomicAdd(sum, e);
}

__global__ void k_check_material_ids(const DsrtTriangle* __restrict__ tris, int n, int nmats, uint32_t* bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (tris[i].material_id < 0 || tris[i].material_id >= nmats)) atomicAdd(bad, 1u);
}

// The reference's bbox_hit rejects a box when t_max <= t_min after any axis (gpu_render.cu:312), so a node whose box
// has zero thickness along an axis is never entered: the triangles of such a LEAF of the reference's own BVH
// (axis-aligned coplanar panels that its median split happened to group together) are invisible in the reference,
// for every ray.  Given that BVH (40-byte GPUBVHNode: min[3], max[3], left, right, tri_offset, tri_count; leaf when
// tri_count > 0), this collapses those triangles to a point in OUR copy of the triangle array, so that they can never
// be hit here either.  One thread per reference node.
__global__ void k_mask_flat_leaves(const float* __restrict__ ref_nodes, int num_nodes, const int* __restrict__ tri_indices,
                                   DsrtTriangle* tris, int ntris, uint32_t* masked) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_nodes) return;
    const float* nd = ref_nodes + (size_t)i * 10;
    const int off = __float_as_int(nd[8]), cnt = __float_as_int(nd[9]);
    if (cnt <= 0) return;
    if (!(nd[0] == nd[3] || nd[1] == nd[4] || nd[2] == nd[5])) return;
    for (int k = 0; k < cnt; ++k) {
        const int t = tri_indices[off + k];
        if (t < 0 || t >= ntris) continue;
        tris[t].v1 = tris[t].v0;
        tris[t].v2 = tris[t].v0;
        atomicAdd(masked, 1u);
    }
}

// Common tail of the two upload entry points.  `tris`/`mats`/`textures`/`pool` are host
// pointers when from_device is false and device pointers (on ctx's GPU) when it is true.
int scene_upload_impl(dsrt_ctx* ctx, bool from_device, const DsrtTriangle* tris, int ntris,
                      const DsrtMaterial* mats, int nmats, const DsrtSphere* spheres, int nspheres,
                      const DsrtTextureHeader* textures, int ntextures,
                      const float* texture_pool, int texture_pool_floats) {
    if (ntris < 0 || nmats < 0 || (ntris > 0 && !tris) || (nmats > 0 && !mats))
        return fail(ctx, DSRT_EINVAL, "dsrt_scene_upload: bad triangle/material arguments");
    if ((ntris > 0 || nspheres > 0) && nmats == 0) return fail(ctx, DSRT_EINVAL, "dsrt_scene_upload: geometry without materials");
    if (ntris >= (1 << kPoolIndexBits))
        return fail(ctx, DSRT_EUNSUPPORTED, "at most %d triangles (record positions are packed into %u bits, csrc/bvh8.cuh)", (1 << kPoolIndexBits) - 1, kPoolIndexBits);
    if (nspheres < 0 || (nspheres > 0 && !spheres)) return fail(ctx, DSRT_EINVAL, "dsrt_scene_upload: bad sphere arguments");
    // Spheres are few and are packed on the host (with the material table at hand, for the emissive ones): arrays that
    // live on the device come back first.
    std::vector<DsrtSphere> h_spheres;
    std::vector<DsrtMaterial> h_mats;
    if (from_device && nspheres > 0) {
        CK(cudaSetDevice(ctx->device));
        h_spheres.resize((size_t)nspheres);
        h_mats.resize((size_t)nmats);
        CK(cudaMemcpy(h_spheres.data(), spheres, sizeof(DsrtSphere) * (size_t)nspheres, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h_mats.data(), mats, sizeof(DsrtMaterial) * (size_t)nmats, cudaMemcpyDeviceToHost));
        spheres = h_spheres.data();
    }
    const DsrtMaterial* mats_host = (from_device && nspheres > 0) ? h_mats.data() : mats;
    for (int i = 0; i < nspheres; ++i)
        if (spheres[i].material_id < 0 || spheres[i].material_id >= nmats)
            return fail(ctx, DSRT_EINVAL, "sphere %d has material_id %d outside [0,%d)", i, spheres[i].material_id, nmats);
    if (ntextures < 0 || (ntextures > 0 && (!textures || !texture_pool || texture_pool_floats <= 0)))
        return fail(ctx, DSRT_EINVAL, "dsrt_scene_upload: bad texture arguments");
    if (!from_device)
        for (int i = 0; i < ntris; ++i)
            if (tris[i].material_id < 0 || tris[i].material_id >= nmats)
                return fail(ctx, DSRT_EINVAL, "triangle %d has material_id %d outside [0,%d)", i, tris[i].material_id, nmats);
    CK(cudaSetDevice(ctx->device));
    {
        const int rc = sync_all_slots(ctx);      // no frame may be in flight on the scene that is replaced
        if (rc) return rc;
    }
    free_scene(ctx);
    cudaEvent_t* ev_frame = ctx->slots[0].ev_frame;
    const cudaMemcpyKind kind = from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;

    DsrtSceneStats st{};
    st.num_triangles = ntris;
    st.max_leaf_tris = kMaxLeafTris;

    if (nmats > 0) {
        static_assert(sizeof(DsrtMaterial) == 48, "material layout");
        CK(cudaMalloc(&ctx->d_mats, sizeof(DsrtMaterial) * (size_t)nmats));
        CK(cudaMemcpyAsync(ctx->d_mats, mats, sizeof(DsrtMaterial) * (size_t)nmats, kind, ctx->stream));
        ctx->nmats = nmats;
        // which material classes occur (k_shade bins its hits by class only when there is more than one)
        std::vector<DsrtMaterial> tmp;
        const DsrtMaterial* hm = mats_host;
        if (from_device && h_mats.empty() && nmats <= kShadeClassTable) {
            tmp.resize((size_t)nmats);
            CK(cudaMemcpy(tmp.data(), mats, sizeof(DsrtMaterial) * (size_t)nmats, cudaMemcpyDeviceToHost));
            hm = tmp.data();
        }
        bool seen[3] = {false, false, false};
        if (!from_device || hm != mats) {
            for (int i = 0; i < nmats; ++i) seen[hm[i].type == DSRT_MAT_METAL ? 1 : (hm[i].type == DSRT_MAT_DIELECTRIC ? 2 : 0)] = true;
            ctx->mat_classes = (int)seen[0] + (int)seen[1] + (int)seen[2];
        } else {
            ctx->mat_classes = 1;       // table too large to bin anyway
        }
    }
    if (ntextures > 0) {
        CK(cudaMalloc(&ctx->d_tex, sizeof(DsrtTextureHeader) * (size_t)ntextures));
        CK(cudaMalloc(&ctx->d_pool, sizeof(float) * (size_t)texture_pool_floats));
        CK(cudaMemcpyAsync(ctx->d_tex, textures, sizeof(DsrtTextureHeader) * (size_t)ntextures, kind, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_pool, texture_pool, sizeof(float) * (size_t)texture_pool_floats, kind, ctx->stream));
        ctx->ntex = ntextures;
        ctx->pool_floats = texture_pool_floats;
    }
    if (nspheres > 0) {
        // Sphere records for the brute-force loop of scene_hit (gpu_render.cu:530-546) and the list
        // of emissive spheres that ray_color rebuilds at every diffuse hit (gpu_render.cu:840-848).
        static_assert(sizeof(DsrtSphere) == 24, "sphere layout");
        std::vector<float4> rec((size_t)nspheres * 2);
        std::vector<int> lights;
        for (int i = 0; i < nspheres; ++i) {
            const DsrtSphere& sp = spheres[i];
            rec[2 * i] = make_float4(sp.center.x, sp.center.y, sp.center.z, sp.radius);
            float mid;
            memcpy(&mid, &sp.material_id, 4);
            rec[2 * i + 1] = make_float4(1.0f / sp.radius, mid, 0.f, 0.f);       // IEEE division, as :501
            const DsrtMaterial& m = mats_host[sp.material_id];
            if (m.type == DSRT_MAT_DIFFUSE_LIGHT && (m.emissive.x > 0 || m.emissive.y > 0 || m.emissive.z > 0))
                lights.push_back(i);
        }
        CK(cudaMalloc(&ctx->d_spheres, sizeof(float4) * rec.size()));
        CK(cudaMemcpy(ctx->d_spheres, rec.data(), sizeof(float4) * rec.size(), cudaMemcpyHostToDevice));
        ctx->nspheres = nspheres;
        if (!lights.empty()) {
            CK(cudaMalloc(&ctx->d_sphere_lights, sizeof(int) * lights.size()));
            CK(cudaMemcpy(ctx->d_sphere_lights, lights.data(), sizeof(int) * lights.size(), cudaMemcpyHostToDevice));
            ctx->nsphere_lights = (int)lights.size();
        }
    }
    if (ntris > 0) {
        static_assert(sizeof(DsrtTriangle) == 116, "triangle layout");
        DsrtTriangle* d_raw = nullptr;
        if (!from_device) {
            CK(cudaMalloc(&d_raw, sizeof(DsrtTriangle) * (size_t)ntris));
            cudaEventRecord(ev_frame[0], ctx->stream);
            cudaError_t e = cudaMemcpyAsync(d_raw, tris, sizeof(DsrtTriangle) * (size_t)ntris, cudaMemcpyHostToDevice, ctx->stream);
            cudaEventRecord(ev_frame[1], ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { cudaFree(d_raw); return fail(ctx, DSRT_ECUDA, "triangle upload: %s", cudaGetErrorString(e)); }
            cudaEventElapsedTime(&st.upload_ms, ev_frame[0], ev_frame[1]);
        }
        if (from_device) {
            // the host path range-checks material ids above; here a kernel does (an id outside the table would index
            // past the material array in k_shade)
            uint32_t* flag = ctx->slots[0].counters + C_COUNT + 1;
            CK(cudaMemsetAsync(flag, 0, sizeof(uint32_t), ctx->stream));
            k_check_material_ids<<<(ntris + 255) / 256, 256, 0, ctx->stream>>>(tris, ntris, nmats, flag);
            uint32_t bad = 0;
            CK(cudaMemcpyAsync(&bad, flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (bad != 0u) return fail(ctx, DSRT_EINVAL, "%u triangles have a material_id outside [0,%d)", bad, nmats);
        }
        {
            float* sum = reinterpret_cast<float*>(ctx->slots[0].counters + C_COUNT + 4);
            CK(cudaMemsetAsync(sum, 0, sizeof(float), ctx->stream));
            k_sum_extent<<<(ntris + 255) / 256, 256, 0, ctx->stream>>>(from_device ? tris : d_raw, ntris, sum);
            float total = 0.f;
            CK(cudaMemcpyAsync(&total, sum, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->mean_tri_extent = total / (float)ntris;
        }
        char msg[256] = "";
        int rc = build_bvh8(ctx->stream, from_device ? tris : d_raw, ntris, ntextures > 0, ctx->builder, &ctx->bvh, msg, sizeof(msg));
        if (d_raw) cudaFree(d_raw);
        if (rc != 0) return fail(ctx, DSRT_ECUDA, "BVH build failed: %s", msg);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    st.num_wide_nodes = ctx->bvh.num_nodes;
    st.num_levels = ctx->bvh.num_levels;
    st.build_ms = ctx->bvh.build_ms;
    st.sah_cost = ctx->bvh.sah_cost;
    for (int a = 0; a < 3; ++a) { st.scene_lo[a] = ctx->bvh.scene_lo[a]; st.scene_hi[a] = ctx->bvh.scene_hi[a]; }
    st.node_bytes = (uint64_t)ctx->bvh.num_nodes * (16u * kNodeWords);
    st.tri_bytes = (uint64_t)ntris * 48u;
    ctx->sstats = st;
    ctx->have_scene = true;
    return DSRT_OK;
}

}  // namespace

extern "C" int dsrt_scene_upload(dsrt_ctx* ctx, const DsrtTriangle* tris, int ntris,
                                 const DsrtMaterial* mats, int nmats,
                                 const DsrtSphere* spheres, int nspheres,
                                 const DsrtTextureHeader* textures, int ntextures,
                                 const float* texture_pool, int texture_pool_floats) {
    if (!ctx) return DSRT_EINVAL;
    return scene_upload_impl(ctx, false, tris, ntris, mats, nmats, spheres, nspheres, textures, ntextures,
                             texture_pool, texture_pool_floats);
}

extern "C" int dsrt_scene_upload_device(dsrt_ctx* ctx, const DsrtTriangle* tris_device, int ntris,
                                        const DsrtMaterial* mats_device, int nmats,
                                        const DsrtSphere* spheres_device, int nspheres,
                                        const DsrtTextureHeader* textures_device, int ntextures,
                                        const float* texture_pool_device, int texture_pool_floats) {
    if (!ctx) return DSRT_EINVAL;
    return scene_upload_impl(ctx, true, tris_device, ntris, mats_device, nmats, spheres_device, nspheres,
                             textures_device, ntextures, texture_pool_device, texture_pool_floats);
}

extern "C" int dsrt_scene_upload_legacy(dsrt_ctx* ctx, const DsrtLegacyScene* scene, int reference_flat_boxes, int* masked_triangles) {
    if (!ctx || !scene) return DSRT_EINVAL;
    if (masked_triangles) *masked_triangles = 0;
    const DsrtTriangle* tris = scene->triangles;
    DsrtTriangle* copy = nullptr;
    if (reference_flat_boxes && scene->num_triangles > 0 && scene->bvh_nodes && scene->num_bvh_nodes > 0 && scene->tri_indices) {
        CK(cudaSetDevice(ctx->device));
        const int rc0 = sync_all_slots(ctx);
        if (rc0) return rc0;
        const size_t bytes = sizeof(DsrtTriangle) * (size_t)scene->num_triangles;
        CK(cudaMalloc(&copy, bytes));
        uint32_t* counter = ctx->slots[0].counters + C_COUNT + 2;
        cudaError_t e = cudaMemcpyAsync(copy, scene->triangles, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), ctx->stream);
        if (e == cudaSuccess) {
            k_mask_flat_leaves<<<(scene->num_bvh_nodes + 127) / 128, 128, 0, ctx->stream>>>(
                (const float*)scene->bvh_nodes, scene->num_bvh_nodes, scene->tri_indices, copy, scene->num_triangles, counter);
            e = cudaGetLastError();
        }
        uint32_t n = 0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&n, counter, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cudaFree(copy); return fail(ctx, DSRT_ECUDA, "flat-leaf masking: %s", cudaGetErrorString(e)); }
        if (masked_triangles) *masked_triangles = (int)n;
        tris = copy;
    }
    const int rc = scene_upload_impl(ctx, true, tris, scene->num_triangles, scene->materials, scene->num_materials,
                                     scene->spheres, scene->num_spheres, scene->textures, scene->num_textures,
                                     scene->texture_pool, scene->texture_pool_floats);
    if (copy) cudaFree(copy);
    return rc;
}

extern "C" int dsrt_scene_stats(dsrt_ctx* ctx, DsrtSceneStats* out) {
    if (!ctx || !out) return DSRT_EINVAL;
    if (!ctx->have_scene) return fail(ctx, DSRT_ENOSCENE, "no scene uploaded");
    *out = ctx->sstats;
    return DSRT_OK;
}

extern "C" int dsrt_scene_download(dsrt_ctx* ctx, void* nodes, size_t node_bytes, void* tri_records, size_t tri_bytes) {
    if (!ctx) return DSRT_EINVAL;
    if (!ctx->have_scene) return fail(ctx, DSRT_ENOSCENE, "no scene uploaded");
    CK(cudaSetDevice(ctx->device));
    const size_t nb = (size_t)ctx->bvh.num_nodes * (16u * kNodeWords), tb = (size_t)ctx->bvh.num_tris * 48u;
    if ((nodes && node_bytes < nb) || (tri_records && tri_bytes < tb))
        return fail(ctx, DSRT_EINVAL, "dsrt_scene_download: need %zu node bytes and %zu triangle bytes", nb, tb);
    if (nodes && nb) CK(cudaMemcpyAsync(nodes, ctx->bvh.nodes, nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (tri_records && tb) CK(cudaMemcpyAsync(tri_records, ctx->bvh.tri_records, tb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DSRT_OK;
}

extern "C" int dsrt_set_stage_timing(dsrt_ctx* ctx, int enabled) {
    if (!ctx) return DSRT_EINVAL;
    ctx->stage_timing = enabled ? 1 : 0;
    return DSRT_OK;
}

extern "C" int dsrt_set_bvh_builder(dsrt_ctx* ctx, int builder) {
    if (!ctx || (builder != DSRT_BUILDER_PLOC && builder != DSRT_BUILDER_LBVH)) return DSRT_EINVAL;
    ctx->builder = builder;
    return DSRT_OK;
}

extern "C" int dsrt_set_samples_per_batch(dsrt_ctx* ctx, int samples) {
    if (!ctx || samples < 0) return DSRT_EINVAL;
    ctx->spb_override = samples;
    return DSRT_OK;
}

namespace {

void fill_scene_consts(dsrt_ctx* ctx, FrameConst& fc) {
    fc.nodes = ctx->bvh.nodes;
    fc.tri_records = ctx->bvh.tri_records;
    fc.shade_records = ctx->bvh.shade_records;
    fc.shade_compact = ctx->bvh.shade_compact ? 1 : 0;
    fc.uv_records = ctx->bvh.uv_records;
    fc.materials = ctx->d_mats;
    fc.textures = ctx->d_tex;
    fc.texture_pool = ctx->d_pool;
    fc.num_nodes = ctx->bvh.num_nodes;
    fc.num_materials = ctx->nmats;
    fc.shade_bins = ctx->mat_classes > 1 && ctx->nmats <= kShadeClassTable;
    {
        float ext = 0.f;
        for (int k = 0; k < 3; ++k) {
            ext = std::max(ext, ctx->bvh.scene_hi[k] - ctx->bvh.scene_lo[k]);
            ext = std::max(ext, std::max(std::fabs(ctx->bvh.scene_lo[k]), std::fabs(ctx->bvh.scene_hi[k])));
        }
        const float pad = ext * (1.0f / 65536.0f);
        for (int k = 0; k < 3; ++k) { fc.scene_lo[k] = ctx->bvh.scene_lo[k] - pad; fc.scene_hi[k] = ctx->bvh.scene_hi[k] + pad; }
    }
    fc.num_textures = ctx->ntex;
    fc.texture_pool_floats = ctx->pool_floats;
    fc.spheres = ctx->d_spheres;
    fc.sphere_lights = ctx->d_sphere_lights;
    fc.num_spheres = ctx->nspheres;
    fc.num_sphere_lights = ctx->nsphere_lights;
}

// Pixel bounds outside of which no camera ray can hit the (padded) scene box: the box corners are
// projected through the reference camera model dir(u,v) = llc + u h + v vert - origin
// (gpu_render.cu:952-967), i.e. corner - origin = lambda (a + u h + v vert) solved for
// (lambda, lambda u, lambda v) in double precision.  A convex box entirely in front of the camera
// projects into the bounding rectangle of its projected corners; the rectangle is widened by two
// pixels, which also covers the jitter ((x + jx) / (W - 1) with 0 <= jx < 1).  If any corner is not
// safely in front, the bounds are the whole frame.  k_primary uses them to end rays that cannot
// hit before generating them; rays inside the bounds still get the exact per-ray box test.
bool scene_screen_bounds(const FrameConst& fc, int W, int H, int out[4]) {
    out[0] = 0; out[1] = 0; out[2] = W - 1; out[3] = H - 1;
    double a[3], h[3], v[3];
    for (int k = 0; k < 3; ++k) { a[k] = (double)fc.cam_llc[k] - fc.cam_o[k]; h[k] = fc.cam_h[k]; v[k] = fc.cam_v[k]; }
    auto det3 = [](const double* c0, const double* c1, const double* c2) {
        return c0[0] * (c1[1] * c2[2] - c1[2] * c2[1]) - c1[0] * (c0[1] * c2[2] - c0[2] * c2[1]) + c2[0] * (c0[1] * c1[2] - c0[2] * c1[1]);
    };
    const double det = det3(a, h, v);
    if (!(std::fabs(det) > 0.0) || !std::isfinite(det)) return false;
    double umin = 1e300, umax = -1e300, vmin = 1e300, vmax = -1e300;
    for (int c = 0; c < 8; ++c) {
        double p[3];
        for (int k = 0; k < 3; ++k) p[k] = (double)((c >> k) & 1 ? fc.scene_hi[k] : fc.scene_lo[k]) - fc.cam_o[k];
        const double lam = det3(p, h, v) / det, lu = det3(a, p, v) / det, lv = det3(a, h, p) / det;
        if (!(lam > 1e-9) || !std::isfinite(lam)) return false;          // a corner beside or behind the camera: no bound
        const double u = lu / lam, vv = lv / lam;
        if (!std::isfinite(u) || !std::isfinite(vv)) return false;
        umin = std::min(umin, u); umax = std::max(umax, u); vmin = std::min(vmin, vv); vmax = std::max(vmax, vv);
    }
    auto clampi = [](double x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : (int)x); };
    out[0] = clampi(std::floor(umin * (W - 1)) - 3.0, 0, W - 1);
    out[2] = clampi(std::ceil(umax * (W - 1)) + 2.0, 0, W - 1);
    out[1] = clampi(std::floor(vmin * (H - 1)) - 3.0, 0, H - 1);
    out[3] = clampi(std::ceil(vmax * (H - 1)) + 2.0, 0, H - 1);
    return true;
}

void fill_camera(const DsrtCamera* cam, FrameConst& fc) {
    const DsrtVec3* src[4] = {&cam->origin, &cam->lower_left_corner, &cam->horizontal, &cam->vertical};
    float* dst[4] = {fc.cam_o, fc.cam_llc, fc.cam_h, fc.cam_v};
    for (int k = 0; k < 4; ++k) { dst[k][0] = src[k]->x; dst[k][1] = src[k]->y; dst[k][2] = src[k]->z; }
}

}  // namespace

namespace {

// What one frame call asks for, validated.
struct FrameArgs {
    int W, H, spp_total, x0, y0, x1, y1, rng_mode, sample_begin, sample_end, sun_enabled;
    bool full_rect;
};

int check_frame_args(dsrt_ctx* ctx, const DsrtCamera* camera, const float sun_dir[3], const float sun_radiance[3], int sun_enabled,
                     const DsrtRenderParams* params, int sample_begin, int sample_end, const int rect[4], FrameArgs* a) {
    if (!camera || !params) return fail(ctx, DSRT_EINVAL, "dsrt_render_frame: camera/params NULL");
    if (!ctx->have_scene) return fail(ctx, DSRT_ENOSCENE, "dsrt_render_frame before dsrt_scene_upload");
    a->W = params->img_width > 0 ? params->img_width : camera->image_width;
    a->H = params->img_height > 0 ? params->img_height : camera->image_height;
    if (a->W < 2 || a->H < 2) return fail(ctx, DSRT_EINVAL, "image must be at least 2x2 (got %dx%d)", a->W, a->H);
    a->spp_total = params->samples_per_pixel < 1 ? 1 : params->samples_per_pixel;
    if (sample_begin < 0 || sample_end < sample_begin || sample_end > a->spp_total)
        return fail(ctx, DSRT_EINVAL, "sample range [%d,%d) outside [0,%d)", sample_begin, sample_end, a->spp_total);
    a->sample_begin = sample_begin; a->sample_end = sample_end;
    a->x0 = 0; a->y0 = 0; a->x1 = a->W; a->y1 = a->H;
    if (rect) { a->x0 = rect[0]; a->y0 = rect[1]; a->x1 = rect[2]; a->y1 = rect[3]; }
    if (a->x0 < 0 || a->y0 < 0 || a->x1 > a->W || a->y1 > a->H || a->x1 <= a->x0 || a->y1 <= a->y0)
        return fail(ctx, DSRT_EINVAL, "bad rectangle [%d,%d)x[%d,%d) for %dx%d", a->x0, a->x1, a->y0, a->y1, a->W, a->H);
    if (sun_enabled && (!sun_dir || !sun_radiance)) return fail(ctx, DSRT_EINVAL, "sun enabled but sun_dir/sun_radiance NULL");
    a->sun_enabled = sun_enabled ? 1 : 0;
    a->rng_mode = params->rng_mode;
    if (a->rng_mode != DSRT_RNG_PHILOX && a->rng_mode != DSRT_RNG_REF_LCG)
        return fail(ctx, DSRT_EINVAL, "unknown rng_mode %d", a->rng_mode);
    a->full_rect = (a->x0 == 0 && a->y0 == 0 && a->x1 == a->W && a->y1 == a->H);
    return DSRT_OK;
}

// The frame constants that do not depend on the slot's work memory.
void fill_frame_consts(dsrt_ctx* ctx, const FrameArgs& a, const DsrtCamera* camera, const float sun_dir[3], const float sun_radiance[3],
                       const DsrtRenderParams* params, uint64_t seed, FrameConst& fc) {
    memset(&fc, 0, sizeof(fc));
    fill_scene_consts(ctx, fc);
    fill_camera(camera, fc);
    fc.W = a.W; fc.H = a.H;
    fc.x0 = a.x0; fc.y0 = a.y0; fc.rw = a.x1 - a.x0; fc.rh = a.y1 - a.y0;
    {
        int hb[4];
        const bool bounded = scene_screen_bounds(fc, a.W, a.H, hb);
        fc.hit_x0 = hb[0]; fc.hit_y0 = hb[1]; fc.hit_x1 = hb[2]; fc.hit_y1 = hb[3];
        fc.cull_pixels = (bounded && fc.num_spheres == 0 && a.rng_mode == DSRT_RNG_PHILOX) ? 1 : 0;
    }
    fc.tiles_x = (fc.rw + 7) / 8;
    const int tiles_y = (fc.rh + 3) / 4;
    fc.slots_per_sample = fc.tiles_x * tiles_y * 32;
    fc.inv_tiles_x = 1.0f / (float)fc.tiles_x;
    fc.inv_slots_per_sample = 1.0f / (float)fc.slots_per_sample;
    fc.spp_total = a.spp_total;
    fc.max_depth = params->max_depth > 0 ? params->max_depth : 12;
    fc.rng_mode = a.rng_mode;
    fc.seed_lo = (uint32_t)(seed & 0xFFFFFFFFu);
    fc.seed_hi = (uint32_t)(seed >> 32);
    fc.sun_enabled = a.sun_enabled;
    if (a.sun_enabled) {
        host_ldir(sun_dir, fc.ldir);
        for (int k = 0; k < 3; ++k) fc.sun_rad[k] = sun_radiance[k];
    }
    // camera rays start at per-tile entry nodes, and tiles with nothing in sight are skipped outright -- unless the
    // scene has spheres (they lie outside the triangle BVH) or the LCG replay runs (its generator advances per pixel)
    fc.tile_walk = (ctx->use_tile_entries && fc.num_nodes > 0) ? 1 : 0;
    fc.cull_tiles = (fc.tile_walk && fc.num_spheres == 0 && a.rng_mode == DSRT_RNG_PHILOX) ? 1 : 0;
}

// Per-tile candidate lists (render.cu: tile_candidates, k_primary_lists) pay when a tile's footprint on the scene is
// of the order of a triangle or smaller; far away a tile covers dozens of triangles, every list would overflow, and
// trying costs ~0.15 ms of a 1080p frame.  The footprint of the 8-pixel tile width at the nearest point of the scene box
// against the mean triangle size decides for the whole frame (a performance choice only: a tile without a list
// traverses).  DSRT_TILE_LISTS=0 switches lists off, =2 always tries.  Not with spheres (outside the triangle BVH).
bool want_tile_lists(const dsrt_ctx* ctx, const FrameConst& fc) {
    if (!ctx->use_tile_lists || !fc.tile_walk || fc.num_spheres != 0) return false;
    if (ctx->use_tile_lists >= 2) return true;
    double d2 = 0.0;
    for (int k = 0; k < 3; ++k) {
        const double o = fc.cam_o[k], lo = ctx->bvh.scene_lo[k], hi = ctx->bvh.scene_hi[k];
        const double g = o < lo ? lo - o : (o > hi ? o - hi : 0.0);
        d2 += g * g;
    }
    // the image plane sits at |llc + h/2 + v/2 - o| along the view axis and is |h| wide for W - 1 pixel steps
    double fd2 = 0.0, h2 = 0.0;
    for (int k = 0; k < 3; ++k) {
        const double c = (double)fc.cam_llc[k] + 0.5 * fc.cam_h[k] + 0.5 * fc.cam_v[k] - fc.cam_o[k];
        fd2 += c * c;
        h2 += (double)fc.cam_h[k] * fc.cam_h[k];
    }
    if (!(fd2 > 0.0) || fc.W < 2) return false;
    const double pixel_per_unit = std::sqrt(h2 / fd2) / (double)(fc.W - 1);
    const double footprint = 8.0 * pixel_per_unit * std::sqrt(d2);
    return footprint <= 4.0 * (double)ctx->mean_tri_extent;
}

void bind_slot(FrameSlot& f, FrameConst& fc) {
    for (int k = 0; k < 2; ++k) { fc.rq_o[k] = f.rq_o[k]; fc.rq_d[k] = f.rq_d[k]; }
    fc.hits = f.hits; fc.st_a = f.st_a; fc.st_b = f.st_b; fc.sq_o = f.sq_o; fc.sq_l = f.sq_l;
    fc.pq = f.pq;
    fc.dbg_leaf = nullptr;
    fc.rng_pix = f.rng_pix;
    fc.accum = f.accum;
    fc.counters = f.counters;
    fc.n_active = f.counters + C_COUNT;
    fc.stats = f.stats;
    fc.tile_entries = f.tile_table;
    fc.active_tiles = f.active_tiles;
    fc.tile_lists = f.tile_lists;
    fc.list_tiles = f.list_tiles;
    fc.n_list = f.counters + C_COUNT + 3;
}

int ensure_tiles(dsrt_ctx* ctx, FrameSlot& f, const FrameConst& fc) {
    const size_t ntiles = (size_t)fc.slots_per_sample >> 5;
    if (ntiles <= f.tile_table_tiles) return DSRT_OK;
    cudaFree(f.tile_table); cudaFree(f.active_tiles); cudaFree(f.tile_lists); cudaFree(f.list_tiles);
    f.tile_table = nullptr; f.active_tiles = nullptr; f.tile_lists = nullptr; f.list_tiles = nullptr; f.tile_table_tiles = 0;
    CK(cudaMalloc(&f.tile_table, ntiles * 8 * sizeof(uint2)));       // up to 8 start nodes per tile (render.cu: DSRT_TILE_ENTRY_COUNT)
    CK(cudaMalloc(&f.active_tiles, ntiles * sizeof(uint32_t)));
    CK(cudaMalloc(&f.tile_lists, ntiles * (size_t)(kTileListCap + 1) * sizeof(uint2)));
    CK(cudaMalloc(&f.list_tiles, ntiles * sizeof(uint32_t)));
    f.tile_table_tiles = ntiles;
    return DSRT_OK;
}

// Enqueues one frame on the slot's stream.  No host synchronisation: queue sizes live in device memory, the launch
// sequence is fixed, and launches behind an emptied queue are no-ops.
int enqueue_frame(dsrt_ctx* ctx, int slot, const DsrtCamera* camera,
                  const float sun_dir[3], const float sun_radiance[3], int sun_enabled,
                  const DsrtRenderParams* params, uint64_t seed,
                  int sample_begin, int sample_end, const int rect[4], int keep_accum) {
    FrameArgs a;
    int rc = check_frame_args(ctx, camera, sun_dir, sun_radiance, sun_enabled, params, sample_begin, sample_end, rect, &a);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    rc = slot_init(ctx, slot);
    if (rc) return rc;
    FrameSlot& f = ctx->slots[slot];
    if (f.pending) return fail(ctx, DSRT_EINVAL, "slot %d still has a frame in flight: call dsrt_frame_wait first", slot);
    bool fresh = false;
    rc = ensure_image(ctx, f, a.W, a.H, &fresh);
    if (rc) return rc;
    if (fresh || !keep_accum)
        CK(cudaMemsetAsync(f.accum, 0, (size_t)a.W * a.H * 3 * sizeof(float), f.stream));

    if (a.rng_mode == DSRT_RNG_REF_LCG) {
        // The replay keeps one generator per pixel that runs through the samples in order.
        if (a.sample_begin == 0) {
            launch_lcg_seed(f.rng_pix, a.W, a.H, (uint32_t)(seed & 0xFFFFFFFFu), f.stream);
        } else if (a.sample_begin != f.lcg_next_sample || !a.full_rect) {
            return fail(ctx, DSRT_EINVAL, "LCG replay renders samples in order from 0 (next expected %d, got %d)",
                        f.lcg_next_sample, a.sample_begin);
        }
        f.lcg_next_sample = a.full_rect ? a.sample_end : -1;
    }

    FrameConst fc;
    fill_frame_consts(ctx, a, camera, sun_dir, sun_radiance, params, seed, fc);

    // Samples of a pixel in flight together.  Large batches keep the deep-bounce launches
    // (few surviving rays) big enough to fill the GPU (profiles/r01_a: 2x from spb 4 -> 32);
    // the default takes as many samples as fit the work-memory budget (ctx->work_gb per slot).
    int spb = 1;
    const int nsamples = a.sample_end - a.sample_begin;
    if (a.rng_mode == DSRT_RNG_PHILOX) {
        if (ctx->spb_override > 0) spb = ctx->spb_override;
        else {
            const long long target = (long long)ctx->work_gb * 1000 * 1000 * 1000 / kWorkBytesPerPath;
            spb = (int)std::max(1ll, std::min(256ll, target / std::max(1, fc.slots_per_sample)));
        }
        spb = std::max(1, std::min(spb, std::max(1, nsamples)));
        // equal batches: 250 samples at a limit of 200 run as 125 + 125, not 200 + 50
        if (ctx->spb_override <= 0 && nsamples > spb) {
            const int nb = (nsamples + spb - 1) / spb;
            spb = (nsamples + nb - 1) / nb;
        }
    }
    // The budget is a wish, not a requirement: when the device cannot give that much (another
    // process on the GPU, another slot), halve the batch until the allocation succeeds.
    for (;;) {
        rc = ensure_work(ctx, f, (size_t)fc.slots_per_sample * spb);
        if (rc == DSRT_OK || spb == 1 || a.rng_mode != DSRT_RNG_PHILOX) break;
        cudaGetLastError();           // clear the allocation failure
        spb = (spb + 1) / 2;
    }
    if (rc) return rc;
    rc = ensure_tiles(ctx, f, fc);
    if (rc) return rc;
    bind_slot(f, fc);
    if (!want_tile_lists(ctx, fc)) fc.n_list = nullptr;

    CK(cudaMemsetAsync(f.stats, 0, sizeof(unsigned long long) * S_COUNT, f.stream));
    CK(cudaEventRecord(f.ev_frame[0], f.stream));

    // traversal entry points of the camera rays and the list of tiles with anything in sight, per 8x4-pixel tile of
    // the rectangle (part of the frame's device time)
    launch_tile_entries(fc, f.tile_table, f.active_tiles, f.stream);

    uint32_t launches = 1, batches = 0;
    f.nev = 0;
    f.ev_kind.clear();
    const bool timing = ctx->stage_timing != 0;
    auto mark = [&](int kind_of_next_interval) {
        if (!timing || f.nev >= (size_t)kMaxEvents) return;
        cudaEvent_t e = pool_event(f, f.nev);
        if (!e) return;
        cudaEventRecord(e, f.stream);
        f.ev_kind.push_back(kind_of_next_interval);
        ++f.nev;
    };

    const int wave_depth = std::min(fc.max_depth, ctx->wave_depth);
    for (int s0 = a.sample_begin; s0 < a.sample_end; s0 += spb) {
        fc.s0 = s0;
        fc.spb = std::min(spb, a.sample_end - s0);
        launch_batch_begin(fc, f.stream);
        mark(2);
        launch_primary(fc, ctx->dims, false, f.stream);
        launch_primary_lists(fc, ctx->dims, false, f.stream);
        launches += fc.n_list ? 3 : 2;
        int cur = 0;
        for (int depth = 0; depth < wave_depth; ++depth) {
            mark(1);
            launch_shade(fc, cur, depth, ctx->dims, f.stream);
            mark(-1);
            launch_advance(fc, depth, f.stream);
            mark(0);
            launch_trace(fc, cur ^ 1, depth & 1, ctx->dims, f.stream);
            launches += 3;
            cur ^= 1;
            if (depth + 1 >= fc.max_depth) break;
            // Few rays left: one thread per path finishes them in a single launch (render.cu: k_tail).  Past the
            // wavefront depth the closing launch takes whatever is left, so no launch count depends on the queue
            // sizes and the host never waits inside a frame.
            const bool closing = depth + 1 == wave_depth;
            if (closing) {
                mark(0);
                launch_tail(fc, cur, depth + 1, kTailAll, ctx->dims, f.stream);
                ++launches;
            } else if (depth >= 1 && ctx->tail_threshold > 0) {
                launch_tail(fc, cur, depth + 1, (uint32_t)ctx->tail_threshold, ctx->dims, f.stream);
                ++launches;
            }
        }
        mark(-1);
        launch_resolve(fc, ctx->dims, f.stream);
        ++launches;
        ++batches;
    }
    mark(-1);
    CK(cudaEventRecord(f.ev_frame[1], f.stream));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(f.h_stats, f.stats, sizeof(unsigned long long) * S_COUNT, cudaMemcpyDeviceToHost, f.stream));
    f.launches = launches;
    f.batches = batches;
    f.pending = true;
    return DSRT_OK;
}

// Waits for everything enqueued on the slot and collects the frame's statistics.
int finish_frame(dsrt_ctx* ctx, int slot) {
    FrameSlot& f = ctx->slots[slot];
    if (!f.created) return fail(ctx, DSRT_EINVAL, "slot %d was never used", slot);
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(f.stream));
    if (!f.pending) return DSRT_OK;
    f.pending = false;
    DsrtFrameStats fs{};
    fs.primary_rays = f.h_stats[S_PRIMARY];
    fs.bounce_rays = f.h_stats[S_BOUNCE];
    fs.shadow_rays = f.h_stats[S_SHADOW];
    fs.paths = f.h_stats[S_PATHS];
    fs.culled_primary_rays = f.h_stats[S_CULLED];
    fs.kernel_launches = f.launches;
    fs.batches = f.batches;
    cudaEventElapsedTime(&fs.device_ms, f.ev_frame[0], f.ev_frame[1]);
    if (f.nev > 1) {
        float trav = 0.f, shade = 0.f, prim = 0.f;
        for (size_t k = 0; k + 1 < f.nev; ++k) {
            float ms = 0.f;
            if (f.ev_kind[k] < 0) continue;
            cudaEventElapsedTime(&ms, f.ev_pool[k], f.ev_pool[k + 1]);
            if (f.ev_kind[k] == 1) shade += ms; else trav += ms;
            if (f.ev_kind[k] == 2) prim += ms;
        }
        fs.extend_ms = trav; fs.primary_ms = prim; fs.shade_ms = shade;
    }
    f.fstats = fs;
    return DSRT_OK;
}

int check_slot(dsrt_ctx* ctx, int slot) {
    if (slot < 0 || slot >= kMaxSlots) return fail(ctx, DSRT_EINVAL, "slot %d outside [0,%d)", slot, kMaxSlots);
    return DSRT_OK;
}

int readback_rgb8_impl(dsrt_ctx* ctx, int slot, int total_spp, float gamma, unsigned char* dst, size_t bytes, bool wait) {
    FrameSlot& f = ctx->slots[slot];
    if (!f.accum) return fail(ctx, DSRT_EINVAL, "nothing rendered yet");
    const size_t need = (size_t)f.W * f.H * 3;
    if (bytes < need) return fail(ctx, DSRT_EINVAL, "dsrt_readback_rgb8: need %zu bytes, got %zu", need, bytes);
    CK(cudaSetDevice(ctx->device));
    const int spp = total_spp < 1 ? 1 : total_spp;
    const float g = gamma > 0.0f ? gamma : 1.0f;                   // gpu_render.cu:1043
    launch_tonemap(f.accum, f.d_rgb, f.W, f.H, 1.0f / (float)spp, 1.0f / g, f.stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dst, f.d_rgb, need, cudaMemcpyDeviceToHost, f.stream));
    if (wait) CK(cudaStreamSynchronize(f.stream));
    return DSRT_OK;
}

}  // namespace

extern "C" int dsrt_render_frame(dsrt_ctx* ctx, const DsrtCamera* camera,
                                 const float sun_dir[3], const float sun_radiance[3], int sun_enabled,
                                 const DsrtRenderParams* params, uint64_t seed,
                                 int sample_begin, int sample_end, const int rect[4], int keep_accum) {
    if (!ctx) return DSRT_EINVAL;
    if (ctx->slots[0].pending) { const int rc = finish_frame(ctx, 0); if (rc) return rc; }
    const int rc = enqueue_frame(ctx, 0, camera, sun_dir, sun_radiance, sun_enabled, params, seed, sample_begin, sample_end, rect, keep_accum);
    if (rc) return rc;
    return finish_frame(ctx, 0);
}

extern "C" int dsrt_render_frame_async(dsrt_ctx* ctx, int slot, const DsrtCamera* camera,
                                       const float sun_dir[3], const float sun_radiance[3], int sun_enabled,
                                       const DsrtRenderParams* params, uint64_t seed,
                                       int sample_begin, int sample_end, const int rect[4], int keep_accum) {
    if (!ctx) return DSRT_EINVAL;
    const int rc = check_slot(ctx, slot);
    if (rc) return rc;
    return enqueue_frame(ctx, slot, camera, sun_dir, sun_radiance, sun_enabled, params, seed, sample_begin, sample_end, rect, keep_accum);
}

extern "C" int dsrt_readback_rgb8_async(dsrt_ctx* ctx, int slot, int total_spp, float gamma, unsigned char* dst, size_t bytes) {
    if (!ctx || !dst) return DSRT_EINVAL;
    const int rc = check_slot(ctx, slot);
    if (rc) return rc;
    return readback_rgb8_impl(ctx, slot, total_spp, gamma, dst, bytes, false);
}

extern "C" int dsrt_frame_wait(dsrt_ctx* ctx, int slot, DsrtFrameStats* out) {
    if (!ctx) return DSRT_EINVAL;
    int rc = check_slot(ctx, slot);
    if (rc) return rc;
    rc = finish_frame(ctx, slot);
    if (rc) return rc;
    if (out) *out = ctx->slots[slot].fstats;
    return DSRT_OK;
}

extern "C" int dsrt_frame_stats(dsrt_ctx* ctx, DsrtFrameStats* out) {
    if (!ctx || !out) return DSRT_EINVAL;
    *out = ctx->slots[0].fstats;
    return DSRT_OK;
}

extern "C" int dsrt_box_screen_bounds(const DsrtCamera* camera, const float lo[3], const float hi[3],
                                      int width, int height, int out[4]) {
    if (!camera || !lo || !hi || !out || width < 2 || height < 2) return 0;
    FrameConst fc;
    memset(&fc, 0, sizeof(fc));
    fill_camera(camera, fc);
    for (int k = 0; k < 3; ++k) { fc.scene_lo[k] = lo[k]; fc.scene_hi[k] = hi[k]; }
    return scene_screen_bounds(fc, width, height, out) ? 1 : 0;
}

extern "C" int dsrt_host_alloc(void** ptr, size_t bytes) {
    if (!ptr || bytes == 0) return DSRT_EINVAL;
    *ptr = nullptr;
    return cudaMallocHost(ptr, bytes) == cudaSuccess ? DSRT_OK : DSRT_ENOMEM;
}

extern "C" void dsrt_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}

extern "C" int dsrt_readback_rgb8(dsrt_ctx* ctx, int total_spp, float gamma, unsigned char* dst, size_t bytes) {
    if (!ctx || !dst) return DSRT_EINVAL;
    return readback_rgb8_impl(ctx, 0, total_spp, gamma, dst, bytes, true);
}

extern "C" int dsrt_readback_accum(dsrt_ctx* ctx, float* dst, size_t bytes) {
    if (!ctx || !dst) return DSRT_EINVAL;
    FrameSlot& f = ctx->slots[0];
    if (!f.accum) return fail(ctx, DSRT_EINVAL, "nothing rendered yet");
    const size_t need = (size_t)f.W * f.H * 3 * sizeof(float);
    if (bytes < need) return fail(ctx, DSRT_EINVAL, "dsrt_readback_accum: need %zu bytes, got %zu", need, bytes);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst, f.accum, need, cudaMemcpyDeviceToHost, f.stream));
    CK(cudaStreamSynchronize(f.stream));
    return DSRT_OK;
}

extern "C" int dsrt_accum_device(dsrt_ctx* ctx, void** device_ptr, size_t* num_floats) {
    if (!ctx || !device_ptr || !num_floats) return DSRT_EINVAL;
    FrameSlot& f = ctx->slots[0];
    if (!f.accum) return fail(ctx, DSRT_EINVAL, "nothing rendered yet");
    *device_ptr = f.accum;
    *num_floats = (size_t)f.W * f.H * 3;
    return DSRT_OK;
}

extern "C" int dsrt_clear_accum(dsrt_ctx* ctx, int width, int height) {
    if (!ctx || width < 2 || height < 2) return DSRT_EINVAL;
    CK(cudaSetDevice(ctx->device));
    FrameSlot& f = ctx->slots[0];
    if (f.pending) { const int rc = finish_frame(ctx, 0); if (rc) return rc; }
    bool fresh = false;
    const int rc = ensure_image(ctx, f, width, height, &fresh);
    if (rc) return rc;
    CK(cudaMemsetAsync(f.accum, 0, (size_t)width * height * 3 * sizeof(float), f.stream));
    CK(cudaStreamSynchronize(f.stream));
    return DSRT_OK;
}

// ncclReduce resolved at run time: ncclResult_t ncclReduce(const void*, void*, size_t, ncclDataType_t,
// ncclRedOp_t, int root, ncclComm_t, cudaStream_t); ncclFloat = 7, ncclSum = 0 (nccl.h, stable ABI).
extern "C" int dsrt_reduce_accum(dsrt_ctx* ctx, void* nccl_comm, int root) {
    if (!ctx || !nccl_comm || root < 0) return DSRT_EINVAL;
    FrameSlot& f = ctx->slots[0];
    if (!f.accum) return fail(ctx, DSRT_EINVAL, "nothing rendered yet");
    typedef int (*reduce_fn)(const void*, void*, size_t, int, int, int, void*, cudaStream_t);
    // looked up once, whichever of the per-GPU worker threads gets here first (C++11 static initialisation is thread-safe)
    static const reduce_fn fn = [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        return h ? reinterpret_cast<reduce_fn>(dlsym(h, "ncclReduce")) : (reduce_fn) nullptr;
    }();
    if (!fn) return fail(ctx, DSRT_EUNSUPPORTED, "dsrt_reduce_accum: libnccl.so.2 with ncclReduce not found");
    CK(cudaSetDevice(ctx->device));
    const size_t nf = (size_t)f.W * f.H * 3;
    const int rc = fn(f.accum, f.accum, nf, /*ncclFloat*/ 7, /*ncclSum*/ 0, root, nccl_comm, ctx->stream);
    if (rc != 0) return fail(ctx, DSRT_ECUDA, "ncclReduce failed with code %d", rc);
    CK(cudaStreamSynchronize(ctx->stream));
    return DSRT_OK;
}

extern "C" int dsrt_upload_accum(dsrt_ctx* ctx, const float* src, size_t bytes) {
    if (!ctx || !src) return DSRT_EINVAL;
    FrameSlot& f = ctx->slots[0];
    if (!f.accum) return fail(ctx, DSRT_EINVAL, "nothing rendered yet");
    const size_t need = (size_t)f.W * f.H * 3 * sizeof(float);
    if (bytes != need) return fail(ctx, DSRT_EINVAL, "dsrt_upload_accum: need %zu bytes, got %zu", need, bytes);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(f.accum, src, need, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DSRT_OK;
}

// ---- parity hooks -------------------------------------------------------------------------
namespace {

int download_hits(dsrt_ctx* ctx, int n, int32_t* tri, float* t, float* u, float* v) {
    std::vector<float4> h((size_t)n);
    CK(cudaMemcpyAsync(h.data(), ctx->t_hits, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; ++i) {
        int32_t id;
        memcpy(&id, &h[i].w, 4);
        if (tri) tri[i] = id;
        if (t) t[i] = h[i].x;
        if (u) u[i] = h[i].y;
        if (v) v[i] = h[i].z;
    }
    return DSRT_OK;
}

}  // namespace

extern "C" int dsrt_trace_closest_device(dsrt_ctx* ctx, const void* rays_device, int n, void* hits_device) {
    if (!ctx || n < 0 || (n > 0 && (!rays_device || !hits_device))) return DSRT_EINVAL;
    if (!ctx->have_scene) return fail(ctx, DSRT_ENOSCENE, "trace before dsrt_scene_upload");
    if (n == 0) return DSRT_OK;
    CK(cudaSetDevice(ctx->device));
    launch_trace_closest(ctx->bvh.nodes, ctx->bvh.tri_records, ctx->bvh.num_nodes, SphereSet{ctx->d_spheres, ctx->nspheres},
                         (const float4*)rays_device, n, (float4*)hits_device,
                         ctx->slots[0].counters + C_FETCH, ctx->dims.trace_blocks, ctx->stream);
    CK(cudaGetLastError());
    return DSRT_OK;
}

extern "C" int dsrt_trace_occluded_device(dsrt_ctx* ctx, const void* rays_device, int n, void* occluded_device) {
    if (!ctx || n < 0 || (n > 0 && (!rays_device || !occluded_device))) return DSRT_EINVAL;
    if (!ctx->have_scene) return fail(ctx, DSRT_ENOSCENE, "trace before dsrt_scene_upload");
    if (n == 0) return DSRT_OK;
    CK(cudaSetDevice(ctx->device));
    launch_trace_occluded(ctx->bvh.nodes, ctx->bvh.tri_records, ctx->bvh.num_nodes, SphereSet{ctx->d_spheres, ctx->nspheres},
                          (const float4*)rays_device, n, (uint8_t*)occluded_device,
                          ctx->slots[0].counters + C_FETCH, ctx->dims.trace_blocks, ctx->stream);
    CK(cudaGetLastError());
    return DSRT_OK;
}

extern "C" int dsrt_trace_closest(dsrt_ctx* ctx, const float* rays, int n,
                                  int32_t* tri, float* t, float* u, float* v) {
    if (!ctx || n < 0 || (n > 0 && !rays)) return DSRT_EINVAL;
    if (!ctx->have_scene) return fail(ctx, DSRT_ENOSCENE, "trace before dsrt_scene_upload");
    if (n == 0) return DSRT_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_trace_scratch(ctx, (size_t)n);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->t_rays, rays, sizeof(float) * 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    rc = dsrt_trace_closest_device(ctx, ctx->t_rays, n, ctx->t_hits);
    if (rc) return rc;
    return download_hits(ctx, n, tri, t, u, v);
}

extern "C" int dsrt_trace_occluded(dsrt_ctx* ctx, const float* rays, int n, uint8_t* occluded) {
    if (!ctx || n < 0 || (n > 0 && (!rays || !occluded))) return DSRT_EINVAL;
    if (!ctx->have_scene) return fail(ctx, DSRT_ENOSCENE, "trace before dsrt_scene_upload");
    if (n == 0) return DSRT_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_trace_scratch(ctx, (size_t)n);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->t_rays, rays, sizeof(float) * 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    rc = dsrt_trace_occluded_device(ctx, ctx->t_rays, n, ctx->t_occ);
    if (rc) return rc;
    CK(cudaMemcpyAsync(occluded, ctx->t_occ, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DSRT_OK;
}

extern "C" int dsrt_first_hits(dsrt_ctx* ctx, const DsrtCamera* camera, int width, int height,
                               float jx, float jy, int32_t* tri, float* t, float* u, float* v) {
    if (!ctx || !camera || width < 2 || height < 2) return DSRT_EINVAL;
    if (!ctx->have_scene) return fail(ctx, DSRT_ENOSCENE, "dsrt_first_hits before dsrt_scene_upload");
    CK(cudaSetDevice(ctx->device));
    const int n = width * height;
    int rc = ensure_trace_scratch(ctx, (size_t)n);
    if (rc) return rc;
    FrameConst fc;
    memset(&fc, 0, sizeof(fc));
    fill_camera(camera, fc);
    fc.W = width; fc.H = height;
    launch_camera_rays(fc, jx, jy, ctx->t_rays, ctx->stream);
    CK(cudaGetLastError());
    rc = dsrt_trace_closest_device(ctx, ctx->t_rays, n, ctx->t_hits);
    if (rc) return rc;
    return download_hits(ctx, n, tri, t, u, v);
}

// The first hits of ONE sample of every pixel, taken from the pipeline's own first stage: k_tile_entries (per-tile
// start nodes, culled tiles) and k_primary (ray generation with the estimator's jitter, traversal from the start
// nodes), exactly as dsrt_render_frame launches them -- unlike dsrt_first_hits, which traces fixed-jitter rays from the
// root through the generic kernel.  Also reports where each hit was found: the BVH8 node whose leaf child held the
// triangle and the triangle's position in the leaf-ordered record array.
extern "C" int dsrt_primary_hits(dsrt_ctx* ctx, const DsrtCamera* camera, const DsrtRenderParams* params, uint64_t seed,
                                 int sample, int32_t* tri, float* t, float* u, float* v,
                                 int32_t* leaf_node, int32_t* leaf_record) {
    if (!ctx) return DSRT_EINVAL;
    FrameArgs a;
    const float zero3[3] = {0.f, 0.f, 0.f};
    int rc = check_frame_args(ctx, camera, zero3, zero3, 0, params, sample, sample + 1, nullptr, &a);
    if (rc) return rc;
    if (a.rng_mode == DSRT_RNG_REF_LCG && sample != 0)
        return fail(ctx, DSRT_EINVAL, "dsrt_primary_hits: the LCG replay can only show sample 0 (later ones depend on the paths before)");
    CK(cudaSetDevice(ctx->device));
    FrameSlot& f = ctx->slots[0];
    if (f.pending) { rc = finish_frame(ctx, 0); if (rc) return rc; }
    bool fresh = false;
    rc = ensure_image(ctx, f, a.W, a.H, &fresh);
    if (rc) return rc;
    f.lcg_next_sample = -1;
    FrameConst fc;
    fill_frame_consts(ctx, a, camera, zero3, zero3, params, seed, fc);
    const size_t paths = (size_t)fc.slots_per_sample;
    rc = ensure_work(ctx, f, paths);
    if (rc) return rc;
    rc = ensure_tiles(ctx, f, fc);
    if (rc) return rc;
    if (paths > f.cap_leaf) {
        cudaFree(f.dbg_leaf); f.dbg_leaf = nullptr; f.cap_leaf = 0;
        CK(cudaMalloc(&f.dbg_leaf, paths * sizeof(uint2)));
        f.cap_leaf = paths;
    }
    bind_slot(f, fc);
    if (!want_tile_lists(ctx, fc)) fc.n_list = nullptr;
    fc.dbg_leaf = f.dbg_leaf;
    fc.s0 = sample;
    fc.spb = 1;
    if (a.rng_mode == DSRT_RNG_REF_LCG) launch_lcg_seed(f.rng_pix, a.W, a.H, (uint32_t)(seed & 0xFFFFFFFFu), f.stream);
    CK(cudaMemsetAsync(f.stats, 0, sizeof(unsigned long long) * S_COUNT, f.stream));
    launch_tile_entries(fc, f.tile_table, f.active_tiles, f.stream);
    launch_batch_begin(fc, f.stream);
    launch_primary(fc, ctx->dims, true, f.stream);
    launch_primary_lists(fc, ctx->dims, true, f.stream);
    CK(cudaGetLastError());
    uint32_t hc[C_COUNT];
    CK(cudaMemcpyAsync(hc, f.counters, sizeof(hc), cudaMemcpyDeviceToHost, f.stream));
    CK(cudaStreamSynchronize(f.stream));
    const size_t n = hc[C_NIN];
    if (n > paths) return fail(ctx, DSRT_ECUDA, "dsrt_primary_hits: %zu hits for %zu paths", n, paths);
    std::vector<float4> hh(n);
    std::vector<uint32_t> hp(n);
    std::vector<uint2> hl(n);
    if (n) {
        CK(cudaMemcpyAsync(hh.data(), f.hits, n * sizeof(float4), cudaMemcpyDeviceToHost, f.stream));
        CK(cudaMemcpyAsync(hp.data(), f.pq, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, f.stream));
        CK(cudaMemcpyAsync(hl.data(), f.dbg_leaf, n * sizeof(uint2), cudaMemcpyDeviceToHost, f.stream));
        CK(cudaStreamSynchronize(f.stream));
    }
    const size_t px = (size_t)a.W * a.H;
    for (size_t i = 0; i < px; ++i) {
        if (tri) tri[i] = -1;
        if (t) t[i] = 1e9f;
        if (u) u[i] = 0.f;
        if (v) v[i] = 0.f;
        if (leaf_node) leaf_node[i] = -1;
        if (leaf_record) leaf_record[i] = -1;
    }
    for (size_t i = 0; i < n; ++i) {
        // one sample in flight: path = slot = tile * 32 + position in the 8x4 tile (render.cu: slot_to_pixel)
        const uint32_t slot = hp[i], tile = slot >> 5, in = slot & 31u;
        const int x = (int)((tile % (uint32_t)fc.tiles_x) * 8u + (in & 7u)), y = (int)((tile / (uint32_t)fc.tiles_x) * 4u + (in >> 3));
        if (x >= a.W || y >= a.H) return fail(ctx, DSRT_ECUDA, "dsrt_primary_hits: a hit outside the image");
        const size_t o = (size_t)y * a.W + x;
        int32_t id;
        memcpy(&id, &hh[i].w, 4);
        if (tri) tri[o] = id;
        if (t) t[o] = hh[i].x;
        if (u) u[o] = hh[i].y;
        if (v) v[o] = hh[i].z;
        if (leaf_node) leaf_node[o] = (int32_t)hl[i].x;
        if (leaf_record) leaf_record[o] = (int32_t)hl[i].y;
    }
    return DSRT_OK;
}
