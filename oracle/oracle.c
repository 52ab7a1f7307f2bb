/* oracle.c -- CPU restatement of the reference's render path.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h for who may use it and for the parity-pinning status).
 *
 * Each function names the reference lines it follows (paths under /root/reference).  Float
 * expressions use fmaf exactly where nvcc 12.9 contracts the reference's source for sm_100a
 * (re-derived from cuobjdump -sass): built with -ffp-contract=off so
 * nothing else is fused.
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ---- minimal parallel-for over rows / ray chunks (pthreads; OpenMP is not in this image) ---- */
typedef void (*or_job_fn)(int index, void* arg);
typedef struct { or_job_fn fn; void* arg; int n; int next; pthread_mutex_t mu; } OrPool;

static void* or_worker(void* p) {
    OrPool* pool = (OrPool*)p;
    for (;;) {
        pthread_mutex_lock(&pool->mu);
        int i = pool->next++;
        pthread_mutex_unlock(&pool->mu);
        if (i >= pool->n) break;
        pool->fn(i, pool->arg);
    }
    return NULL;
}

static int or_threads = 0;
void or_set_threads(int n) { or_threads = n; }
int or_get_threads(void) {
    if (or_threads > 0) return or_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void or_parallel_for(int n, or_job_fn fn, void* arg) {
    int nt = or_get_threads();
    if (nt > n) nt = n;
    if (nt <= 1) { for (int i = 0; i < n; ++i) fn(i, arg); return; }
    OrPool pool = {fn, arg, n, 0, PTHREAD_MUTEX_INITIALIZER};
    pthread_t th[256];
    if (nt > 256) nt = 256;
    for (int t = 0; t < nt; ++t) pthread_create(&th[t], NULL, or_worker, &pool);
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
}

typedef struct { float x, y, z; } f3;

static inline f3 F3(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 f3_sub(f3 a, f3 b) { return F3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 f3_mul(f3 a, f3 b) { return F3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3 f3_scale(f3 a, float s) { return F3(a.x * s, a.y * s, a.z * s); }
static inline f3 f3_neg(f3 a) { return F3(-a.x, -a.y, -a.z); }
/* gpu_render.cu:31-33 as compiled: fma(a.z,b.z, fma(a.x,b.x, rn(a.y*b.y))) */
static inline float f3_dot(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.x, b.x, a.y * b.y)); }
/* gpu_render.cu:35-41 as compiled: p*q - r*s -> fma(p,q,-rn(r*s)) */
static inline f3 f3_cross(f3 a, f3 b) {
    return F3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
/* a + b*s as compiled */
static inline f3 f3_madd(f3 b, float s, f3 a) { return F3(fmaf(b.x, s, a.x), fmaf(b.y, s, a.y), fmaf(b.z, s, a.z)); }
/* gpu_render.cu:51-56 */
static inline f3 f3_norm(f3 a) {
    float L = sqrtf(f3_dot(a, a));
    if (L <= 0.0f) return F3(0.f, 0.f, 0.f);
    float inv = 1.0f / L;
    return F3(a.x * inv, a.y * inv, a.z * inv);
}
static inline float clamp01(float x) { return fminf(1.0f, fmaxf(0.0f, x)); }

/* ---- layouts shared with the reference ------------------------------------------------- */
typedef struct {           /* GPUMaterial, gpu_scene.h:34-45 */
    int type, albedo_tex, pad0, pad1;
    float albedo[3], emissive[3];
    float fuzz, ref_idx;
} OrMaterial;

typedef struct {           /* GPUBVHNode semantics, gpu_scene.h:77-85 */
    float bmin[3], bmax[3];
    int left, right, tri_offset, tri_count;
} OrNode;

typedef struct { float center[3], radius; int material_id, pad; } OrSphere;      /* GPUSphere, gpu_scene.h:50-55 */
typedef struct { int width, height, offset; } OrTexHeader;                       /* GPUTextureHeader, gpu_scene.h:28-32 */

struct OrScene {
    int n, nmats;
    float* tris;           /* n x 29 */
    OrMaterial* mats;
    int* indices;
    OrNode* nodes;
    int num_nodes, cap_nodes;
    int max_leaf;
    OrSphere* spheres; int num_spheres;
    OrTexHeader* textures; int num_textures;
    float* texture_pool; int texture_pool_floats;
};

#define MAT_LAMBERTIAN 0
#define MAT_METAL 1
#define MAT_DIELECTRIC 2
#define MAT_DIFFUSE_LIGHT 3

/* ---- BVH build: gpu_scene_builder.cpp:144-178, 343-459 ------------------------------------ */
static inline const float* tri_ptr(const OrScene* s, int i) { return s->tris + (size_t)i * 29; }

static void tri_bounds(const float* t, float lo[3], float hi[3]) {
    for (int a = 0; a < 3; ++a) {
        lo[a] = fminf(fminf(t[a], t[3 + a]), t[6 + a]);
        hi[a] = fmaxf(fmaxf(t[a], t[3 + a]), t[6 + a]);
    }
}
static inline float tri_centroid(const float* t, int a) { return (t[a] + t[3 + a] + t[6 + a]) / 3.f; }

/* std::nth_element semantics: afterwards idx[mid] is the element a full sort would put
 * there, everything before is not greater, everything after not smaller.  (The arrangement
 * inside the halves is implementation-defined in the reference too -- Appendix A-13.) */
static void nth_element_axis(const OrScene* s, int* idx, int lo, int hi, int mid, int axis) {
    while (hi - lo > 1) {
        int m = lo + (hi - lo) / 2;
        float pv = tri_centroid(tri_ptr(s, idx[m]), axis);
        int i = lo, j = hi - 1;
        while (i <= j) {
            while (tri_centroid(tri_ptr(s, idx[i]), axis) < pv) ++i;
            while (tri_centroid(tri_ptr(s, idx[j]), axis) > pv) --j;
            if (i <= j) { int t = idx[i]; idx[i] = idx[j]; idx[j] = t; ++i; --j; }
        }
        if (mid <= j) hi = j + 1;
        else if (mid >= i) lo = i;
        else return;
    }
}

static int build_recursive(OrScene* s, int start, int end) {
    const int node_index = s->num_nodes++;
    OrNode* node = &s->nodes[node_index];
    float lo[3], hi[3];
    tri_bounds(tri_ptr(s, s->indices[start]), lo, hi);
    for (int i = start + 1; i < end; ++i) {
        float l2[3], h2[3];
        tri_bounds(tri_ptr(s, s->indices[i]), l2, h2);
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], l2[a]); hi[a] = fmaxf(hi[a], h2[a]); }
    }
    for (int a = 0; a < 3; ++a) { node->bmin[a] = lo[a]; node->bmax[a] = hi[a]; }
    node->left = node->right = -1;
    const int prim_count = end - start;
    node->tri_offset = start;
    node->tri_count = prim_count;
    if (prim_count <= 4) {                                   /* max_leaf_size, :371 */
        if (prim_count > s->max_leaf) s->max_leaf = prim_count;
        return node_index;
    }
    float cmin[3], cmax[3];
    for (int a = 0; a < 3; ++a) cmin[a] = cmax[a] = tri_centroid(tri_ptr(s, s->indices[start]), a);
    for (int i = start + 1; i < end; ++i)
        for (int a = 0; a < 3; ++a) {
            float c = tri_centroid(tri_ptr(s, s->indices[i]), a);
            cmin[a] = fminf(cmin[a], c);
            cmax[a] = fmaxf(cmax[a], c);
        }
    const float dx = cmax[0] - cmin[0], dy = cmax[1] - cmin[1], dz = cmax[2] - cmin[2];
    int axis = 0;                                            /* :403-406, asymmetric tie-breaks */
    if (dy > dx && dy >= dz) axis = 1;
    else if (dz > dx && dz >= dy) axis = 2;
    if ((axis == 0 && dx == 0.0f) || (axis == 1 && dy == 0.0f) || (axis == 2 && dz == 0.0f)) {
        if (prim_count > s->max_leaf) s->max_leaf = prim_count;
        return node_index;
    }
    const int mid = (start + end) / 2;
    nth_element_axis(s, s->indices, start, end, mid, axis);
    node->tri_offset = 0;
    node->tri_count = 0;
    const int l = build_recursive(s, start, mid);
    const int r = build_recursive(s, mid, end);
    s->nodes[node_index].left = l;
    s->nodes[node_index].right = r;
    return node_index;
}

OrScene* or_scene_create(const float* tris, int n, const void* mats, int nmats) {
    OrScene* s = (OrScene*)calloc(1, sizeof(OrScene));
    s->n = n; s->nmats = nmats;
    s->tris = (float*)malloc(sizeof(float) * 29 * (size_t)(n > 0 ? n : 1));
    memcpy(s->tris, tris, sizeof(float) * 29 * (size_t)n);
    s->mats = (OrMaterial*)malloc(sizeof(OrMaterial) * (size_t)(nmats > 0 ? nmats : 1));
    memcpy(s->mats, mats, sizeof(OrMaterial) * (size_t)nmats);
    s->indices = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) s->indices[i] = i;
    s->cap_nodes = 2 * (n > 0 ? n : 1);
    s->nodes = (OrNode*)malloc(sizeof(OrNode) * (size_t)s->cap_nodes);
    s->num_nodes = 0;
    if (n > 0) build_recursive(s, 0, n);
    return s;
}

/* GPUScene.spheres / .textures / .texture_pool (gpu_scene.h:116-173) */
void or_scene_set_spheres(OrScene* s, const void* spheres, int n) {
    free(s->spheres);
    s->spheres = (OrSphere*)malloc(sizeof(OrSphere) * (size_t)(n > 0 ? n : 1));
    memcpy(s->spheres, spheres, sizeof(OrSphere) * (size_t)n);
    s->num_spheres = n;
}
void or_scene_set_textures(OrScene* s, const void* headers, int n, const float* pool, int pool_floats) {
    free(s->textures); free(s->texture_pool);
    s->textures = (OrTexHeader*)malloc(sizeof(OrTexHeader) * (size_t)(n > 0 ? n : 1));
    memcpy(s->textures, headers, sizeof(OrTexHeader) * (size_t)n);
    s->texture_pool = (float*)malloc(sizeof(float) * (size_t)(pool_floats > 0 ? pool_floats : 1));
    memcpy(s->texture_pool, pool, sizeof(float) * (size_t)pool_floats);
    s->num_textures = n; s->texture_pool_floats = pool_floats;
}

void or_scene_destroy(OrScene* s) {
    if (!s) return;
    free(s->tris); free(s->mats); free(s->indices); free(s->nodes);
    free(s->spheres); free(s->textures); free(s->texture_pool); free(s);
}
int or_scene_num_nodes(const OrScene* s) { return s->num_nodes; }
int or_scene_max_leaf(const OrScene* s) { return s->max_leaf; }

/* ---- traversal ---------------------------------------------------------------------------- */
typedef struct { f3 orig, dir; } Ray;

typedef struct {
    float t; f3 p; f3 normal; int mat_id, tri_tex_id, tri_index; float u, v; int front_face;
    int sphere_index;      /* not in the reference's HitRecord: which sphere, for the hooks (-1 = triangle) */
} Hit;

/* gpu_render.cu:285-315.  Note `t_max <= t_min`: a box of zero thickness is never hit. */
static int bbox_hit(const OrNode* node, const Ray* r, float t_min, float t_max) {
    const float o[3] = {r->orig.x, r->orig.y, r->orig.z}, d[3] = {r->dir.x, r->dir.y, r->dir.z};
    for (int a = 0; a < 3; ++a) {
        float invD = 1.0f / d[a];
        float t0 = (node->bmin[a] - o[a]) * invD;
        float t1 = (node->bmax[a] - o[a]) * invD;
        if (invD < 0.0f) { float tmp = t0; t0 = t1; t1 = tmp; }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min) return 0;
    }
    return 1;
}

/* gpu_render.cu:322-380 */
static int hit_triangle_index(const OrScene* s, int tri_index, const Ray* ray, float t_min, float t_max, Hit* rec) {
    const float* T = tri_ptr(s, tri_index);
    const f3 v0 = F3(T[0], T[1], T[2]), v1 = F3(T[3], T[4], T[5]), v2 = F3(T[6], T[7], T[8]);
    const f3 edge1 = f3_sub(v1, v0), edge2 = f3_sub(v2, v0);
    const f3 pvec = f3_cross(ray->dir, edge2);
    const float det = f3_dot(edge1, pvec);
    if (fabsf(det) < 1e-8f) return 0;
    const float invDet = 1.0f / det;
    const f3 tvec = f3_sub(ray->orig, v0);
    const float u = f3_dot(tvec, pvec) * invDet;
    if (u < 0.0f || u > 1.0f) return 0;
    const f3 qvec = f3_cross(tvec, edge1);
    const float v = f3_dot(ray->dir, qvec) * invDet;
    if (v < 0.0f || u + v > 1.0f) return 0;
    const float t = f3_dot(edge2, qvec) * invDet;
    if (t < t_min || t > t_max) return 0;
    rec->t = t;
    rec->p = f3_madd(ray->dir, t, ray->orig);
    const float w = 1.0f - u - v;
    const f3 n0 = F3(T[9], T[10], T[11]), n1 = F3(T[12], T[13], T[14]), n2 = F3(T[15], T[16], T[17]);
    /* w*n0 + u*n1 + v*n2 as compiled: fma(v,n2, fma(w,n0, rn(u*n1))) */
    f3 nrm = F3(fmaf(v, n2.x, fmaf(w, n0.x, u * n1.x)),
                fmaf(v, n2.y, fmaf(w, n0.y, u * n1.y)),
                fmaf(v, n2.z, fmaf(w, n0.z, u * n1.z)));
    nrm = f3_norm(nrm);
    rec->front_face = f3_dot(ray->dir, nrm) < 0.0f;
    rec->normal = rec->front_face ? nrm : f3_neg(nrm);
    rec->u = u; rec->v = v;
    memcpy(&rec->mat_id, &T[27], 4);
    memcpy(&rec->tri_tex_id, &T[28], 4);
    rec->tri_index = tri_index;
    rec->sphere_index = -1;
    return 1;
}

/* gpu_render.cu:387-473 */
static int bvh_hit_closest(const OrScene* s, const Ray* ray, float t_min, float t_max, Hit* out_rec) {
    if (s->num_nodes <= 0) return 0;
    int stack[64];
    int stack_size = 0, node_index = 0, hit_anything = 0;
    float closest = t_max;
    Hit tmp;
    for (;;) {
        const OrNode* node = &s->nodes[node_index];
        if (bbox_hit(node, ray, t_min, closest)) {
            if (node->tri_count > 0) {
                for (int i = 0; i < node->tri_count; ++i) {
                    int tri_idx = s->indices[node->tri_offset + i];
                    if (hit_triangle_index(s, tri_idx, ray, t_min, closest, &tmp)) {
                        hit_anything = 1;
                        closest = tmp.t;
                        *out_rec = tmp;
                    }
                }
                if (stack_size == 0) break;
                node_index = stack[--stack_size];
            } else {
                const OrNode* left = &s->nodes[node->left];
                const OrNode* right = &s->nodes[node->right];
                const int hit_left = bbox_hit(left, ray, t_min, closest);
                const int hit_right = bbox_hit(right, ray, t_min, closest);
                if (hit_left && hit_right) {
                    /* 0.5f*(min+max) - orig contracts to fma(sum, 0.5f, -orig) */
                    f3 dl = F3(fmaf(left->bmin[0] + left->bmax[0], 0.5f, -ray->orig.x),
                               fmaf(left->bmin[1] + left->bmax[1], 0.5f, -ray->orig.y),
                               fmaf(left->bmin[2] + left->bmax[2], 0.5f, -ray->orig.z));
                    f3 dr = F3(fmaf(right->bmin[0] + right->bmax[0], 0.5f, -ray->orig.x),
                               fmaf(right->bmin[1] + right->bmax[1], 0.5f, -ray->orig.y),
                               fmaf(right->bmin[2] + right->bmax[2], 0.5f, -ray->orig.z));
                    const float dL = f3_dot(dl, ray->dir), dR = f3_dot(dr, ray->dir);
                    const int near_idx = (dL < dR) ? node->left : node->right;
                    const int far_idx = (dL < dR) ? node->right : node->left;
                    stack[stack_size++] = far_idx;
                    node_index = near_idx;
                } else if (hit_left) {
                    node_index = node->left;
                } else if (hit_right) {
                    node_index = node->right;
                } else {
                    if (stack_size == 0) break;
                    node_index = stack[--stack_size];
                }
            }
        } else {
            if (stack_size == 0) break;
            node_index = stack[--stack_size];
        }
    }
    return hit_anything;
}

/* gpu_render.cu:478-504 as compiled for sm_100a (cuobjdump of hit_sphere):
 * c = fma(-r, r, dot(oc, oc)), discriminant = fma(hb, hb, -rn(a * c)), IEEE sqrt and divisions,
 * p = fma(dir, t, orig), normal = (p - centre) * (1 / r). */
/* x_first: dot(dir, dir) as render_kernel's closest-hit loop has it after hoisting,
 * fma(z,z, fma(y,y, rn(x*x))); elsewhere (shadow loop, stand-alone scene_hit) the usual f3_dot. */
static int hit_sphere(const OrSphere* sph, const Ray* ray, float t_min, float t_max, int x_first, float* t_out, f3* normal_out) {
    const f3 centre = F3(sph->center[0], sph->center[1], sph->center[2]);
    const f3 oc = f3_sub(ray->orig, centre);
    const f3 d = ray->dir;
    const float a = x_first ? fmaf(d.z, d.z, fmaf(d.y, d.y, d.x * d.x)) : f3_dot(d, d);
    const float half_b = f3_dot(oc, ray->dir);
    const float c = fmaf(-sph->radius, sph->radius, f3_dot(oc, oc));
    const float disc = fmaf(half_b, half_b, -(a * c));
    if (disc < 0.0f) return 0;
    const float sqrtd = sqrtf(disc);
    float root = (-half_b - sqrtd) / a;
    if (root < t_min || root > t_max) {
        root = (-half_b + sqrtd) / a;
        if (root < t_min || root > t_max) return 0;
    }
    *t_out = root;
    const f3 p = f3_madd(ray->dir, root, ray->orig);
    *normal_out = f3_scale(f3_sub(p, centre), 1.0f / sph->radius);
    return 1;
}

/* gpu_render.cu:509-551: BVH triangles, then the brute-force sphere loop on [t_min, closest] */
static int scene_hit_ctx(const OrScene* s, const Ray* ray, float t_min, float t_max, int x_first, Hit* rec) {
    Hit best;
    memset(&best, 0, sizeof(best));
    int hit_any = 0;
    float closest = t_max;
    Hit tri_rec;
    memset(&tri_rec, 0, sizeof(tri_rec));
    if (bvh_hit_closest(s, ray, t_min, closest, &tri_rec)) { hit_any = 1; closest = tri_rec.t; best = tri_rec; }
    for (int i = 0; i < s->num_spheres; ++i) {
        float t_hit; f3 n_hit;
        if (hit_sphere(&s->spheres[i], ray, t_min, closest, x_first, &t_hit, &n_hit)) {
            hit_any = 1;
            closest = t_hit;
            best.t = t_hit;
            best.p = f3_madd(ray->dir, t_hit, ray->orig);
            best.front_face = f3_dot(ray->dir, n_hit) < 0.0f;             /* set_face_normal, :275-278 */
            best.normal = best.front_face ? n_hit : f3_neg(n_hit);
            best.mat_id = s->spheres[i].material_id;
            best.tri_tex_id = -1;
            best.tri_index = -1;
            best.sphere_index = i;
            best.u = 0.0f; best.v = 0.0f;
        }
    }
    if (hit_any) *rec = best;
    return hit_any;
}

static int scene_hit(const OrScene* s, const Ray* ray, float t_min, float t_max, Hit* rec) {
    return scene_hit_ctx(s, ray, t_min, t_max, 0, rec);
}

/* tex2D, gpu_render.cu:232-259: nearest texel, wrap by x - floor(x), white when out of range */
static f3 tex2D(const OrScene* s, int tex_id, float u, float v) {
    if (tex_id < 0 || tex_id >= s->num_textures || !s->textures || !s->texture_pool) return F3(1.f, 1.f, 1.f);
    const OrTexHeader th = s->textures[tex_id];
    u = u - floorf(u);
    v = v - floorf(v);
    const int i = (int)(u * (float)(th.width - 1));
    const int j = (int)((1.0f - v) * (float)(th.height - 1));
    const int idx = th.offset + (j * th.width + i) * 3;
    if (idx < 0 || idx + 2 >= s->texture_pool_floats) return F3(1.f, 1.f, 1.f);
    return F3(s->texture_pool[idx], s->texture_pool[idx + 1], s->texture_pool[idx + 2]);
}

/* ---- random numbers ----------------------------------------------------------------------- */
static void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
void or_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    memcpy(out, c, sizeof(c));
}

typedef struct { int mode; uint32_t state, pixel, sample, k0, k1; } Rng;

static float rand01(Rng* r) {
    uint32_t bits24;
    if (r->mode == 1) {                                      /* gpu_render.cu:77-80 */
        r->state = r->state * 1664525u + 1013904223u;
        bits24 = r->state & 0x00FFFFFFu;
    } else {                                                 /* product stream, render.cu */
        uint32_t ctr[4] = {r->pixel, r->sample, r->state >> 2, 0u}, key[2] = {r->k0, r->k1}, o[4];
        or_philox4x32_10(ctr, key, o);
        bits24 = o[r->state & 3u] >> 8;
        r->state += 1u;
    }
    return (float)bits24 * 5.9604644775390625e-08f;           /* / 16777216.0f */
}

/* ---- estimator ----------------------------------------------------------------------------- */
static const float PI_F = 3.14159265358979323846f;           /* gpu_render.cu:96 */

static f3 random_in_unit_sphere(Rng* rng) {                   /* gpu_render.cu:82-91 */
    for (;;) {
        float x = fmaf(rand01(rng), 2.0f, -1.0f);
        float y = fmaf(rand01(rng), 2.0f, -1.0f);
        float z = fmaf(rand01(rng), 2.0f, -1.0f);
        f3 p = F3(x, y, z);
        if (f3_dot(p, p) >= 1.0f) continue;
        return p;
    }
}

static f3 reflect_ref(f3 v, f3 n) {                           /* gpu_render.cu:195-197 */
    float d = f3_dot(v, n);
    float two_d = d + d;
    return F3(fmaf(-n.x, two_d, v.x), fmaf(-n.y, two_d, v.y), fmaf(-n.z, two_d, v.z));
}

/* sample_cosine_hemisphere + build_onb + random_cosine_direction, gpu_render.cu:99-141 */
static f3 sample_cosine_hemisphere(f3 normal, Rng* rng, float* pdf_out) {
    const f3 w = f3_norm(normal);
    const f3 a = fabsf(w.x) > 0.9f ? F3(0.f, 1.f, 0.f) : F3(1.f, 0.f, 0.f);
    const f3 vv = f3_norm(f3_cross(w, a));
    const f3 uu = f3_cross(vv, w);
    const float r1 = rand01(rng), r2 = rand01(rng);
    const float lz = sqrtf(1.0f - r2);
    const float phi = (PI_F + PI_F) * r1;
    const float sr2 = sqrtf(r2);
    const float lx = cosf(phi) * sr2, ly = sinf(phi) * sr2;
    f3 wd = F3(fmaf(w.x, lz, fmaf(uu.x, lx, vv.x * ly)),
               fmaf(w.y, lz, fmaf(uu.y, lx, vv.y * ly)),
               fmaf(w.z, lz, fmaf(uu.z, lx, vv.z * ly)));
    wd = f3_norm(wd);
    const float cos_b = fmaxf(0.0f, f3_dot(wd, normal));
    *pdf_out = cos_b > 0.0f ? cos_b / PI_F : 0.0f;
    return wd;
}

/* sample_sphere_light_direction, gpu_render.cu:145-188 */
static f3 sample_sphere_light_direction(const OrSphere* sph, f3 origin, Rng* rng, float* pdf_out) {
    const float z = fmaf(2.0f, rand01(rng), -1.0f);
    const float phi = (PI_F + PI_F) * rand01(rng);
    const float r = sqrtf(fmaxf(0.0f, fmaf(-z, z, 1.0f)));
    const f3 local = F3(r * cosf(phi), r * sinf(phi), z);
    const f3 centre = F3(sph->center[0], sph->center[1], sph->center[2]);
    const f3 p_light = f3_madd(local, sph->radius, centre);
    const f3 to_light = f3_sub(p_light, origin);
    const float dist2 = f3_dot(to_light, to_light);
    const float dist = sqrtf(dist2);
    if (dist <= 0.0f) { *pdf_out = 0.0f; return F3(0.f, 0.f, 1.f); }
    const f3 wi = f3_scale(to_light, 1.0f / dist);
    const f3 n_light = f3_norm(f3_sub(p_light, centre));
    const float cos_l = fmaxf(0.0f, f3_dot(n_light, f3_scale(wi, -1.0f)));
    if (cos_l <= 0.0f) { *pdf_out = 0.0f; return wi; }
    const float area = ((4.0f * PI_F) * sph->radius) * sph->radius;
    *pdf_out = dist2 / (cos_l * area);
    return wi;
}

typedef struct { uint64_t primary, bounce, shadow; } Counts;

/* ray_color, gpu_render.cu:715-936 */
static f3 ray_color(const OrScene* s, Ray ray, Rng* rng, int max_depth_param, int sun_enabled,
                    f3 sun_dir, f3 sun_rad, Counts* cnt) {
    f3 L = F3(0, 0, 0), T = F3(1, 1, 1);
    const int max_depth = max_depth_param > 0 ? max_depth_param : 12;
    for (int depth = 0; depth < max_depth; ++depth) {
        if (depth >= 5) {
            float p = fmaxf(T.x, fmaxf(T.y, T.z));
            p = fminf(p, 0.95f);
            if (rand01(rng) > p) break;
            T = f3_scale(T, 1.0f / p);
        }
        Hit rec;
        if (depth == 0) cnt->primary++; else cnt->bounce++;
        if (!scene_hit_ctx(s, &ray, 0.001f, 1e9f, 1, &rec)) break;      /* render_kernel's closest-hit context */
        const OrMaterial* mat = &s->mats[rec.mat_id];
        if (mat->type == MAT_DIFFUSE_LIGHT) {
            L = F3(fmaf(T.x, mat->emissive[0], L.x), fmaf(T.y, mat->emissive[1], L.y), fmaf(T.z, mat->emissive[2], L.z));
            break;
        }
        f3 albedo = F3(mat->albedo[0], mat->albedo[1], mat->albedo[2]);
        if (rec.tri_tex_id >= 0) {                             /* gpu_render.cu:765-774 */
            const float* T9 = tri_ptr(s, rec.tri_index);
            const float w = 1.0f - rec.u - rec.v;
            /* w*uv0 + u*uv1 + v*uv2 as compiled: fma(v,uv2, fma(w,uv0, rn(u*uv1))) */
            const float u_tex = fmaf(rec.v, T9[24], fmaf(w, T9[18], rec.u * T9[21]));
            const float v_tex = fmaf(rec.v, T9[25], fmaf(w, T9[19], rec.u * T9[22]));
            albedo = f3_mul(albedo, tex2D(s, rec.tri_tex_id, u_tex, v_tex));
        }
        if (mat->type == MAT_DIELECTRIC) {                     /* gpu_render.cu:621-661 */
            float eta = mat->ref_idx;
            if (eta <= 0.0f || !isfinite(eta)) eta = 1.5f;
            const float ratio = rec.front_face ? (1.0f / eta) : eta;
            const f3 ud = f3_norm(ray.dir);
            const float cos_t = fminf(-f3_dot(ud, rec.normal), 1.0f);
            const float sin_t = sqrtf(fmaxf(0.0f, fmaf(-cos_t, cos_t, 1.0f)));
            const int cannot = ratio * sin_t > 1.0f;
            float r0 = (1.0f - ratio) / (1.0f + ratio);
            r0 = r0 * r0;
            const float refl = fmaf(1.0f - r0, powf(1.0f - cos_t, 5.0f), r0);
            f3 nd;
            if (cannot || refl > rand01(rng)) {
                nd = reflect_ref(ud, rec.normal);
            } else {                                           /* refract, gpu_render.cu:199-206 */
                const f3 uv = f3_norm(ud);
                const float c2 = fminf(-f3_dot(uv, rec.normal), 1.0f);
                const f3 perp = f3_scale(f3_madd(rec.normal, c2, uv), ratio);
                const float k = -sqrtf(fabsf(1.0f - f3_dot(perp, perp)));
                nd = f3_madd(rec.normal, k, perp);
            }
            ray.orig = rec.p; ray.dir = nd;
            continue;
        }
        if (mat->type == MAT_METAL) {                          /* gpu_render.cu:603-619 */
            const f3 refl = reflect_ref(f3_norm(ray.dir), rec.normal);
            const float fuzz = fmaxf(0.0f, fminf(1.0f, mat->fuzz));
            const f3 sph = random_in_unit_sphere(rng);
            const f3 nd = f3_madd(sph, fuzz, refl);
            if (!(f3_dot(nd, rec.normal) > 0.0f)) break;
            T = f3_mul(T, albedo);
            ray.orig = rec.p; ray.dir = nd;
            continue;
        }
        if (sun_enabled) {                                     /* gpu_render.cu:800-836 */
            const f3 Ldir = f3_norm(f3_neg(sun_dir));
            const float cos_l = fmaxf(0.0f, f3_dot(rec.normal, Ldir));
            if (cos_l > 0.0f) {
                Ray sh;
                sh.orig = f3_madd(rec.normal, 1e-3f, rec.p);
                sh.dir = Ldir;
                Hit srec;
                cnt->shadow++;
                if (!scene_hit(s, &sh, 0.001f, 1e9f, &srec)) {
                    const float pdf = cos_l / PI_F;
                    const float mix = fmaf(pdf, 0.5f, 0.5f);
                    const float wgt = pdf / mix;
                    const f3 y = F3(albedo.x * (sun_rad.x * wgt), albedo.y * (sun_rad.y * wgt), albedo.z * (sun_rad.z * wgt));
                    L = F3(fmaf(y.x, T.x, L.x), fmaf(y.y, T.y, L.y), fmaf(y.z, T.z, L.z));
                }
            }
        }
        /* gpu_render.cu:838-850: emissive spheres are re-counted at every diffuse hit */
        int num_lights = 0;
        for (int i = 0; i < s->num_spheres; ++i) {
            const OrMaterial* lm = &s->mats[s->spheres[i].material_id];
            if (lm->type == MAT_DIFFUSE_LIGHT && (lm->emissive[0] > 0 || lm->emissive[1] > 0 || lm->emissive[2] > 0)) num_lights++;
        }
        if (num_lights == 0) {                                 /* gpu_render.cu:852-866 */
            float pdf_b;
            const f3 wd = sample_cosine_hemisphere(rec.normal, rng, &pdf_b);
            if (pdf_b <= 0.0f) break;
            const float cos_b = fmaxf(0.0f, f3_dot(wd, rec.normal));
            T = f3_mul(T, f3_scale(albedo, (cos_b / PI_F) / pdf_b));
            ray.orig = rec.p; ray.dir = wd;
            continue;
        }
        {                                                      /* gpu_render.cu:868-932 */
            f3 dir;
            float pdf_val = 0.0f;
            const float choose = rand01(rng);
            if (choose < 0.5f) {
                int k = (int)(rand01(rng) * (float)num_lights);
                if (k >= num_lights) k = num_lights - 1;
                int found = 0, light_idx = -1;
                for (int i = 0; i < s->num_spheres; ++i) {
                    const OrMaterial* lm = &s->mats[s->spheres[i].material_id];
                    if (lm->type == MAT_DIFFUSE_LIGHT && (lm->emissive[0] > 0 || lm->emissive[1] > 0 || lm->emissive[2] > 0)) {
                        if (found == k) { light_idx = i; break; }
                        found++;
                    }
                }
                float pdf_cond = 0.0f;
                dir = sample_sphere_light_direction(&s->spheres[light_idx], rec.p, rng, &pdf_cond);
                if (pdf_cond <= 0.0f) break;
                const float cos_s = fmaxf(0.0f, f3_dot(dir, rec.normal));
                if (cos_s <= 0.0f) break;
                const float pdf_light = pdf_cond / (float)num_lights;
                const float pdf_brdf = cos_s / PI_F;
                pdf_val = fmaf(0.5f, pdf_light, 0.5f * pdf_brdf);
            } else {
                float pdf_b = 0.0f;
                dir = sample_cosine_hemisphere(rec.normal, rng, &pdf_b);
                if (pdf_b <= 0.0f) break;
                pdf_val = 0.5f * pdf_b;
            }
            const float cos_s = fmaxf(0.0f, f3_dot(dir, rec.normal));
            const float wgt = (cos_s / PI_F) / pdf_val;
            T = f3_mul(T, f3_scale(albedo, wgt));
            ray.orig = rec.p; ray.dir = dir;
        }
    }
    return F3(clamp01(L.x), clamp01(L.y), clamp01(L.z));
}

/* make_camera_ray_jittered, gpu_render.cu:941-968 as compiled */
static Ray camera_ray(const float c[12], int px, int py, int W, int H, float jx, float jy) {
    const float u = ((float)px + jx) / (float)(W - 1);
    const float v = ((float)py + jy) / (float)(H - 1);
    Ray r;
    r.orig = F3(c[0], c[1], c[2]);
    r.dir = F3(fmaf(v, c[9], fmaf(u, c[6], c[3])) - c[0],
               fmaf(v, c[10], fmaf(u, c[7], c[4])) - c[1],
               fmaf(v, c[11], fmaf(u, c[8], c[5])) - c[2]);
    return r;
}

typedef struct {
    const OrScene* s; const float* cam12; int W, H; float jx, jy;
    int32_t* tri; float *t, *u, *v;
} FirstHitJob;

static void first_hit_row(int y, void* arg) {
    FirstHitJob* J = (FirstHitJob*)arg;
    for (int x = 0; x < J->W; ++x) {
        Ray r = camera_ray(J->cam12, x, y, J->W, J->H, J->jx, J->jy);
        Hit rec;
        int hit = scene_hit(J->s, &r, 0.001f, 1e9f, &rec);
        int i = x + y * J->W;
        J->tri[i] = hit ? (rec.sphere_index >= 0 ? -2 - rec.sphere_index : rec.tri_index) : -1;
        J->t[i] = hit ? rec.t : 1e9f;
        J->u[i] = hit ? rec.u : 0.f;
        J->v[i] = hit ? rec.v : 0.f;
    }
}

void or_first_hits(const OrScene* s, const float cam12[12], int W, int H, float jx, float jy,
                   int32_t* tri, float* t, float* u, float* v) {
    FirstHitJob J = {s, cam12, W, H, jx, jy, tri, t, u, v};
    or_parallel_for(H, first_hit_row, &J);
}

/* First hit of ONE sample of every pixel with the jitter the estimator really uses
 * (gpu_render.cu:990-998): rng_mode 1 = the reference LCG, sample 0 only (state = pixel ^ seed32,
 * :990); rng_mode 0 = the product's Philox stream, any sample (counter (pixel, sample, 0, 0)).
 * The pipeline's own first-hit kernel is checked against this (tests/test_gpu_parity.py). */
typedef struct {
    const OrScene* s; const float* cam12; int W, H, spp, sample, rng_mode; uint64_t seed;
    int32_t* tri; float *t, *u, *v;
} PrimaryHitJob;

static void primary_hit_row(int y, void* arg) {
    PrimaryHitJob* J = (PrimaryHitJob*)arg;
    for (int x = 0; x < J->W; ++x) {
        Rng rng;
        memset(&rng, 0, sizeof(rng));
        rng.mode = J->rng_mode;
        rng.pixel = (uint32_t)(x + y * J->W);
        rng.sample = (uint32_t)J->sample;
        rng.k0 = (uint32_t)(J->seed & 0xFFFFFFFFu);
        rng.k1 = (uint32_t)(J->seed >> 32);
        if (J->rng_mode == 1) rng.state = rng.pixel ^ (uint32_t)(J->seed & 0xFFFFFFFFu);
        const float jx = ((float)J->sample + rand01(&rng)) / (float)J->spp;
        const float jy = ((float)J->sample + rand01(&rng)) / (float)J->spp;
        Ray r = camera_ray(J->cam12, x, y, J->W, J->H, jx, jy);
        Hit rec;
        int hit = scene_hit_ctx(J->s, &r, 0.001f, 1e9f, 1, &rec);
        int i = x + y * J->W;
        J->tri[i] = hit ? (rec.sphere_index >= 0 ? -2 - rec.sphere_index : rec.tri_index) : -1;
        J->t[i] = hit ? rec.t : 1e9f;
        J->u[i] = hit ? rec.u : 0.f;
        J->v[i] = hit ? rec.v : 0.f;
    }
}

void or_primary_hits(const OrScene* s, const float cam12[12], int W, int H, int spp_total, int sample,
                     uint64_t seed, int rng_mode, int32_t* tri, float* t, float* u, float* v) {
    PrimaryHitJob J = {s, cam12, W, H, spp_total < 1 ? 1 : spp_total, sample, rng_mode, seed, tri, t, u, v};
    or_parallel_for(H, primary_hit_row, &J);
}

#define OR_CHUNK 1024
typedef struct {
    const OrScene* s; const float* rays; int n;
    int32_t* tri; float *t, *u, *v; uint8_t* occ;
} TraceJob;

static void trace_chunk(int c, void* arg) {
    TraceJob* J = (TraceJob*)arg;
    int end = (c + 1) * OR_CHUNK < J->n ? (c + 1) * OR_CHUNK : J->n;
    for (int i = c * OR_CHUNK; i < end; ++i) {
        const float* q = J->rays + (size_t)i * 8;
        Ray r;
        r.orig = F3(q[0], q[1], q[2]);
        r.dir = F3(q[4], q[5], q[6]);
        Hit rec;
        int hit = scene_hit(J->s, &r, q[3], q[7], &rec);
        if (J->occ) { J->occ[i] = (uint8_t)hit; continue; }
        J->tri[i] = hit ? (rec.sphere_index >= 0 ? -2 - rec.sphere_index : rec.tri_index) : -1;
        J->t[i] = hit ? rec.t : q[7];
        J->u[i] = hit ? rec.u : 0.f;
        J->v[i] = hit ? rec.v : 0.f;
    }
}

void or_trace(const OrScene* s, const float* rays, int n, int32_t* tri, float* t, float* u, float* v) {
    TraceJob J = {s, rays, n, tri, t, u, v, NULL};
    or_parallel_for((n + OR_CHUNK - 1) / OR_CHUNK, trace_chunk, &J);
}

void or_occluded(const OrScene* s, const float* rays, int n, uint8_t* occluded) {
    TraceJob J = {s, rays, n, NULL, NULL, NULL, NULL, occluded};
    or_parallel_for((n + OR_CHUNK - 1) / OR_CHUNK, trace_chunk, &J);
}

/* gpu_render.cu:1003-1030 */
void or_tonemap(const float* accum, int W, int H, int spp_total, float gamma, unsigned char* rgb) {
    const float inv_spp = 1.0f / (float)(spp_total < 1 ? 1 : spp_total);
    const float inv_gamma = 1.0f / (gamma > 0.0f ? gamma : 1.0f);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < 3; ++c) {
                float v = accum[((size_t)y * W + x) * 3 + c] * inv_spp;
                v = fmaxf(v, 0.0f);
                v = fminf(v, 10.0f);
                v = powf(v, inv_gamma);
                v = clamp01(v);
                rgb[((size_t)(H - 1 - y) * W + x) * 3 + c] = (unsigned char)(255.99f * v);
            }
}

/* render_kernel, gpu_render.cu:973-1031 */
typedef struct {
    const OrScene* s; const float* cam12; int W, H, spp, max_depth, sun_enabled, rng_mode, s_begin, s_end;
    f3 sd, sr; uint64_t seed; float* accum; uint32_t* lcg_state;
    uint64_t counts[3]; pthread_mutex_t mu;
} RenderJob;

static void render_row(int y, void* arg) {
    RenderJob* J = (RenderJob*)arg;
    Counts cnt = {0, 0, 0};
    for (int x = 0; x < J->W; ++x) {
        Rng rng;
        memset(&rng, 0, sizeof(rng));
        rng.mode = J->rng_mode;
        rng.pixel = (uint32_t)(x + y * J->W);
        rng.k0 = (uint32_t)(J->seed & 0xFFFFFFFFu);
        rng.k1 = (uint32_t)(J->seed >> 32);
        if (J->rng_mode == 1)
            rng.state = (J->lcg_state && J->s_begin > 0) ? J->lcg_state[rng.pixel]
                                                         : (rng.pixel ^ (uint32_t)(J->seed & 0xFFFFFFFFu));   /* :990 */
        float* a = J->accum + ((size_t)y * J->W + x) * 3;
        for (int sidx = J->s_begin; sidx < J->s_end; ++sidx) {
            if (J->rng_mode != 1) { rng.state = 0u; rng.sample = (uint32_t)sidx; }
            const float jx = ((float)sidx + rand01(&rng)) / (float)J->spp;      /* :995-996 */
            const float jy = ((float)sidx + rand01(&rng)) / (float)J->spp;
            Ray ray = camera_ray(J->cam12, x, y, J->W, J->H, jx, jy);
            f3 c = ray_color(J->s, ray, &rng, J->max_depth, J->sun_enabled, J->sd, J->sr, &cnt);
            a[0] += c.x; a[1] += c.y; a[2] += c.z;
        }
        if (J->rng_mode == 1 && J->lcg_state) J->lcg_state[rng.pixel] = rng.state;
    }
    pthread_mutex_lock(&J->mu);
    J->counts[0] += cnt.primary; J->counts[1] += cnt.bounce; J->counts[2] += cnt.shadow;
    pthread_mutex_unlock(&J->mu);
}

void or_render(const OrScene* s, const float cam12[12], int W, int H, int spp_total, int max_depth,
               const float sun_dir[3], const float sun_radiance[3], int sun_enabled,
               uint64_t seed, int rng_mode, float gamma, int s_begin, int s_end,
               float* accum, unsigned char* rgb, uint32_t* lcg_state, uint64_t ray_counts[3]) {
    RenderJob J;
    memset(&J, 0, sizeof(J));
    J.s = s; J.cam12 = cam12; J.W = W; J.H = H;
    J.spp = spp_total < 1 ? 1 : spp_total;
    J.max_depth = max_depth; J.sun_enabled = sun_enabled; J.rng_mode = rng_mode;
    J.s_begin = s_begin; J.s_end = s_end;
    J.sd = F3(sun_dir[0], sun_dir[1], sun_dir[2]);
    J.sr = F3(sun_radiance[0], sun_radiance[1], sun_radiance[2]);
    J.seed = seed; J.accum = accum; J.lcg_state = lcg_state;
    pthread_mutex_init(&J.mu, NULL);
    or_parallel_for(H, render_row, &J);
    pthread_mutex_destroy(&J.mu);
    if (ray_counts) { ray_counts[0] += J.counts[0]; ray_counts[1] += J.counts[1]; ray_counts[2] += J.counts[2]; }
    if (rgb) or_tonemap(accum, W, H, J.spp, gamma, rgb);
}

/* ---- host side ----------------------------------------------------------------------------- */
int or_read_pose_file(const char* path, double* poses9, int max_poses) {
    /* main.cpp:139-173: skip empty lines and lines whose first character is '#'; a line needs
     * nine leading numbers (istream >> double), anything after them is ignored. */
    FILE* f = fopen(path, "r");
    if (!f) return -5;
    char line[4096];
    int n = 0;
    while (fgets(line, sizeof(line), f)) {
        size_t len = strlen(line);
        if (len && line[len - 1] == '\n') line[--len] = 0;
        if (len == 0 || line[0] == '#') continue;
        double v[9];
        char* p = line;
        int ok = 1;
        for (int k = 0; k < 9; ++k) {
            char* e;
            v[k] = strtod(p, &e);
            if (e == p) { ok = 0; break; }
            p = e;
        }
        if (!ok) continue;
        if (poses9 && n < max_poses) {
            for (int k = 0; k < 6; ++k) poses9[n * 9 + k] = v[k];
            for (int k = 6; k < 9; ++k) poses9[n * 9 + k] = (double)(float)v[k];
        }
        ++n;
    }
    fclose(f);
    return n;
}

void or_pose_to_model_frame(const double p[9], float cam_in_model[3], float sun_dir_model[3], double* sep_m) {
    /* main.cpp:334-357, rotate_yaw_deg_d :118-128, degrees_to_radians rtweekend.h:32-34 */
    const double pi = 3.1415926535897932385;
    const double cx = p[0] - p[3], cy = p[1] - p[4], cz = p[2] - p[5];
    const double lx = 0.0 - p[3], ly = 0.0 - p[4], lz = 0.0 - p[5];
    if (sep_m) *sep_m = sqrt(cx * cx + cy * cy + cz * cz);
    const double rad = (-p[6]) * pi / 180.0;
    const double c = cos(rad), s = sin(rad);
    cam_in_model[0] = (float)(c * cx + s * cz);
    cam_in_model[1] = (float)cy;
    cam_in_model[2] = (float)(-s * cx + c * cz);
    const double mx = c * lx + s * lz, my = ly, mz = -s * lx + c * lz;
    const double L = sqrt(mx * mx + my * my + mz * mz);
    if (L == 0.0) { sun_dir_model[0] = sun_dir_model[1] = sun_dir_model[2] = 0.f; return; }
    const double inv = 1.0 / L;
    sun_dir_model[0] = (float)(inv * mx);
    sun_dir_model[1] = (float)(inv * my);
    sun_dir_model[2] = (float)(inv * mz);
}

void or_camera_init(const float lf[3], float vfov_deg, int W, int H, float cam12[12], float uvw9[9]) {
    /* point_camera_at main.cpp:178-187 + camera::initialize camera.h:91-116; host float math,
     * vec3 operators of inc/vec3.h (division = multiplication by the rounded reciprocal). */
    const double pi = 3.1415926535897932385;
    const float lx = lf[0], ly = lf[1], lz = lf[2];
    const float focus = sqrtf(lx * lx + ly * ly + lz * lz);
    const float aspect = (float)W / (float)H;
    const float theta = (float)((double)vfov_deg * pi / 180.0);
    const float h = tanf(theta / 2.0f);
    const float vh = 2.0f * h, vw = aspect * vh;
    float inv = 1.0f / focus;
    const float wx = lx * inv, wy = ly * inv, wz = lz * inv;
    /* cross(vup=(0,1,0), w) */
    float ux = 1.0f * wz - 0.0f * wy, uy = 0.0f * wx - 0.0f * wz, uz = 0.0f * wy - 1.0f * wx;
    inv = 1.0f / sqrtf(ux * ux + uy * uy + uz * uz);
    ux *= inv; uy *= inv; uz *= inv;
    const float vx = wy * uz - wz * uy, vy = wz * ux - wx * uz, vz = wx * uy - wy * ux;
    const float fw = focus * vw, fh = focus * vh;
    const float hx = fw * ux, hy = fw * uy, hz = fw * uz;
    const float tx = fh * vx, ty = fh * vy, tz = fh * vz;
    cam12[0] = lx; cam12[1] = ly; cam12[2] = lz;
    cam12[3] = lx - 0.5f * hx - 0.5f * tx - focus * wx;
    cam12[4] = ly - 0.5f * hy - 0.5f * ty - focus * wy;
    cam12[5] = lz - 0.5f * hz - 0.5f * tz - focus * wz;
    cam12[6] = hx; cam12[7] = hy; cam12[8] = hz;
    cam12[9] = tx; cam12[10] = ty; cam12[11] = tz;
    if (uvw9) { uvw9[0] = ux; uvw9[1] = uy; uvw9[2] = uz; uvw9[3] = vx; uvw9[4] = vy; uvw9[5] = vz; uvw9[6] = wx; uvw9[7] = wy; uvw9[8] = wz; }
}
