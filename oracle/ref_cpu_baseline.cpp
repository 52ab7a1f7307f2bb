// ref_cpu_baseline.cpp -- TEST/BENCH INFRASTRUCTURE ONLY: the reported CPU baseline.
//
// BASELINE.json asks for "the reference's CPU camera/bvh_node path from inc/ timed on the
// box's host cores".  This file drives the reference's UNMODIFIED CPU classes, compiled from
// /root/reference/inc where they lie (oracle/Makefile -> oracle/_ref/libref_cpu.so):
//   camera::initialize   inc/camera.h:91-116      triangle::hit      inc/triangle.h:31-61
//   bvh_node (x-sorted)  inc/bvh.h:14-136         aabb::hit          inc/aabb.h:33-56
//   hittable_list        inc/hittable_list.h      interval           inc/interval.h
//   lambertian/metal/dielectric getters           inc/material.h
// The reference ships NO CPU render loop, so the loop below
// restates the GPU estimator ray_color (/root/reference/src/gpu_render.cu:715-936) on top of
// bvh_node::hit, with the reference's device LCG (inc/rtweekend.h:126-133) and a std::thread
// pool over image rows.  It is a timing baseline, not a parity oracle: triangle::hit works
// in double with an open t-interval and a 1e-9 determinant threshold, unlike the GPU path.
#include "rtweekend.h"
#include "camera.h"
#include "hittable_list.h"
#include "triangle.h"
#include "bvh.h"
#include "material.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace {

struct RefCpuMaterial {   // GPUMaterial layout
    int type, albedo_tex, pad0, pad1;
    float albedo[3], emissive[3];
    float fuzz, ref_idx;
};

struct RefCpu {
    hittable_list world;
    std::shared_ptr<bvh_node> bvh;
    double build_s = 0;
};

struct Counts { unsigned long long primary = 0, bounce = 0, shadow = 0; };

inline float rnd(uint32_t& s) { return random_float_device(s); }

inline vec3 rand_in_unit_sphere(uint32_t& s) {
    while (true) {
        vec3 p(rnd(s) * 2.0f - 1.0f, rnd(s) * 2.0f - 1.0f, rnd(s) * 2.0f - 1.0f);
        if (p.length_squared() >= 1.0f) continue;
        return p;
    }
}

inline vec3 clamp01v(const vec3& a) {
    return vec3(fminf(1.f, fmaxf(0.f, a.x())), fminf(1.f, fmaxf(0.f, a.y())), fminf(1.f, fmaxf(0.f, a.z())));
}

const float PI_F = 3.14159265358979323846f;

vec3 ray_color_cpu(const hittable& scene, ray r, uint32_t& rng, int max_depth_in, const vec3& sun_dir,
                   const vec3& sun_rad, Counts& cnt) {
    vec3 L(0, 0, 0), T(1, 1, 1);
    const int max_depth = max_depth_in > 0 ? max_depth_in : 12;
    for (int depth = 0; depth < max_depth; ++depth) {
        if (depth >= 5) {
            float p = fminf(fmaxf(T.x(), fmaxf(T.y(), T.z())), 0.95f);
            if (rnd(rng) > p) break;
            T = T * (1.0f / p);
        }
        hit_record rec;
        if (depth == 0) cnt.primary++; else cnt.bounce++;
        if (!scene.hit(r, interval(0.001, 1e9), rec)) break;
        const material* m = rec.mat_ptr.get();
        if (auto em = dynamic_cast<const diffuse_light*>(m)) { L += T * em->emit_value(); break; }
        if (auto di = dynamic_cast<const dielectric*>(m)) {
            float eta = (float)di->ior_value();
            if (eta <= 0.0f || !std::isfinite(eta)) eta = 1.5f;
            float ratio = rec.front_face ? (1.0f / eta) : eta;
            vec3 ud = unit_vector(r.direction());
            float cos_t = fminf(dot(-ud, rec.normal), 1.0f);
            float sin_t = sqrtf(fmaxf(0.0f, 1.0f - cos_t * cos_t));
            float r0 = (1.0f - ratio) / (1.0f + ratio);
            r0 = r0 * r0;
            float refl = r0 + (1.0f - r0) * powf(1.0f - cos_t, 5.0f);
            vec3 nd = (ratio * sin_t > 1.0f || refl > rnd(rng)) ? reflect(ud, rec.normal) : refract(unit_vector(ud), rec.normal, ratio);
            r = ray(rec.p, nd);
            continue;
        }
        if (auto me = dynamic_cast<const metal*>(m)) {
            vec3 refl = reflect(unit_vector(r.direction()), rec.normal);
            float fuzz = fmaxf(0.0f, fminf(1.0f, (float)me->fuzz_value()));
            vec3 nd = refl + fuzz * rand_in_unit_sphere(rng);
            if (!(dot(nd, rec.normal) > 0.0f)) break;
            T = T * me->albedo_value();
            r = ray(rec.p, nd);
            continue;
        }
        vec3 albedo(0.73f, 0.73f, 0.73f);
        if (auto la = dynamic_cast<const lambertian*>(m)) albedo = la->albedo_value();
        {
            vec3 Ldir = unit_vector(-sun_dir);
            float cos_l = fmaxf(0.0f, dot(rec.normal, Ldir));
            if (cos_l > 0.0f) {
                ray sh(rec.p + 1e-3f * rec.normal, Ldir);
                hit_record srec;
                cnt.shadow++;
                if (!scene.hit(sh, interval(0.001, 1e9), srec)) {
                    float pdf = cos_l / PI_F;
                    float wgt = pdf / (0.5f + 0.5f * pdf);
                    L += T * (albedo * (sun_rad * wgt));
                }
            }
        }
        vec3 w = unit_vector(rec.normal);
        vec3 a = (fabsf(w.x()) > 0.9f) ? vec3(0, 1, 0) : vec3(1, 0, 0);
        vec3 vv = unit_vector(cross(w, a));
        vec3 uu = cross(vv, w);
        float r1 = rnd(rng), r2 = rnd(rng);
        float lz = sqrtf(1.0f - r2), phi = 2.0f * PI_F * r1;
        vec3 wd = unit_vector(uu * (cosf(phi) * sqrtf(r2)) + vv * (sinf(phi) * sqrtf(r2)) + w * lz);
        float cos_b = fmaxf(0.0f, dot(wd, rec.normal));
        if (!(cos_b > 0.0f)) break;
        T = T * albedo;
        r = ray(rec.p, wd);
    }
    return clamp01v(L);
}

}  // namespace

extern "C" {

void* refcpu_create(const float* tris, int n, const RefCpuMaterial* mats, int nmats, double* build_seconds) {
    std::vector<std::shared_ptr<material>> M;
    for (int i = 0; i < nmats; ++i) {
        const RefCpuMaterial& m = mats[i];
        vec3 a(m.albedo[0], m.albedo[1], m.albedo[2]);
        if (m.type == 1) M.push_back(std::make_shared<metal>(a, (double)m.fuzz));
        else if (m.type == 2) M.push_back(std::make_shared<dielectric>((double)m.ref_idx));
        else if (m.type == 3) M.push_back(std::make_shared<diffuse_light>(vec3(m.emissive[0], m.emissive[1], m.emissive[2])));
        else M.push_back(std::make_shared<lambertian>(a));
    }
    RefCpu* R = new RefCpu();
    for (int i = 0; i < n; ++i) {
        const float* p = tris + (size_t)i * 29;
        int mid;
        std::memcpy(&mid, p + 27, 4);
        R->world.add(std::make_shared<triangle>(vec3(p[0], p[1], p[2]), vec3(p[3], p[4], p[5]), vec3(p[6], p[7], p[8]), M[(size_t)mid]));
    }
    auto t0 = std::chrono::steady_clock::now();
    if (n > 0) R->bvh = std::make_shared<bvh_node>(R->world);
    R->build_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (build_seconds) *build_seconds = R->build_s;
    return R;
}

void refcpu_destroy(void* h) { delete (RefCpu*)h; }

// Renders samples [s_begin,s_end) of the pixels (x0 + i*step_x, y0 + j*step_y) inside the
// image: a bounded, evenly spread sample of the full frame's workload.  accum (may be NULL):
// W*H*3 floats, row 0 = bottom, receives the sum of clamped samples.  Returns wall seconds.
double refcpu_render(void* h, const float lookfrom[3], float vfov, int W, int H, int spp_total, int max_depth,
                     const float sun_dir[3], const float sun_rad[3], unsigned long long seed,
                     int step_x, int step_y, int s_begin, int s_end, int threads,
                     float* accum, unsigned long long counts[3]) {
    RefCpu* R = (RefCpu*)h;
    camera cam;
    cam.image_width = W; cam.image_height = H;
    cam.vfov = vfov; cam.aperture = 0.0f;
    cam.lookfrom = vec3(lookfrom[0], lookfrom[1], lookfrom[2]);
    cam.lookat = vec3(0, 0, 0);
    cam.vup = vec3(0, 1, 0);
    cam.focus_dist = (cam.lookfrom - cam.lookat).length();
    cam.initialize();
    const vec3 sd(sun_dir[0], sun_dir[1], sun_dir[2]), sr(sun_rad[0], sun_rad[1], sun_rad[2]);
    const int spp = spp_total < 1 ? 1 : spp_total;
    if (step_x < 1) step_x = 1;
    if (step_y < 1) step_y = 1;
    if (threads < 1) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    const hittable* scene = R->bvh ? (const hittable*)R->bvh.get() : (const hittable*)&R->world;
    std::atomic<int> next_row(0);
    std::vector<Counts> per_thread((size_t)threads);
    const int rows = (H + step_y - 1) / step_y;
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid) {
        Counts c;
        for (;;) {
            int j = next_row.fetch_add(1);
            if (j >= rows) break;
            int y = j * step_y;
            for (int x = 0; x < W; x += step_x) {
                uint32_t rng = (uint32_t)(x + y * W) ^ (uint32_t)(seed & 0xFFFFFFFFu);
                vec3 acc(0, 0, 0);
                for (int s = s_begin; s < s_end; ++s) {
                    float jx = ((float)s + rnd(rng)) / (float)spp;
                    float jy = ((float)s + rnd(rng)) / (float)spp;
                    float u = ((float)x + jx) / (float)(W - 1), v = ((float)y + jy) / (float)(H - 1);
                    vec3 dir = cam.lower_left_corner + u * cam.horizontal + v * cam.vertical - cam.origin;
                    acc += ray_color_cpu(*scene, ray(cam.origin, dir), rng, max_depth, sd, sr, c);
                }
                if (accum) {
                    float* a = accum + ((size_t)y * W + x) * 3;
                    a[0] += acc.x(); a[1] += acc.y(); a[2] += acc.z();
                }
            }
        }
        per_thread[(size_t)tid] = c;
    };
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (counts) for (const Counts& c : per_thread) { counts[0] += c.primary; counts[1] += c.bounce; counts[2] += c.shadow; }
    return secs;
}

int refcpu_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
