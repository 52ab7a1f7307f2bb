// pose_camera.cpp -- host side of the render path that feeds the GPU: pose file, the
// double-precision world -> model-frame transform, and the pinhole camera set-up.
//
// Mirrors (behaviour, not code) /root/reference/src/main.cpp:139-173 (read_pose_file),
// :118-128 (rotate_yaw_deg_d), :334-357 (per-frame transform), :178-187 (point_camera_at)
// and /root/reference/inc/camera.h:91-116 (camera::initialize).  Built with
// -ffp-contract=off so the float/double expressions round exactly like the reference's
// host build (no FMA contraction on the host).
#include "dsrt.h"

#include <cmath>
#include <fstream>
#include <sstream>
#include <string>

namespace {

// rt_pi() of the reference, inc/rtweekend.h:27-29; the literal rounds to the double nearest pi.
constexpr double kPi = 3.1415926535897932385;

struct F3 { float x, y, z; };

inline F3 f3(float x, float y, float z) { return F3{x, y, z}; }
inline F3 sub(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline F3 scale(F3 a, float s) { return f3(s * a.x, s * a.y, s * a.z); }
inline float length(F3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
// vec3 operator/ multiplies by the rounded reciprocal (inc/vec3.h:101-105).
inline F3 unit(F3 a) { float inv = 1.0f / length(a); return f3(a.x * inv, a.y * inv, a.z * inv); }
inline F3 cross(F3 u, F3 v) {
    return f3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
inline DsrtVec3 out(F3 a) { DsrtVec3 r; r.x = a.x; r.y = a.y; r.z = a.z; return r; }

}  // namespace

extern "C" int dsrt_read_pose_file(const char* path, double* poses9, int max_poses) {
    if (!path) return DSRT_EINVAL;
    std::ifstream in(path);
    if (!in) return DSRT_EIO;
    int n = 0;
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;      // main.cpp:150-151
        std::istringstream iss(line);
        double v[9];
        bool ok = true;
        for (int k = 0; k < 9 && ok; ++k) ok = static_cast<bool>(iss >> v[k]);
        if (!ok) continue;                                   // malformed line: skipped silently
        if (poses9 && n < max_poses) {
            for (int k = 0; k < 6; ++k) poses9[n * 9 + k] = v[k];
            // yaw/pitch/roll are stored in a float vec3 by the reference (main.cpp:166).
            for (int k = 6; k < 9; ++k) poses9[n * 9 + k] = (double)(float)v[k];
        }
        ++n;
    }
    return n;
}

extern "C" void dsrt_pose_to_model_frame(const double p[9], float cam_in_model[3],
                                         float sun_dir_model[3], double* sep_m) {
    // camera and light (world origin) relative to the model, world axes
    const double crx = p[0] - p[3], cry = p[1] - p[4], crz = p[2] - p[5];
    const double lrx = 0.0 - p[3], lry = 0.0 - p[4], lrz = 0.0 - p[5];
    if (sep_m) *sep_m = std::sqrt(crx * crx + cry * cry + crz * crz);

    // R_world_model^T ~ yaw about +Y by -yaw degrees
    const double yaw_deg = -p[6];
    const double rad = yaw_deg * kPi / 180.0;
    const double c = std::cos(rad), s = std::sin(rad);
    const double cmx = c * crx + s * crz, cmy = cry, cmz = -s * crx + c * crz;
    const double lmx = c * lrx + s * lrz, lmy = lry, lmz = -s * lrx + c * lrz;

    cam_in_model[0] = (float)cmx; cam_in_model[1] = (float)cmy; cam_in_model[2] = (float)cmz;

    const double L = std::sqrt(lmx * lmx + lmy * lmy + lmz * lmz);
    if (L == 0.0) {
        sun_dir_model[0] = sun_dir_model[1] = sun_dir_model[2] = 0.0f;
    } else {
        const double inv = 1.0 / L;
        sun_dir_model[0] = (float)(inv * lmx);
        sun_dir_model[1] = (float)(inv * lmy);
        sun_dir_model[2] = (float)(inv * lmz);
    }
}

extern "C" void dsrt_camera_look_at_origin(const float lookfrom_[3], float vfov_deg, int width,
                                           int height, int spp, int max_depth, DsrtCamera* cam) {
    const F3 lookfrom = f3(lookfrom_[0], lookfrom_[1], lookfrom_[2]);
    const F3 lookat = f3(0.0f, 0.0f, 0.0f);
    const F3 vup = f3(0.0f, 1.0f, 0.0f);
    const float focus_dist = length(sub(lookfrom, lookat));

    const float aspect = float(width) / float(height);
    const float theta = (float)((double)vfov_deg * kPi / 180.0);
    const float h = tanf(theta / 2.0f);
    const float viewport_height = 2.0f * h;
    const float viewport_width = aspect * viewport_height;

    const F3 w = unit(sub(lookfrom, lookat));
    const F3 u = unit(cross(vup, w));
    const F3 v = cross(w, u);

    const F3 horizontal = scale(u, focus_dist * viewport_width);
    const F3 vertical = scale(v, focus_dist * viewport_height);
    const F3 llc = sub(sub(sub(lookfrom, scale(horizontal, 0.5f)), scale(vertical, 0.5f)),
                       scale(w, focus_dist));

    cam->origin = out(lookfrom);
    cam->lower_left_corner = out(llc);
    cam->horizontal = out(horizontal);
    cam->vertical = out(vertical);
    cam->u = out(u); cam->v = out(v); cam->w = out(w);
    cam->lens_radius = 0.0f * 0.5f;
    cam->image_width = width;
    cam->image_height = height;
    cam->samples_per_pixel = spp;
    cam->max_depth = max_depth;
}
