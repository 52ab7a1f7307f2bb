// ref_host_probe.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Exposes the reference's own host-side functions of the render path so golden vectors can
// be produced from the reference itself: this file #includes the UNMODIFIED
// /root/reference/src/main.cpp (found through -I/root/reference/src; its main() is renamed)
// and wraps its static functions read_pose_file (main.cpp:139-173), rotate_yaw_deg_d
// (:118-128), dnormalize/to_float_vec3 (:74-86), point_camera_at (:178-187) and, through it,
// camera::initialize (inc/camera.h:91-116).  Linked into oracle/_ref/libref_gpu.so.
#define main ref_main_not_used
#include "main.cpp"
#undef main

extern "C" {

int refhost_read_pose_file(const char* path, double* poses9, int max_poses) {
    std::vector<PoseEntry> poses;
    if (!read_pose_file(path, poses)) return 0;
    int n = (int)poses.size();
    for (int i = 0; i < n && i < max_poses; ++i) {
        const PoseEntry& p = poses[i];
        double* o = poses9 + 9 * i;
        o[0] = p.cam_pos_world_d.x; o[1] = p.cam_pos_world_d.y; o[2] = p.cam_pos_world_d.z;
        o[3] = p.model_pos_world_d.x; o[4] = p.model_pos_world_d.y; o[5] = p.model_pos_world_d.z;
        o[6] = p.model_euler_deg.x(); o[7] = p.model_euler_deg.y(); o[8] = p.model_euler_deg.z();
    }
    return n;
}

// main.cpp:334-357, using the reference's own helpers
void refhost_pose_to_model_frame(const double p[9], float cam_in_model[3], float sun_dir_model[3], double* sep_m) {
    dvec3 cam_w(p[0], p[1], p[2]), model_w(p[3], p[4], p[5]), light(0.0, 0.0, 0.0);
    double yaw_deg = (double)(float)p[6];
    dvec3 cam_rel = cam_w - model_w;
    dvec3 light_rel = light - model_w;
    *sep_m = dlength(cam_rel);
    dvec3 cam_m = rotate_yaw_deg_d(cam_rel, -yaw_deg);
    dvec3 light_m = rotate_yaw_deg_d(light_rel, -yaw_deg);
    vec3 c = to_float_vec3(cam_m);
    vec3 s = to_float_vec3(dnormalize(light_m));
    cam_in_model[0] = c.x(); cam_in_model[1] = c.y(); cam_in_model[2] = c.z();
    sun_dir_model[0] = s.x(); sun_dir_model[1] = s.y(); sun_dir_model[2] = s.z();
}

// point_camera_at(cam, lookfrom, origin) with the reference's default vfov handling
void refhost_camera(const float lookfrom[3], float vfov, int W, int H, float cam12[12], float uvw9[9]) {
    camera cam;
    cam.image_width = W; cam.image_height = H;
    cam.vfov = vfov; cam.aperture = 0.0;
    point_camera_at(cam, vec3(lookfrom[0], lookfrom[1], lookfrom[2]), vec3(0, 0, 0));
    GPUCamera g = cam.toGPUCamera();
    const vec3* v[4] = {&g.origin, &g.lower_left_corner, &g.horizontal, &g.vertical};
    for (int k = 0; k < 4; ++k) { cam12[3 * k] = v[k]->x(); cam12[3 * k + 1] = v[k]->y(); cam12[3 * k + 2] = v[k]->z(); }
    const vec3* b[3] = {&g.u, &g.v, &g.w};
    for (int k = 0; k < 3; ++k) { uvw9[3 * k] = b[k]->x(); uvw9[3 * k + 1] = b[k]->y(); uvw9[3 * k + 2] = b[k]->z(); }
}

}  // extern "C"
