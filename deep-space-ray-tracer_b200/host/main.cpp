// main.cpp -- the `ray_tracer` command line, drop-in for /root/reference/src/main.cpp.
//
// Same flags (--input_txt <file>, --output_dir <dir>, --upscale; unknown flags ignored,
// main.cpp:194-215), same pose file and fallback pose (:139-173, :272-287), same per-frame
// banner and skip rule (:319-345), same defaults (800x450, 1000 spp, depth 50, vfov 40,
// :254-260), same outputs (<out>/frame_%04zu.ppm + .png, :418-426; the output directory is
// emptied first, :41-50).  What changes is behind the C ABI: the scene is uploaded and its
// BVH built once, each frame only sets camera + sun, and frames are spread over the GPUs of
// the node (frame-parallel, no collective) or one frame's samples are sharded across them
// and summed with a single NCCL reduce.
//
// The mesh: the reference hard-codes ../../iss_model/ISS_stationary.obj (main.cpp:239) and
// silently renders an empty scene when it is missing; here --obj <path> (default: that same
// path) is loaded by dsrt_mesh_load_obj (OBJ + MTL + map_Kd textures, host/obj_loader.cpp), and
// when the file cannot be opened the deterministic synthetic station stands in (the NASA model
// is not redistributable with either tree).
//
// Runs over several GPUs use one PROCESS per GPU (forked before the first CUDA call), the same
// model as `torchrun bench.py`.  CUDA start-up costs ~2.4 s per process on a B200 box (measured;
// it is paid before the first kernel and is outside this library's control) and is reported
// separately from the frame loop in --stats_json.  Each worker owns its GPU, renders frames
// k = g, g+G, ... and writes its own files; there is no collective.
//
// Frames are written by a writer thread per GPU (PPM, then the in-process PNG encoder that
// replaces the reference's `magick` subprocess, main.cpp:28-36), so a GPU renders frame k+1
// while frame k is being compressed.
//
// Extra, optional flags (defaults reproduce the reference): --width --height --spp
// --max_depth --seed --obj <path> --obj_scale S --synthetic_tris N --diffuse_only --gpus G
// --shard frame|sample --frames a:b --rng philox|lcg --no_png --stats_json <file>.
#include "dsrt.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <sys/wait.h>
#include <unistd.h>

#ifdef DSRT_WITH_NCCL
#include <cuda_runtime.h>
#include <nccl.h>
#endif

namespace fs = std::filesystem;

namespace {

struct Options {
    std::string pose_file, out_dir = "output", stats_json, shard = "frame", rng = "philox";
    std::string obj_path = "../../iss_model/ISS_stationary.obj";      // main.cpp:239
    double obj_scale = 1.0;
    bool upscale = false, png = true, diffuse_only = false;
    int width = 800, height = 450, spp = 1000, max_depth = 50;
    int synthetic_tris = 1000000, gpus = 1, frame_begin = 0, frame_end = -1;
    float vfov = 40.0f;
    unsigned long long seed = 1337ULL;        // gpu_scene_builder.cpp:583
};

void parse_args(int argc, char** argv, Options& o) {
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&](const char* name) -> const char* { return (a == name && i + 1 < argc) ? argv[++i] : nullptr; };
        if (const char* v = next("--input_txt")) o.pose_file = v;
        else if (const char* v = next("--output_dir")) o.out_dir = v;
        else if (a == "--upscale") o.upscale = true;
        else if (const char* v = next("--width")) o.width = std::atoi(v);
        else if (const char* v = next("--height")) o.height = std::atoi(v);
        else if (const char* v = next("--spp")) o.spp = std::atoi(v);
        else if (const char* v = next("--max_depth")) o.max_depth = std::atoi(v);
        else if (const char* v = next("--seed")) o.seed = std::strtoull(v, nullptr, 10);
        else if (const char* v = next("--synthetic_tris")) o.synthetic_tris = std::atoi(v);
        else if (const char* v = next("--obj")) o.obj_path = v;
        else if (const char* v = next("--obj_scale")) o.obj_scale = std::atof(v);
        else if (a == "--diffuse_only") o.diffuse_only = true;
        else if (const char* v = next("--gpus")) o.gpus = std::atoi(v);
        else if (const char* v = next("--shard")) o.shard = v;
        else if (const char* v = next("--rng")) o.rng = v;
        else if (a == "--no_png") o.png = false;
        else if (const char* v = next("--stats_json")) o.stats_json = v;
        else if (const char* v = next("--frames")) {
            const char* c = std::strchr(v, ':');
            o.frame_begin = std::atoi(v);
            o.frame_end = c ? std::atoi(c + 1) : o.frame_begin + 1;
        }
        // anything else: ignored, as in the reference
    }
}

void prepare_output_dir(const std::string& dir) {
    if (fs::exists(dir)) {
        for (auto& entry : fs::directory_iterator(dir)) fs::remove_all(entry.path());
    } else {
        fs::create_directories(dir);
    }
}

// Host copy of the scene in the layouts dsrt_scene_upload takes.
struct SceneData {
    std::vector<DsrtTriangle> tris;
    std::vector<DsrtMaterial> mats;
    std::vector<DsrtTextureHeader> tex;
    std::vector<float> pool;
};

int upload(dsrt_ctx* ctx, const SceneData& sc) {
    return dsrt_scene_upload(ctx, sc.tris.data(), (int)sc.tris.size(), sc.mats.data(), (int)sc.mats.size(), nullptr, 0,
                             sc.tex.empty() ? nullptr : sc.tex.data(), (int)sc.tex.size(),
                             sc.pool.empty() ? nullptr : sc.pool.data(), (int)sc.pool.size());
}

struct FrameJob {
    size_t index;
    double pose[9];
    float cam_in_model[3];
    float sun_dir[3];
};

struct FrameResult {
    double device_ms = 0, rays = 0;
    bool done = false;
};

// Wall-clock split of a run: what is paid once per GPU (CUDA context, scene upload, BVH build,
// work-memory allocation inside the first frame) versus the steady frame loop.
struct SetupTimes {
    double create_s = 0, upload_s = 0, first_frame_s = 0, loop_s = 0, teardown_s = 0;
};
inline double seconds_since(std::chrono::high_resolution_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
}

const float kSunRadiance[3] = {100000.0f, 95000.0f, 90000.0f};   // gpu_scene_builder.cpp:598

DsrtRenderParams make_params(const Options& o) {
    DsrtRenderParams p;
    std::memset(&p, 0, sizeof(p));
    p.img_width = o.width; p.img_height = o.height;
    p.samples_per_pixel = o.spp; p.max_depth = o.max_depth;
    p.use_bvh = 1;
    p.rng_mode = o.rng == "lcg" ? DSRT_RNG_REF_LCG : DSRT_RNG_PHILOX;
    p.gamma = 2.0f; p.exposure = 50.0f;                       // gpu_scene_builder.cpp:579-580
    return p;
}

bool save_frame(const Options& o, size_t index, const unsigned char* rgb) {
    char name[64];
    std::snprintf(name, sizeof(name), "frame_%04zu.ppm", index);
    const std::string ppm = o.out_dir + "/" + name;
    std::snprintf(name, sizeof(name), "frame_%04zu.png", index);
    const std::string png = o.out_dir + "/" + name;
    if (dsrt_write_ppm(ppm.c_str(), rgb, o.width, o.height) != DSRT_OK) return false;
    if (o.png && dsrt_write_png(png.c_str(), rgb, o.width, o.height) != DSRT_OK) return false;
    return true;
}

std::mutex g_print;

// Writes finished frames on its own thread so the GPU that produced them keeps rendering.
// Frames travel in three page-locked buffers that circulate between the render loop (acquire ->
// read back -> push) and the writer (save -> release), so at most two frames wait to be written
// and the producer blocks beyond that.  Created after the ctx (page-locked memory needs CUDA).
class FrameWriter {
public:
    FrameWriter(const Options& o, std::vector<FrameResult>& results)
        : o_(o), results_(results), bytes_((size_t)o.width * o.height * 3) {
        for (int k = 0; k < 3; ++k) {
            void* p = nullptr;
            if (dsrt_host_alloc(&p, bytes_) != DSRT_OK) { p = std::malloc(bytes_); pageable_.push_back(p); }
            all_.push_back((unsigned char*)p);
            free_.push_back((unsigned char*)p);
        }
        th_ = std::thread([this] { run(); });
    }
    ~FrameWriter() {
        finish();
        for (unsigned char* p : all_) {
            bool pageable = false;
            for (void* q : pageable_) pageable = pageable || q == p;
            if (pageable) std::free(p); else dsrt_host_free(p);
        }
    }
    size_t bytes() const { return bytes_; }
    unsigned char* acquire() {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [this] { return !free_.empty(); });
        unsigned char* p = free_.back();
        free_.pop_back();
        return p;
    }
    void release(unsigned char* p) {
        std::lock_guard<std::mutex> l(mu_);
        free_.push_back(p);
        cv_.notify_all();
    }
    void push(size_t slot, size_t index, unsigned char* rgb) {
        std::lock_guard<std::mutex> l(mu_);
        q_.push_back(Item{slot, index, rgb});
        cv_.notify_all();
    }
    void finish() {
        {
            std::lock_guard<std::mutex> l(mu_);
            if (closed_) return;
            closed_ = true;
        }
        cv_.notify_all();
        if (th_.joinable()) th_.join();
    }
private:
    struct Item { size_t slot, index; unsigned char* rgb; };
    void run() {
        for (;;) {
            Item it;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [this] { return !q_.empty() || closed_; });
                if (q_.empty()) return;
                it = q_.front();
                q_.pop_front();
            }
            results_[it.slot].done = save_frame(o_, it.index, it.rgb);
            release(it.rgb);
        }
    }
    const Options& o_;
    std::vector<FrameResult>& results_;
    size_t bytes_;
    std::vector<unsigned char*> all_, free_;
    std::vector<void*> pageable_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Item> q_;
    bool closed_ = false;
    std::thread th_;
};

// One GPU renders the frames jobs[k] with k % stride == offset.
void frame_worker(int device, int label, const Options& o, const SceneData& scene, const std::vector<FrameJob>& jobs,
                  int offset, int stride, std::vector<FrameResult>& results, std::atomic<int>& errors, SetupTimes& times) {
    dsrt_ctx* ctx = nullptr;
    auto t0 = std::chrono::high_resolution_clock::now();
    if (dsrt_create(&ctx, device) != DSRT_OK) {
        std::lock_guard<std::mutex> l(g_print);
        std::fprintf(stderr, "GPU %d: %s\n", label, dsrt_last_error(nullptr));
        ++errors;
        return;
    }
    times.create_s = seconds_since(t0);
    t0 = std::chrono::high_resolution_clock::now();
    if (upload(ctx, scene) != DSRT_OK) {
        std::lock_guard<std::mutex> l(g_print);
        std::fprintf(stderr, "GPU %d: %s\n", label, dsrt_last_error(ctx));
        ++errors;
        dsrt_destroy(ctx);
        return;
    }
    times.upload_s = seconds_since(t0);
    {
        DsrtSceneStats st;
        dsrt_scene_stats(ctx, &st);
        std::lock_guard<std::mutex> l(g_print);
        std::cout << "GPU " << label << " scene build time: " << st.build_ms << " ms (" << st.num_wide_nodes
                  << " wide nodes, upload " << st.upload_ms << " ms)\n";
    }
    const DsrtRenderParams params = make_params(o);
    FrameWriter writer(o, results);
    t0 = std::chrono::high_resolution_clock::now();
    for (size_t k = (size_t)offset; k < jobs.size(); k += (size_t)stride) {
        const FrameJob& j = jobs[k];
        unsigned char* rgb = writer.acquire();
        DsrtCamera cam;
        dsrt_camera_look_at_origin(j.cam_in_model, o.vfov, o.width, o.height, o.spp, o.max_depth, &cam);
        int rc = dsrt_render_frame(ctx, &cam, j.sun_dir, kSunRadiance, 1, &params, o.seed, 0, o.spp, nullptr, 0);
        if (rc == DSRT_OK) rc = dsrt_readback_rgb8(ctx, o.spp, params.gamma, rgb, writer.bytes());
        if (k == (size_t)offset) times.first_frame_s = seconds_since(t0);
        if (rc != DSRT_OK) {
            std::lock_guard<std::mutex> l(g_print);
            std::fprintf(stderr, "frame %zu on GPU %d: %s\n", j.index, label, dsrt_last_error(ctx));
            ++errors;
            writer.release(rgb);
            continue;                       // no stale file is renamed into place
        }
        DsrtFrameStats fs;
        dsrt_frame_stats(ctx, &fs);
        results[k].device_ms = fs.device_ms;
        results[k].rays = (double)(fs.primary_rays + fs.bounce_rays + fs.shadow_rays);
        writer.push(k, j.index, rgb);
        std::lock_guard<std::mutex> l(g_print);
        char name[64];
        std::snprintf(name, sizeof(name), "frame_%04zu.png", j.index);
        std::cout << "Saved " << o.out_dir << "/" << name << "  [GPU " << label << ", " << fs.device_ms << " ms, "
                  << results[k].rays / (fs.device_ms * 1e3) << " Mrays/s]\n";
    }
    writer.finish();
    times.loop_s = seconds_since(t0);
    t0 = std::chrono::high_resolution_clock::now();
    dsrt_destroy(ctx);
    times.teardown_s = seconds_since(t0);
}

// Gives a forked worker process its own GPU: with `isolate` it sees only that device (so the CUDA
// runtime initialises one device instead of all of them) and uses ordinal 0, otherwise it uses
// ordinal g.  Returns the ordinal to pass to dsrt_create.
int claim_gpu(int g, bool isolate) {
    if (!isolate) return g;
    std::string pick = std::to_string(g);
    if (const char* vis = std::getenv("CUDA_VISIBLE_DEVICES")) {
        std::vector<std::string> tok;
        std::string cur;
        for (const char* c = vis;; ++c) {
            if (*c == ',' || *c == 0) { tok.push_back(cur); cur.clear(); if (*c == 0) break; }
            else cur.push_back(*c);
        }
        if ((size_t)g < tok.size() && !tok[(size_t)g].empty()) pick = tok[(size_t)g];
    }
    setenv("CUDA_VISIBLE_DEVICES", pick.c_str(), 1);
    return 0;
}

bool write_all(int fd, const void* src, size_t n) {
    size_t put = 0;
    while (put < n) {
        const ssize_t r = write(fd, (const char*)src + put, n - put);
        if (r <= 0) return false;
        put += (size_t)r;
    }
    return true;
}
bool read_all(int fd, void* dst, size_t n) {
    size_t got = 0;
    while (got < n) {
        const ssize_t r = read(fd, (char*)dst + got, n - got);
        if (r <= 0) return false;
        got += (size_t)r;
    }
    return true;
}

struct Wire { size_t k; FrameResult r; };

#ifdef DSRT_WITH_NCCL
// One frame at a time, its samples split over the GPUs -- one worker process per GPU (as in the
// frame-parallel mode, and for the same reason).  Worker g renders the global sample indices
// [spp*g/G, spp*(g+1)/G) of every pixel; one ncclReduce(sum) of the float accumulation buffer over
// NVLink brings the frame to worker 0, which tone-maps and writes it.  The NCCL id is created by
// worker 0 and relayed by the parent, which itself never touches CUDA.
int sample_worker(int g, int G, int id_fd, int out_fd, const Options& o, const SceneData& scene,
                  const std::vector<FrameJob>& jobs, bool isolate) {
    const int device = claim_gpu(g, isolate);
    SetupTimes times;
    auto t0 = std::chrono::high_resolution_clock::now();
    dsrt_ctx* ctx = nullptr;
    if (dsrt_create(&ctx, device) != DSRT_OK) { std::fprintf(stderr, "GPU %d: %s\n", g, dsrt_last_error(nullptr)); return 1; }
    times.create_s = seconds_since(t0);
    ncclUniqueId id;
    if (g == 0) {
        if (ncclGetUniqueId(&id) != ncclSuccess || !write_all(out_fd, &id, sizeof(id))) { std::fprintf(stderr, "ncclGetUniqueId failed\n"); return 1; }
    }
    if (!read_all(id_fd, &id, sizeof(id))) { std::fprintf(stderr, "GPU %d: no NCCL id\n", g); return 1; }
    t0 = std::chrono::high_resolution_clock::now();
    if (upload(ctx, scene) != DSRT_OK) { std::fprintf(stderr, "GPU %d: %s\n", g, dsrt_last_error(ctx)); return 1; }
    ncclComm_t comm;
    cudaSetDevice(device);
    if (ncclCommInitRank(&comm, G, id, g) != ncclSuccess) { std::fprintf(stderr, "GPU %d: ncclCommInitRank failed\n", g); return 1; }
    times.upload_s = seconds_since(t0);

    const DsrtRenderParams params = make_params(o);
    std::vector<FrameResult> results(jobs.size());
    int errors = 0;
    {
        FrameWriter writer(o, results);
        t0 = std::chrono::high_resolution_clock::now();
        for (size_t k = 0; k < jobs.size(); ++k) {
            const FrameJob& j = jobs[k];
            DsrtCamera cam;
            dsrt_camera_look_at_origin(j.cam_in_model, o.vfov, o.width, o.height, o.spp, o.max_depth, &cam);
            // global sample indices stay global: the jitter depends on s/spp (gpu_render.cu:995-996)
            const int s0 = (int)((long long)o.spp * g / G), s1 = (int)((long long)o.spp * (g + 1) / G);
            const bool rendered = dsrt_render_frame(ctx, &cam, j.sun_dir, kSunRadiance, 1, &params, o.seed, s0, s1, nullptr, 0) == DSRT_OK;
            if (!rendered) { std::fprintf(stderr, "frame %zu on GPU %d: %s\n", j.index, g, dsrt_last_error(ctx)); ++errors; }
            DsrtFrameStats fs{};
            dsrt_frame_stats(ctx, &fs);
            // every worker takes part in the reduce even after a failed render (a zero buffer), so the
            // collective cannot dead-lock
            if (dsrt_reduce_accum(ctx, comm, 0) != DSRT_OK) { std::fprintf(stderr, "GPU %d: %s\n", g, dsrt_last_error(ctx)); ++errors; break; }
            results[k].device_ms = rendered ? fs.device_ms : 0.0;
            results[k].rays = rendered ? (double)(fs.primary_rays + fs.bounce_rays + fs.shadow_rays) : 0.0;
            if (g == 0) {
                unsigned char* rgb = writer.acquire();
                if (dsrt_readback_rgb8(ctx, o.spp, params.gamma, rgb, writer.bytes()) != DSRT_OK) { writer.release(rgb); ++errors; continue; }
                writer.push(k, j.index, rgb);
                std::cout << "Saved frame " << j.index << "  [" << G << " GPUs sample-sharded, GPU 0: " << fs.device_ms << " ms]\n";
            } else {
                results[k].done = rendered;
            }
            if (k == 0) times.first_frame_s = seconds_since(t0);
        }
        writer.finish();
        times.loop_s = seconds_since(t0);
    }
    t0 = std::chrono::high_resolution_clock::now();
    ncclCommDestroy(comm);
    dsrt_destroy(ctx);
    times.teardown_s = seconds_since(t0);
    std::cout.flush();
    bool ok = write_all(out_fd, &times, sizeof(times));
    for (size_t k = 0; k < jobs.size() && ok; ++k) {
        const Wire w{k, results[k]};
        ok = write_all(out_fd, &w, sizeof(w));
    }
    return (errors || !ok) ? 1 : 0;
}

int sample_sharded(const Options& o, const SceneData& scene, const std::vector<FrameJob>& jobs,
                   std::vector<FrameResult>& results, std::vector<SetupTimes>& times, bool isolate) {
    const int G = o.gpus;
    std::cout.flush();
    std::fflush(stdout);
    std::vector<pid_t> pids((size_t)G, -1);
    std::vector<int> to_child((size_t)G, -1), from_child((size_t)G, -1);
    for (int g = 0; g < G; ++g) {
        int down[2], up[2];
        if (pipe(down) != 0 || pipe(up) != 0) { std::perror("pipe"); return 1; }
        const pid_t pid = fork();
        if (pid < 0) { std::perror("fork"); return 1; }
        if (pid == 0) {
            close(down[1]); close(up[0]);
            for (int q = 0; q < g; ++q) { close(to_child[(size_t)q]); close(from_child[(size_t)q]); }
            const int rc = sample_worker(g, G, down[0], up[1], o, scene, jobs, isolate);
            close(up[1]);
            _exit(rc);
        }
        close(down[0]); close(up[1]);
        pids[(size_t)g] = pid; to_child[(size_t)g] = down[1]; from_child[(size_t)g] = up[0];
    }
    int errors = 0;
    ncclUniqueId id;
    if (read_all(from_child[0], &id, sizeof(id))) {
        for (int g = 0; g < G; ++g) if (!write_all(to_child[(size_t)g], &id, sizeof(id))) ++errors;
    } else {
        ++errors;
    }
    for (int g = 0; g < G; ++g) close(to_child[(size_t)g]);       // a worker waiting for an id that never comes sees EOF
    std::vector<std::vector<FrameResult>> per((size_t)G, std::vector<FrameResult>(jobs.size()));
    for (int g = 0; g < G; ++g) {
        if (!read_all(from_child[(size_t)g], &times[(size_t)g], sizeof(SetupTimes))) ++errors;
        Wire w;
        while (read_all(from_child[(size_t)g], &w, sizeof(w)))
            if (w.k < jobs.size()) per[(size_t)g][w.k] = w.r;
        close(from_child[(size_t)g]);
        int status = 0;
        waitpid(pids[(size_t)g], &status, 0);
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) ++errors;
    }
    for (size_t k = 0; k < jobs.size(); ++k) {
        bool done = true;
        double mx = 0, total = 0;
        for (int g = 0; g < G; ++g) {
            done = done && per[(size_t)g][k].done;
            mx = std::max(mx, per[(size_t)g][k].device_ms);
            total += per[(size_t)g][k].rays;
        }
        results[k].done = done; results[k].device_ms = mx; results[k].rays = total;
    }
    return errors ? 1 : 0;
}
#endif

}  // namespace

int main(int argc, char** argv) {
    using clk = std::chrono::high_resolution_clock;
    Options o;
    parse_args(argc, argv, o);
    prepare_output_dir(o.out_dir);

    std::cout << "Using input_txt : " << (o.pose_file.empty() ? "(none)" : o.pose_file) << "\n";
    std::cout << "Using output_dir: " << o.out_dir << "\n";
    if (o.upscale) std::cout << "--upscale: the Real-ESRGAN post-process is outside this renderer; ignored.\n";
    const auto total_start = clk::now();

    // The mesh is static in the model frame: loaded (or generated) once, uploaded once per GPU.
    const auto t1 = clk::now();
    SceneData scene;
    std::string scene_desc;
    {
        dsrt_mesh* mesh = nullptr;
        if (dsrt_mesh_load_obj(o.obj_path.c_str(), o.obj_scale, &mesh) == DSRT_OK) {
            DsrtMeshView v;
            dsrt_mesh_view(mesh, &v);
            scene.tris.assign(v.triangles, v.triangles + v.num_triangles);
            scene.mats.assign(v.materials, v.materials + v.num_materials);
            scene.tex.assign(v.textures, v.textures + v.num_textures);
            scene.pool.assign(v.texture_pool, v.texture_pool + v.texture_pool_floats);
            scene_desc = std::to_string(v.num_triangles) + " triangles, " + std::to_string(v.num_materials) + " materials, " +
                         std::to_string(v.num_textures) + " textures from " + o.obj_path;
            if (v.skipped_faces) scene_desc += " (" + std::to_string(v.skipped_faces) + " triangles with bad indices dropped)";
            dsrt_mesh_free(mesh);
        } else {
            scene.tris.resize((size_t)std::max(0, o.synthetic_tris));
            scene.mats.resize(32);
            const int nm = dsrt_synth_mesh((int)scene.tris.size(), 1337u, o.diffuse_only ? 1 : 0, scene.tris.data(),
                                           scene.mats.data(), (int)scene.mats.size());
            if (nm < 0) { std::fprintf(stderr, "synthetic mesh generation failed\n"); return 1; }
            scene.mats.resize((size_t)nm);
            scene_desc = o.obj_path + " not found: " + std::to_string(scene.tris.size()) + " synthetic triangles";
        }
    }
    const auto t2 = clk::now();
    std::cout << "Base ISS + material load time: "
              << std::chrono::duration_cast<std::chrono::milliseconds>(t2 - t1).count() << " ms (" << scene_desc << ")\n";

    // Poses, or the reference's single default pose.
    std::vector<double> poses;
    int np = 0;
    if (!o.pose_file.empty()) {
        np = dsrt_read_pose_file(o.pose_file.c_str(), nullptr, 0);
        if (np > 0) { poses.resize((size_t)np * 9); dsrt_read_pose_file(o.pose_file.c_str(), poses.data(), np); }
    }
    if (np <= 0) {
        std::cout << "No valid pose file found; using single default pose.\n";
        const double def[9] = {0.0, 50.0, 200.0, 0.0, -100.0, 0.0, 0.0, 0.0, 0.0};
        poses.assign(def, def + 9);
        np = 1;
    } else {
        std::cout << "Loaded " << np << " poses.\n";
    }

    std::vector<FrameJob> jobs;
    const int fb = std::max(0, o.frame_begin), fe = o.frame_end < 0 ? np : std::min(np, o.frame_end);
    for (int i = fb; i < fe; ++i) {
        const double* p = &poses[(size_t)i * 9];
        FrameJob j;
        j.index = (size_t)i;
        std::memcpy(j.pose, p, sizeof(j.pose));
        double sep = 0;
        dsrt_pose_to_model_frame(p, j.cam_in_model, j.sun_dir, &sep);
        std::cout << "\n=== Frame " << i << " ===\n";
        std::cout << "Camera world: (" << p[0] << ", " << p[1] << ", " << p[2] << ")\n";
        std::cout << "Model world : (" << p[3] << ", " << p[4] << ", " << p[5] << ")\n";
        std::cout << "Model yaw/pitch/roll (deg): (" << (float)p[6] << ", " << (float)p[7] << ", " << (float)p[8] << ")\n";
        std::cout << "  sep(cam, model) = " << sep << " m\n";
        if (sep < 1.0) {
            std::cout << "  [!] Camera is inside/too close to ISS mesh. Skipping frame.\n";
            continue;
        }
        std::cout << "  sun_dir_model   = (" << j.sun_dir[0] << ", " << j.sun_dir[1] << ", " << j.sun_dir[2] << ")\n";
        jobs.push_back(j);
    }

    std::vector<FrameResult> results(jobs.size());
    int rc = 0;
    const int G = std::max(1, o.gpus);
    std::vector<SetupTimes> times((size_t)G);
    // DSRT_CLI_ISOLATE=1 hides the other GPUs from each worker process.  Off by default: measured on B200 boxes it
    // does not shorten CUDA start-up (2.4 s vs 3.6 s per process at 2 GPUs) and slows NCCL's set-up.
    const bool isolate = std::getenv("DSRT_CLI_ISOLATE") && std::atoi(std::getenv("DSRT_CLI_ISOLATE")) != 0;
    const auto render_start = clk::now();
    if (o.shard == "sample" && G > 1) {
#ifdef DSRT_WITH_NCCL
        rc = sample_sharded(o, scene, jobs, results, times, isolate);
#else
        std::fprintf(stderr, "--shard sample needs a build with NCCL\n");
        rc = 1;
#endif
    } else {
        std::atomic<int> errors(0);
        if (G == 1) {
            frame_worker(0, 0, o, scene, jobs, 0, 1, results, errors, times[0]);
        } else {
            // one worker process per GPU; results come back through a pipe
            std::cout.flush();
            std::fflush(stdout);
            std::vector<pid_t> pids((size_t)G, -1);
            std::vector<int> rfd((size_t)G, -1);
            for (int g = 0; g < G; ++g) {
                int fd[2];
                if (pipe(fd) != 0) { std::perror("pipe"); return 1; }
                const pid_t pid = fork();
                if (pid < 0) { std::perror("fork"); return 1; }
                if (pid == 0) {
                    close(fd[0]);
                    for (int q = 0; q < g; ++q) close(rfd[(size_t)q]);
                    const int device = claim_gpu(g, isolate);
                    frame_worker(device, g, o, scene, jobs, g, G, results, errors, times[(size_t)g]);
                    std::cout.flush();
                    bool ok = write_all(fd[1], &times[(size_t)g], sizeof(SetupTimes));
                    for (size_t k = (size_t)g; k < jobs.size() && ok; k += (size_t)G) {
                        const Wire w{k, results[k]};
                        ok = write_all(fd[1], &w, sizeof(w));
                    }
                    close(fd[1]);
                    _exit((errors || !ok) ? 1 : 0);
                }
                close(fd[1]);
                pids[(size_t)g] = pid;
                rfd[(size_t)g] = fd[0];
            }
            for (int g = 0; g < G; ++g) {
                if (!read_all(rfd[(size_t)g], &times[(size_t)g], sizeof(SetupTimes))) ++errors;
                Wire w;
                while (read_all(rfd[(size_t)g], &w, sizeof(w)))
                    if (w.k < results.size()) results[w.k] = w.r;
                close(rfd[(size_t)g]);
                int status = 0;
                waitpid(pids[(size_t)g], &status, 0);
                if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) ++errors;
            }
        }
        rc = errors ? 1 : 0;
    }

    const auto total_end = clk::now();
    const double total_s = std::chrono::duration<double>(total_end - total_start).count();
    // upload + BVH build + every frame rendered, read back and written (PPM + PNG) on all GPUs
    const double render_wall_s = std::chrono::duration<double>(total_end - render_start).count();
    double rays = 0, ms = 0;
    size_t done = 0;
    for (const auto& r : results) { if (r.done) { ++done; rays += r.rays; ms += r.device_ms; } }
    SetupTimes worst;
    for (const auto& t : times) {
        worst.create_s = std::max(worst.create_s, t.create_s); worst.upload_s = std::max(worst.upload_s, t.upload_s);
        worst.first_frame_s = std::max(worst.first_frame_s, t.first_frame_s); worst.loop_s = std::max(worst.loop_s, t.loop_s);
        worst.teardown_s = std::max(worst.teardown_s, t.teardown_s);
    }
    std::cout << "\nPer-GPU wall clock (max over GPUs): context " << worst.create_s << " s, scene upload + BVH " << worst.upload_s
              << " s, first frame " << worst.first_frame_s << " s, frame loop " << worst.loop_s << " s, teardown " << worst.teardown_s << " s\n";
    std::cout << "Total runtime: " << (long long)total_s << " s\n";
    if (done) std::cout << "Frames: " << done << ", device time " << ms << " ms, " << rays / (ms * 1e3) << " Mrays/s per GPU-time\n";
    if (!o.stats_json.empty()) {
        if (FILE* f = std::fopen(o.stats_json.c_str(), "w")) {
            std::fprintf(f, "{\"frames\": %zu, \"gpus\": %d, \"shard\": \"%s\", \"total_s\": %.3f, \"render_wall_s\": %.3f, "
                            "\"frames_per_s\": %.4f, \"device_ms\": %.3f, \"rays\": %.0f, \"width\": %d, \"height\": %d, \"spp\": %d, "
                            "\"max_depth\": %d, \"triangles\": %zu, \"png\": %s, \"context_s\": %.3f, \"upload_s\": %.3f, \"first_frame_s\": %.3f, "
                            "\"frame_loop_s\": %.3f, \"teardown_s\": %.3f, \"loop_frames_per_s\": %.4f}\n",
                         done, G, o.shard.c_str(), total_s, render_wall_s, render_wall_s > 0 ? done / render_wall_s : 0.0, ms, rays,
                         o.width, o.height, o.spp, o.max_depth, scene.tris.size(), o.png ? "true" : "false",
                         worst.create_s, worst.upload_s, worst.first_frame_s, worst.loop_s, worst.teardown_s,
                         worst.loop_s > 0 ? done / worst.loop_s : 0.0);
            std::fclose(f);
        }
    }
    std::cout << (o.upscale ? "" : "Upscaling disabled (use --upscale to enable).\n");
    std::cout << "Done.\n";
    return rc;
}
