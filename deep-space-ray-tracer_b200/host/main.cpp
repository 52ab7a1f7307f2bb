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
// silently renders an empty scene when it is missing.  Here --obj <path> (default: that same
// path) is loaded by dsrt_mesh_load_obj (OBJ + MTL + map_Kd textures, host/obj_loader.cpp).
// An --obj that was asked for and cannot be opened is an error (exit 1); the missing DEFAULT
// path renders the reference's empty scene, with a warning.  --synthetic_tris N replaces the
// OBJ by the deterministic synthetic station (the NASA model is not redistributable with
// either tree).
//
// Several GPUs are driven from ONE process, one worker thread per GPU, each with its own
// dsrt_ctx (round 1 forked one process per GPU: eight CUDA runtimes starting at once took
// ~11 s on a B200 box against ~2.4 s for one; here the runtime starts once and only the
// per-device contexts are created in parallel).  Frames are handed out dynamically from one
// atomic counter -- late poses of the shipped sequence cost ~3x the early ones -- and each
// worker keeps TWO frames in flight on two frame slots (dsrt_render_frame_async), so the
// latency-bound end of frame k, its tone map and its copy to the host run under the camera
// rays of frame k+1.  There is no collective in this mode.
//
// Frames are written by a writer thread per GPU (PPM, then the in-process PNG encoder that
// replaces the reference's `magick` subprocess, main.cpp:28-36), so a GPU renders on
// while finished frames are being compressed.
//
// Extra, optional flags (defaults reproduce the reference): --width --height --spp
// --max_depth --seed --obj <path> --obj_scale S --synthetic_tris N --diffuse_only --gpus G
// --shard frame|sample --frames a:b --rng philox|lcg --no_png --stats_json <file>
// --dump_accum (also writes <frame>.accum, the raw float accumulation buffer) --sync_frames
// (one frame at a time per GPU, as round 1 did) --resume (keep the output directory and render
// only the frames whose files are missing).
#include "dsrt.h"

#include <algorithm>
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
    bool upscale = false, png = true, diffuse_only = false, dump_accum = false, sync_frames = false, resume = false;
    bool obj_given = false;
    int width = 800, height = 450, spp = 1000, max_depth = 50;
    int synthetic_tris = -1, gpus = 1, frame_begin = 0, frame_end = -1;      // -1: no synthetic stand-in
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
        else if (const char* v = next("--obj")) { o.obj_path = v; o.obj_given = true; }
        else if (a == "--dump_accum") o.dump_accum = true;
        else if (a == "--sync_frames") o.sync_frames = true;
        else if (a == "--resume") o.resume = true;
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
    double sep;              // camera-model separation [m]: the closer, the more of the frame the station covers
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
// Frames travel in four page-locked buffers that circulate between the render loop (acquire ->
// read back -> push; two frames are in flight per GPU) and the writer (save -> release), so at most two
// frames wait to be written and the producer blocks beyond that.  Created after the ctx (page-locked memory needs CUDA).
class FrameWriter {
public:
    FrameWriter(const Options& o, std::vector<FrameResult>& results)
        : o_(o), results_(results), bytes_((size_t)o.width * o.height * 3) {
        for (int k = 0; k < 4; ++k) {
            void* p = nullptr;
            if (dsrt_host_alloc(&p, bytes_) != DSRT_OK) {
                p = std::malloc(bytes_);
                if (!p) { std::fprintf(stderr, "out of host memory for the frame buffers\n"); std::abort(); }
                pageable_.push_back(p);
            }
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

// One GPU of a frame-parallel run: takes frames from the shared counter until none are left and keeps two of
// them in flight on two frame slots of its ctx.
void frame_worker(int device, const Options& o, const SceneData& scene, const std::vector<FrameJob>& jobs,
                  std::atomic<size_t>& next, std::vector<FrameResult>& results, std::atomic<int>& errors, SetupTimes& times) {
    dsrt_ctx* ctx = nullptr;
    auto t0 = std::chrono::high_resolution_clock::now();
    if (dsrt_create(&ctx, device) != DSRT_OK) {
        std::lock_guard<std::mutex> l(g_print);
        std::fprintf(stderr, "GPU %d: %s\n", device, dsrt_last_error(nullptr));
        ++errors;
        return;
    }
    times.create_s = seconds_since(t0);
    t0 = std::chrono::high_resolution_clock::now();
    if (upload(ctx, scene) != DSRT_OK) {
        std::lock_guard<std::mutex> l(g_print);
        std::fprintf(stderr, "GPU %d: %s\n", device, dsrt_last_error(ctx));
        ++errors;
        dsrt_destroy(ctx);
        return;
    }
    times.upload_s = seconds_since(t0);
    {
        DsrtSceneStats st;
        dsrt_scene_stats(ctx, &st);
        std::lock_guard<std::mutex> l(g_print);
        std::cout << "GPU " << device << " scene build time: " << st.build_ms << " ms (" << st.num_wide_nodes
                  << " wide nodes, upload " << st.upload_ms << " ms)\n";
    }
    const DsrtRenderParams params = make_params(o);
    FrameWriter writer(o, results);
    const int nslots = (o.sync_frames || o.dump_accum) ? 1 : 2;
    struct InFlight { bool busy = false; size_t k = 0; unsigned char* rgb = nullptr; } fl[2];
    bool first = true;
    t0 = std::chrono::high_resolution_clock::now();

    auto retire = [&](int s) {
        if (!fl[s].busy) return;
        fl[s].busy = false;
        const size_t k = fl[s].k;
        DsrtFrameStats fs{};
        if (dsrt_frame_wait(ctx, s, &fs) != DSRT_OK) {
            std::lock_guard<std::mutex> l(g_print);
            std::fprintf(stderr, "frame %zu on GPU %d: %s\n", jobs[k].index, device, dsrt_last_error(ctx));
            ++errors;
            writer.release(fl[s].rgb);
            return;                         // no stale file is renamed into place
        }
        if (first) { times.first_frame_s = seconds_since(t0); first = false; }
        results[k].device_ms = fs.device_ms;
        results[k].rays = (double)(fs.primary_rays + fs.bounce_rays + fs.shadow_rays);
        if (o.dump_accum) {
            std::vector<float> acc((size_t)o.width * o.height * 3);
            char name[64];
            std::snprintf(name, sizeof(name), "/frame_%04zu.accum", jobs[k].index);
            FILE* f = dsrt_readback_accum(ctx, acc.data(), acc.size() * sizeof(float)) == DSRT_OK ? std::fopen((o.out_dir + name).c_str(), "wb") : nullptr;
            if (f) { std::fwrite(acc.data(), sizeof(float), acc.size(), f); std::fclose(f); } else ++errors;
        }
        writer.push(k, jobs[k].index, fl[s].rgb);
        std::lock_guard<std::mutex> l(g_print);
        char name[64];
        std::snprintf(name, sizeof(name), "frame_%04zu.png", jobs[k].index);
        std::cout << "Saved " << o.out_dir << "/" << name << "  [GPU " << device << ", " << fs.device_ms << " ms, "
                  << results[k].rays / (fs.device_ms * 1e3) << " Mrays/s]\n";
    };

    int s = 0;
    for (;;) {
        const size_t k = next.fetch_add(1);
        if (k >= jobs.size()) break;
        retire(s);                          // the slot's previous frame (the older of the two in flight)
        const FrameJob& j = jobs[k];
        unsigned char* rgb = writer.acquire();
        DsrtCamera cam;
        dsrt_camera_look_at_origin(j.cam_in_model, o.vfov, o.width, o.height, o.spp, o.max_depth, &cam);
        int rc = dsrt_render_frame_async(ctx, s, &cam, j.sun_dir, kSunRadiance, 1, &params, o.seed, 0, o.spp, nullptr, 0);
        if (rc == DSRT_OK) {
            fl[s].busy = true; fl[s].k = k; fl[s].rgb = rgb;
            rc = dsrt_readback_rgb8_async(ctx, s, o.spp, params.gamma, rgb, writer.bytes());
        }
        if (rc != DSRT_OK) {
            {
                std::lock_guard<std::mutex> l(g_print);
                std::fprintf(stderr, "frame %zu on GPU %d: %s\n", j.index, device, dsrt_last_error(ctx));
            }
            ++errors;
            if (fl[s].busy) { dsrt_frame_wait(ctx, s, nullptr); fl[s].busy = false; }
            writer.release(rgb);
            continue;
        }
        if (nslots == 1) retire(s); else s ^= 1;
    }
    retire(s);
    retire(s ^ 1);
    writer.finish();
    times.loop_s = seconds_since(t0);
    t0 = std::chrono::high_resolution_clock::now();
    dsrt_destroy(ctx);
    times.teardown_s = seconds_since(t0);
}

#ifdef DSRT_WITH_NCCL
// One frame at a time, its samples split over the GPUs, one worker thread per GPU.  Worker g renders the global
// sample indices [spp*g/G, spp*(g+1)/G) of every pixel; one ncclReduce(sum) of the float accumulation buffer over
// NVLink brings the frame to worker 0, which tone-maps and writes it.
// A worker whose render fails still takes part in the reduce with a zeroed buffer and flags the frame, so the
// collective neither dead-locks nor sums a partial frame; a failing collective aborts every communicator.
struct SampleShared {
    std::vector<ncclComm_t> comms;
    std::vector<std::atomic<int>> comm_ready;
    std::vector<std::atomic<int>> frame_failed;
    std::atomic<int> abort_all{0};
    std::vector<std::vector<FrameResult>> per;       // [gpu][frame]
    SampleShared(int G, size_t frames) : comms((size_t)G, nullptr), comm_ready((size_t)G), frame_failed(frames), per((size_t)G, std::vector<FrameResult>(frames)) {
        for (auto& a : comm_ready) a = 0;
        for (auto& a : frame_failed) a = 0;
    }
    void abort() {
        if (abort_all.exchange(1)) return;
        for (size_t g = 0; g < comms.size(); ++g)
            if (comm_ready[g].load()) ncclCommAbort(comms[g]);
    }
};

void sample_worker(int g, int G, const ncclUniqueId& id, SampleShared& sh, const Options& o, const SceneData& scene,
                   const std::vector<FrameJob>& jobs, std::atomic<int>& errors, SetupTimes& times) {
    auto t0 = std::chrono::high_resolution_clock::now();
    dsrt_ctx* ctx = nullptr;
    bool ok = dsrt_create(&ctx, g) == DSRT_OK;
    if (!ok) { std::lock_guard<std::mutex> l(g_print); std::fprintf(stderr, "GPU %d: %s\n", g, dsrt_last_error(nullptr)); }
    times.create_s = seconds_since(t0);
    t0 = std::chrono::high_resolution_clock::now();
    if (ok && upload(ctx, scene) != DSRT_OK) {
        std::lock_guard<std::mutex> l(g_print);
        std::fprintf(stderr, "GPU %d: %s\n", g, dsrt_last_error(ctx));
        ok = false;
    }
    // every worker joins the communicator set-up, also one that has already failed: its peers are waiting in it
    cudaSetDevice(g);
    ncclComm_t comm = nullptr;
    if (ncclCommInitRank(&comm, G, id, g) != ncclSuccess) {
        std::lock_guard<std::mutex> l(g_print);
        std::fprintf(stderr, "GPU %d: ncclCommInitRank failed\n", g);
        ++errors;
        sh.abort();
        if (ctx) dsrt_destroy(ctx);
        return;
    }
    sh.comms[(size_t)g] = comm;
    sh.comm_ready[(size_t)g] = 1;
    times.upload_s = seconds_since(t0);
    if (!ok) ++errors;

    const DsrtRenderParams params = make_params(o);
    std::vector<FrameResult>& results = sh.per[(size_t)g];
    {
        FrameWriter writer(o, results);
        t0 = std::chrono::high_resolution_clock::now();
        for (size_t k = 0; k < jobs.size() && !sh.abort_all.load(); ++k) {
            const FrameJob& j = jobs[k];
            DsrtCamera cam;
            dsrt_camera_look_at_origin(j.cam_in_model, o.vfov, o.width, o.height, o.spp, o.max_depth, &cam);
            // global sample indices stay global: the jitter depends on s/spp (gpu_render.cu:995-996)
            const int s0 = (int)((long long)o.spp * g / G), s1 = (int)((long long)o.spp * (g + 1) / G);
            bool rendered = ok && dsrt_render_frame(ctx, &cam, j.sun_dir, kSunRadiance, 1, &params, o.seed, s0, s1, nullptr, 0) == DSRT_OK;
            DsrtFrameStats fs{};
            if (rendered) dsrt_frame_stats(ctx, &fs);
            if (!rendered) {
                {
                    std::lock_guard<std::mutex> l(g_print);
                    std::fprintf(stderr, "frame %zu on GPU %d: %s\n", j.index, g, ctx ? dsrt_last_error(ctx) : "no context");
                }
                ++errors;
                sh.frame_failed[k] = 1;
                if (!ctx || dsrt_clear_accum(ctx, o.width, o.height) != DSRT_OK) { sh.abort(); break; }
            }
            if (dsrt_reduce_accum(ctx, comm, 0) != DSRT_OK) {
                if (!sh.abort_all.load()) {
                    std::lock_guard<std::mutex> l(g_print);
                    std::fprintf(stderr, "GPU %d: %s\n", g, dsrt_last_error(ctx));
                }
                ++errors;
                sh.abort();
                break;
            }
            results[k].device_ms = rendered ? fs.device_ms : 0.0;
            results[k].rays = rendered ? (double)(fs.primary_rays + fs.bounce_rays + fs.shadow_rays) : 0.0;
            if (g == 0) {
                if (sh.frame_failed[k].load()) continue;             // a shard is missing: no file for this frame
                if (o.dump_accum) {
                    std::vector<float> acc((size_t)o.width * o.height * 3);
                    char name[64];
                    std::snprintf(name, sizeof(name), "/frame_%04zu.accum", j.index);
                    FILE* f = dsrt_readback_accum(ctx, acc.data(), acc.size() * sizeof(float)) == DSRT_OK ? std::fopen((o.out_dir + name).c_str(), "wb") : nullptr;
                    if (f) { std::fwrite(acc.data(), sizeof(float), acc.size(), f); std::fclose(f); } else ++errors;
                }
                unsigned char* rgb = writer.acquire();
                if (dsrt_readback_rgb8(ctx, o.spp, params.gamma, rgb, writer.bytes()) != DSRT_OK) { writer.release(rgb); ++errors; continue; }
                writer.push(k, j.index, rgb);
                std::lock_guard<std::mutex> l(g_print);
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
    if (!sh.abort_all.load()) ncclCommDestroy(comm);
    if (ctx) dsrt_destroy(ctx);
    times.teardown_s = seconds_since(t0);
}

int sample_sharded(const Options& o, const SceneData& scene, const std::vector<FrameJob>& jobs,
                   std::vector<FrameResult>& results, std::vector<SetupTimes>& times) {
    const int G = o.gpus;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) { std::fprintf(stderr, "ncclGetUniqueId failed\n"); return 1; }
    SampleShared sh(G, jobs.size());
    std::atomic<int> errors(0);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g)
        th.emplace_back([&, g] { sample_worker(g, G, id, sh, o, scene, jobs, errors, times[(size_t)g]); });
    for (auto& t : th) t.join();
    for (size_t k = 0; k < jobs.size(); ++k) {
        bool done = !sh.frame_failed[k].load();
        double mx = 0, total = 0;
        for (int g = 0; g < G; ++g) {
            done = done && sh.per[(size_t)g][k].done;
            mx = std::max(mx, sh.per[(size_t)g][k].device_ms);
            total += sh.per[(size_t)g][k].rays;
        }
        results[k].done = done; results[k].device_ms = mx; results[k].rays = total;
    }
    return errors.load() ? 1 : 0;
}
#endif

}  // namespace

int main(int argc, char** argv) {
    using clk = std::chrono::high_resolution_clock;
    Options o;
    parse_args(argc, argv, o);
    // --resume: frames are independent, so a sequence that was interrupted continues with the frames whose files are
    // missing; the output directory is then NOT emptied (the reference always empties it, main.cpp:41-50)
    if (o.resume) fs::create_directories(o.out_dir); else prepare_output_dir(o.out_dir);

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
        if (o.synthetic_tris >= 0) {
            scene.tris.resize((size_t)o.synthetic_tris);
            scene.mats.resize(32);
            const int nm = dsrt_synth_mesh((int)scene.tris.size(), 1337u, o.diffuse_only ? 1 : 0, scene.tris.data(),
                                           scene.mats.data(), (int)scene.mats.size());
            if (nm < 0) { std::fprintf(stderr, "synthetic mesh generation failed\n"); return 1; }
            scene.mats.resize((size_t)nm);
            scene_desc = std::to_string(scene.tris.size()) + " synthetic triangles (--synthetic_tris)";
        } else if (dsrt_mesh_load_obj(o.obj_path.c_str(), o.obj_scale, &mesh) == DSRT_OK) {
            DsrtMeshView v;
            dsrt_mesh_view(mesh, &v);
            scene.tris.assign(v.triangles, v.triangles + v.num_triangles);
            scene.mats.assign(v.materials, v.materials + v.num_materials);
            scene.tex.assign(v.textures, v.textures + v.num_textures);
            scene.pool.assign(v.texture_pool, v.texture_pool + v.texture_pool_floats);
            scene_desc = std::to_string(v.num_triangles) + " triangles, " + std::to_string(v.num_materials) + " materials, " +
                         std::to_string(v.num_textures) + " textures from " + o.obj_path;
            if (v.skipped_faces) scene_desc += " (" + std::to_string(v.skipped_faces) + " triangles with bad indices dropped)";
            if (v.textures_unsupported) {
                // the reference decodes map_Kd files with stb_image (GIF, PSD, HDR ... as well); a texture that exists
                // but that this loader cannot decode would silently change the picture, so it is an error, not a white
                // stand-in (a MISSING file is white in the reference too)
                std::fprintf(stderr, "%d texture file(s) of %s exist but could not be decoded (supported: PNG, JPEG, PNM, BMP, TGA); "
                                     "convert them, or set DSRT_ALLOW_MISSING_TEXTURES=1 to render them white\n",
                             v.textures_unsupported, o.obj_path.c_str());
                if (!std::getenv("DSRT_ALLOW_MISSING_TEXTURES")) { dsrt_mesh_free(mesh); return 1; }
            }
            dsrt_mesh_free(mesh);
        } else if (o.obj_given) {
            std::fprintf(stderr, "cannot open OBJ %s\n", o.obj_path.c_str());
            return 1;
        } else {
            // the reference's behaviour for its hard-coded path (triangle_mesh.h:172-173): an empty mesh, black frames
            std::fprintf(stderr, "WARNING: %s not found: rendering an EMPTY scene as the reference does "
                                 "(give --obj <path>, or --synthetic_tris N for the synthetic station)\n", o.obj_path.c_str());
            scene_desc = o.obj_path + " not found: empty scene";
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
        j.sep = sep;
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
        if (o.resume) {
            char name[64];
            std::snprintf(name, sizeof(name), "/frame_%04zu.ppm", j.index);
            std::error_code ec;
            const auto sz = fs::file_size(o.out_dir + name, ec);
            std::snprintf(name, sizeof(name), "/frame_%04zu.png", j.index);
            if (!ec && sz >= (uintmax_t)o.width * o.height * 3 && (!o.png || fs::exists(o.out_dir + name))) {
                std::cout << "  already rendered (--resume)\n";
                continue;
            }
        }
        jobs.push_back(j);
    }

    std::vector<FrameResult> results(jobs.size());
    int rc = 0;
    const int G = std::max(1, o.gpus);
    // Several GPUs, frame-parallel: hand the frames out most expensive first (longest-processing-time order).  The cost
    // of a frame grows with the share of the image the station covers, i.e. as the camera closes in (in the shipped
    // sequence the last poses cost ~6x the first ones); taken in file order, the dearest frames would start last and
    // the other GPUs would wait for them.  Files are named by pose index, so the output does not depend on the order.
    if (G > 1 && o.shard != "sample")
        std::stable_sort(jobs.begin(), jobs.end(), [](const FrameJob& a, const FrameJob& b) { return a.sep < b.sep; });
    std::vector<SetupTimes> times((size_t)G);
    // The CUDA driver initialises every VISIBLE device when the first context is created (~0.8 s per B200 on an
    // 8-GPU node, measured: 6.3 s with eight visible, whether one or eight are used), so only the devices this run
    // uses stay visible -- before the first CUDA call of the process.
    {
        std::string keep;
        const char* vis = std::getenv("CUDA_VISIBLE_DEVICES");
        if (vis && *vis) {
            int n = 0;
            for (const char* c = vis; *c && n < G; ++c) { if (*c == ',') { if (++n == G) break; } keep.push_back(*c); }
        } else {
            for (int g = 0; g < G; ++g) keep += (g ? "," : "") + std::to_string(g);
        }
        if (!keep.empty()) setenv("CUDA_VISIBLE_DEVICES", keep.c_str(), 1);
    }
    const auto render_start = clk::now();
    if (jobs.empty()) {
        std::cout << "Nothing to render.\n";
    } else if (o.shard == "sample" && G > 1) {
#ifdef DSRT_WITH_NCCL
        rc = sample_sharded(o, scene, jobs, results, times);
#else
        std::fprintf(stderr, "--shard sample needs a build with NCCL\n");
        rc = 1;
#endif
    } else {
        std::atomic<int> errors(0);
        std::atomic<size_t> next(0);
        if (G == 1) {
            frame_worker(0, o, scene, jobs, next, results, errors, times[0]);
        } else {
            std::vector<std::thread> th;
            for (int g = 0; g < G; ++g)
                th.emplace_back([&, g] { frame_worker(g, o, scene, jobs, next, results, errors, times[(size_t)g]); });
            for (auto& t : th) t.join();
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
