// Where CUDA start-up time goes on a multi-GPU node (GPU box; not part of the product or of pytest).
//   cuda_startup seq            one process: cuInit, then a primary context on each visible device, one after another
//   cuda_startup par            one process: one thread per visible device creates its context at the same time
//   cuda_startup one <t0>       a context on device 0 of whatever is visible; prints seconds since the epoch time <t0>
// Build: nvcc -O2 -o tests/c/cuda_startup tests/c/cuda_startup.cu -lcuda -lpthread
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() {
    return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
    const char* mode = argc > 1 ? argv[1] : "seq";
    const double t_start = now();
    double t = now();
    CUresult r = cuInit(0);
    std::printf("%s: cuInit %.3f s (rc %d)\n", mode, now() - t, (int)r);
    int n = 0;
    cudaGetDeviceCount(&n);
    std::printf("%s: %d visible devices, %.3f s since start\n", mode, n, now() - t_start);
    if (!std::strcmp(mode, "one")) {
        t = now();
        cudaSetDevice(0);
        cudaFree(0);
        void* p = nullptr;
        cudaMalloc(&p, 1 << 20);
        const double t0 = argc > 2 ? atof(argv[2]) : t_start;
        std::printf("one (CUDA_VISIBLE_DEVICES=%s): context %.3f s, ready %.3f s after the common start\n",
                    getenv("CUDA_VISIBLE_DEVICES") ? getenv("CUDA_VISIBLE_DEVICES") : "", now() - t, now() - t0);
        return 0;
    }
    if (!std::strcmp(mode, "seq")) {
        for (int g = 0; g < n; ++g) {
            t = now();
            cudaSetDevice(g);
            cudaFree(0);
            std::printf("seq: device %d context %.3f s, ready %.3f s since start\n", g, now() - t, now() - t_start);
        }
        return 0;
    }
    std::vector<std::thread> th;
    std::vector<double> ready((size_t)n), took((size_t)n);
    for (int g = 0; g < n; ++g)
        th.emplace_back([g, t_start, &ready, &took] {
            const double t1 = now();
            cudaSetDevice(g);
            cudaFree(0);
            took[(size_t)g] = now() - t1;
            ready[(size_t)g] = now() - t_start;
        });
    for (auto& x : th) x.join();
    for (int g = 0; g < n; ++g) std::printf("par: device %d context %.3f s, ready %.3f s since start\n", g, took[(size_t)g], ready[(size_t)g]);
    return 0;
}
