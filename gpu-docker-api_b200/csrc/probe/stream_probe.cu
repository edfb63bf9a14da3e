// stream_probe.cu -- host-path experiment (not part of libvmig): does it pay to keep the pinned staging
// buffers cache-resident?  R reader threads pread 4 MiB blocks of tmpfs files into pinned memory and push
// them to HBM; W writer threads (one per destination file) pull 4 MiB blocks back and pwrite them.  The
// same loop runs with (a) a small private double buffer per thread (fits the LLC) and (b) buffers that
// rotate through a 512 MiB ring (what libvmig's slot rings do).  No kernel runs on the GPU.
//   nvcc -O2 -o stream_probe stream_probe.cu -lpthread ; ./stream_probe /dev/shm/vmig_sprobe
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
static void par(int T, const std::function<void(int)>& f) { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(f, t); for (auto& x : th) x.join(); }
int main(int argc, char** argv) {
    std::string dir = argc > 1 ? argv[1] : "/dev/shm/vmig_sprobe";
    const int F = 10; const size_t G = 1ull << 30, CH = 4 << 20, NB = F * G / CH;
    mkdir(dir.c_str(), 0755);
    std::vector<int> sfd(F);
    { char* buf = (char*)malloc(CH); memset(buf, 5, CH);
      for (int f = 0; f < F; f++) { std::string p = dir + "/s" + std::to_string(f); sfd[f] = open(p.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
          for (size_t o = 0; o < G; o += CH) if (pwrite(sfd[f], buf, CH, o) != (ssize_t)CH) return 1; } free(buf); }
    char* dev; CK(cudaMalloc(&dev, F * G));
    char* ring; const size_t RING = 512ull << 20; CK(cudaHostAlloc(&ring, 2 * RING, cudaHostAllocDefault)); memset(ring, 0, 2 * RING);
    for (int mode = 0; mode < 4; mode++) {
        const bool small = (mode & 1) == 0; const size_t sub = mode >= 2 ? (1 << 20) : CH;   // D2H granularity
        for (int R : {8}) {
            const int W = F;
            std::vector<int> dfd(F);
            for (int f = 0; f < F; f++) { std::string p = dir + "/d" + std::to_string(f); unlink(p.c_str()); dfd[f] = open(p.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); }
            std::atomic<size_t> next{0}; std::vector<std::atomic<int>> landed(NB); for (auto& x : landed) x = 0;
            double t0 = now();
            std::vector<std::thread> th;
            for (int r = 0; r < R; r++) th.emplace_back([&, r] {
                cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
                size_t rot = 0;
                for (;;) { size_t k = next.fetch_add(1); if (k >= NB) break; int f = k % F; size_t o = (k / F) * CH;
                    char* b = small ? ring + (size_t)r * 2 * CH + (rot++ & 1) * CH : ring + ((k * CH) % RING);
                    if (pread(sfd[f], b, CH, o) != (ssize_t)CH) exit(1);
                    CK(cudaMemcpyAsync(dev + f * G + o, b, CH, cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st));
                    landed[k].store(1); }
            });
            for (int w = 0; w < W; w++) th.emplace_back([&, w] {
                cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
                size_t rot = 0;
                for (size_t j = 0; j < G / CH; j++) { size_t k = j * F + w; while (!landed[k].load()) sched_yield();
                    char* b = small ? ring + RING + (size_t)w * 2 * CH + (rot++ & 1) * CH : ring + RING + ((k * CH) % RING);
                    for (size_t s = 0; s < CH; s += sub) {
                        CK(cudaMemcpyAsync(b + s, dev + w * G + j * CH + s, sub, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
                        if (pwrite(dfd[w], b + s, sub, j * CH + s) != (ssize_t)sub) exit(1); } }
            });
            for (auto& t : th) t.join();
            printf("%s staging buffers, D2H+pwrite granularity %zu KiB, %d readers + %d writers: %.2f GB/s\n", small ? "small private (LLC-sized)" : "512 MiB rotating ring   ", sub >> 10, R, W, F * G / (now() - t0) / 1e9);
            for (int f = 0; f < F; f++) close(dfd[f]);
        }
    }
    std::string rm = "rm -rf " + dir; return system(rm.c_str());
}
