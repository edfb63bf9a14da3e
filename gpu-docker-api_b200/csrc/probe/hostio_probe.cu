// hostio_probe.cu -- one-off measurement tool (not part of libvmig): which host-side
// data path can feed a B200's PCIe Gen5 link from/to tmpfs?  Results are recorded in
// profiles/ and drive the engine's I/O design (DESIGN.md "host data path").
//   build: nvcc -O2 -o hostio_probe hostio_probe.cu -lpthread
//   run  : ./hostio_probe /dev/shm/vmig_probe [GiB=4]
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <string>
#include <atomic>
#include <functional>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

static void par(int T, const std::function<void(int)>& f) {
    std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(f, t); for (auto& x : th) x.join();
}

int main(int argc, char** argv) {
    std::string dir = argc > 1 ? argv[1] : "/dev/shm/vmig_probe";
    size_t GiB = argc > 2 ? atoi(argv[2]) : 4;
    size_t N = GiB << 30; const size_t CH = 4 << 20;
    mkdir(dir.c_str(), 0755);
    std::string src = dir + "/src.bin";
    // --- create source file (parallel mmap fill)
    { int fd = open(src.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); if (ftruncate(fd, N)) return 1;
      char* m = (char*)mmap(0, N, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      double t0 = now(); par(16, [&](int t) { size_t per = N / 16; memset(m + t * per, 0x5a + t, per); });
      printf("create src via mmap+memset 16T: %.2f GB/s\n", N / (now() - t0) / 1e9); munmap(m, N); close(fd); }
    char* pin; CK(cudaHostAlloc(&pin, N > (2ull << 30) ? (2ull << 30) : N, cudaHostAllocDefault));
    size_t PN = N > (2ull << 30) ? (2ull << 30) : N;
    double t0 = now(); memset(pin, 1, PN); printf("touch pinned: %.2f GB/s\n", PN / (now() - t0) / 1e9);
    char* dev; CK(cudaMalloc(&dev, PN));
    // --- pinned DMA rates
    for (int rep = 0; rep < 2; rep++) {
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); float ms;
        CK(cudaEventRecord(a)); CK(cudaMemcpyAsync(dev, pin, PN, cudaMemcpyHostToDevice)); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b)); printf("H2D pinned %zu MiB: %.2f GB/s\n", PN >> 20, PN / ms / 1e6);
        CK(cudaEventRecord(a)); CK(cudaMemcpyAsync(pin, dev, PN, cudaMemcpyDeviceToHost)); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b)); printf("D2H pinned %zu MiB: %.2f GB/s\n", PN >> 20, PN / ms / 1e6);
        // bidirectional on two streams
        cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
        char* dev2; CK(cudaMalloc(&dev2, PN / 2));
        double w0 = now();
        CK(cudaMemcpyAsync(dev, pin, PN / 2, cudaMemcpyHostToDevice, s1));
        CK(cudaMemcpyAsync(pin + PN / 2, dev2, PN / 2, cudaMemcpyDeviceToHost, s2));
        CK(cudaDeviceSynchronize()); double w = now() - w0;
        printf("bidir H2D+D2H concurrently: %.2f GB/s each direction\n", PN / 2 / w / 1e9);
        CK(cudaFree(dev2));
    }
    // --- pread tmpfs -> pinned, T threads
    for (int T : {1, 4, 8, 16, 32}) {
        int fd = open(src.c_str(), O_RDONLY); std::atomic<size_t> next{0};
        double t0 = now();
        par(T, [&](int) { for (;;) { size_t o = next.fetch_add(CH); if (o >= N) break; if (pread(fd, pin + (o % PN), CH, o) != (ssize_t)CH) { perror("pread"); exit(1);} } });
        printf("pread tmpfs->pinned %2dT: %.2f GB/s\n", T, N / (now() - t0) / 1e9); close(fd);
    }
    // --- pwrite pinned -> tmpfs single new file, T threads (inode lock?)
    for (int T : {1, 4, 16}) {
        std::string d = dir + "/dst_pw.bin"; unlink(d.c_str());
        int fd = open(d.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); std::atomic<size_t> next{0};
        double t0 = now();
        par(T, [&](int) { for (;;) { size_t o = next.fetch_add(CH); if (o >= N) break; if (pwrite(fd, pin + (o % PN), CH, o) != (ssize_t)CH) { perror("pwrite"); exit(1);} } });
        printf("pwrite pinned->tmpfs ONE new file %2dT: %.2f GB/s\n", T, N / (now() - t0) / 1e9); close(fd); unlink(d.c_str());
    }
    // --- pwrite to T distinct new files
    for (int T : {4, 16, 32}) {
        double t0 = now();
        par(T, [&](int t) { std::string d = dir + "/dst_pw" + std::to_string(t); int fd = open(d.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
            size_t per = N / T; for (size_t o = 0; o < per; o += CH) if (pwrite(fd, pin + ((t * per + o) % PN), CH, o) != (ssize_t)CH) { perror("pwrite"); exit(1);} close(fd); });
        printf("pwrite pinned->tmpfs %2d distinct new files: %.2f GB/s\n", T, N / (now() - t0) / 1e9);
        for (int t = 0; t < T; t++) unlink((dir + "/dst_pw" + std::to_string(t)).c_str());
    }
    // --- mmap + memcpy into ONE new file, T threads
    for (int T : {1, 4, 16, 32}) {
        std::string d = dir + "/dst_mm.bin"; unlink(d.c_str());
        int fd = open(d.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); if (ftruncate(fd, N)) return 1;
        char* m = (char*)mmap(0, N, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0); std::atomic<size_t> next{0};
        double t0 = now();
        par(T, [&](int) { for (;;) { size_t o = next.fetch_add(CH); if (o >= N) break; memcpy(m + o, pin + (o % PN), CH); } });
        printf("mmap+memcpy pinned->tmpfs ONE new file %2dT: %.2f GB/s\n", T, N / (now() - t0) / 1e9);
        munmap(m, N); close(fd); unlink(d.c_str());
    }
    // --- mmap + memcpy overwrite EXISTING pages (diff path: in-place) 16T
    { std::string d = dir + "/dst_mm2.bin"; int fd = open(d.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); if (ftruncate(fd, N)) return 1;
      char* m = (char*)mmap(0, N, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      par(16, [&](int t) { size_t per = N / 16; memset(m + t * per, 1, per); });
      std::atomic<size_t> next{0}; double t0 = now();
      par(16, [&](int) { for (;;) { size_t o = next.fetch_add(CH); if (o >= N) break; memcpy(m + o, pin + (o % PN), CH); } });
      printf("mmap+memcpy overwrite existing pages 16T: %.2f GB/s\n", N / (now() - t0) / 1e9);
      // pwrite in place same file 16T
      next = 0; t0 = now();
      par(16, [&](int) { for (;;) { size_t o = next.fetch_add(CH); if (o >= N) break; if (pwrite(fd, pin + (o % PN), CH, o) != (ssize_t)CH) exit(1); } });
      printf("pwrite overwrite existing pages ONE file 16T: %.2f GB/s\n", N / (now() - t0) / 1e9);
      munmap(m, N); close(fd); unlink(d.c_str()); }
    // --- zero-copy: cudaHostRegister of mmap'd tmpfs src (page-cache hot) in 64 MiB pieces, T threads, then DMA
    for (int T : {1, 4, 16}) {
        int fd = open(src.c_str(), O_RDWR); char* m = (char*)mmap(0, N, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        const size_t RC = 64 << 20; std::atomic<size_t> next{0}; std::atomic<int> fails{0};
        double t0 = now();
        par(T, [&](int) { for (;;) { size_t o = next.fetch_add(RC); if (o >= N) break; if (cudaHostRegister(m + o, RC, cudaHostRegisterDefault) != cudaSuccess) fails++; } });
        double treg = now() - t0;
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); float ms = 0;
        if (!fails) { CK(cudaEventRecord(a)); for (size_t o = 0; o < N; o += RC) CK(cudaMemcpyAsync(dev + (o % PN), m + o, RC, cudaMemcpyHostToDevice)); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); }
        t0 = now(); next = 0;
        par(T, [&](int) { for (;;) { size_t o = next.fetch_add(RC); if (o >= N) break; cudaHostUnregister(m + o); } });
        double tun = now() - t0; cudaGetLastError();
        printf("register mmap'd tmpfs SRC (hot) %2dT: reg %.2f GB/s, H2D from it %.2f GB/s, unreg %.2f GB/s, fails=%d\n", T, N / treg / 1e9, ms ? N / ms / 1e6 : 0.0, N / tun / 1e9, fails.load());
        munmap(m, N); close(fd);
    }
    // --- zero-copy dst: new file, ftruncate, mmap, register (faults pages), D2H into it
    for (int T : {1, 4, 16}) {
        std::string d = dir + "/dst_reg.bin"; unlink(d.c_str());
        int fd = open(d.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); if (ftruncate(fd, N)) return 1;
        char* m = (char*)mmap(0, N, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        const size_t RC = 64 << 20; std::atomic<size_t> next{0}; std::atomic<int> fails{0};
        double t0 = now();
        par(T, [&](int) { for (;;) { size_t o = next.fetch_add(RC); if (o >= N) break; if (cudaHostRegister(m + o, RC, cudaHostRegisterDefault) != cudaSuccess) fails++; } });
        double treg = now() - t0; float ms = 0;
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        if (!fails) { CK(cudaEventRecord(a)); for (size_t o = 0; o < N; o += RC) CK(cudaMemcpyAsync(m + o, dev + (o % PN), RC, cudaMemcpyDeviceToHost)); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); }
        t0 = now(); next = 0;
        par(T, [&](int) { for (;;) { size_t o = next.fetch_add(RC); if (o >= N) break; cudaHostUnregister(m + o); } });
        double tun = now() - t0; cudaGetLastError();
        printf("register mmap'd NEW tmpfs DST %2dT: reg(+fault) %.2f GB/s, D2H into it %.2f GB/s, unreg %.2f GB/s, fails=%d\n", T, N / treg / 1e9, ms ? N / ms / 1e6 : 0.0, N / tun / 1e9, fails.load());
        munmap(m, N); close(fd); unlink(d.c_str());
    }
    // --- the reference: tar | tar on a 1-file tree
    { std::string s = dir + "/tsrc", d = dir + "/tdst"; mkdir(s.c_str(), 0755); mkdir(d.c_str(), 0755);
      rename(src.c_str(), (s + "/src.bin").c_str());
      double t0 = now(); std::string cmd = "(cd " + s + "; tar c .) | (cd " + d + "; tar x)"; int rc = system(cmd.c_str());
      printf("reference tar|tar %zu GiB: %.2f GB/s (rc=%d)\n", GiB, N / (now() - t0) / 1e9, rc);
      std::string rm = "rm -rf " + dir; rc = system(rm.c_str()); }
    return 0;
}
