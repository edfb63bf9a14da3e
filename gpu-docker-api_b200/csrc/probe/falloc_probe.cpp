// falloc_probe.cpp -- one-off host measurement (not part of libvmig): how fast can NEW tmpfs files
// be filled when page allocation (fallocate) is split from the copy (mmap memcpy, no inode lock)?
//   g++ -O2 -pthread -o falloc_probe falloc_probe.cpp ; ./falloc_probe /dev/shm/vmig_fprobe
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
static void par(int T, const std::function<void(int)>& f) { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(f, t); for (auto& x : th) x.join(); }
int main(int argc, char** argv) {
    std::string dir = argc > 1 ? argv[1] : "/dev/shm/vmig_fprobe";
    mkdir(dir.c_str(), 0755);
    const size_t G = 1ull << 30, CH = 4 << 20;
    char* srcbuf = (char*)aligned_alloc(4096, 256 << 20); memset(srcbuf, 7, 256 << 20);
    auto path = [&](int i) { return dir + "/f" + std::to_string(i); };
    // 1. fallocate rate, one new 2 GiB file
    { int fd = open(path(0).c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); double t0 = now();
      if (fallocate(fd, 0, 0, 2 * G)) perror("fallocate"); double dt = now() - t0;
      printf("fallocate 2 GiB new file, 1 thread: %.2f GB/s\n", 2 * G / dt / 1e9);
      // pwrite into fallocated pages, 1 thread
      t0 = now(); for (size_t o = 0; o < 2 * G; o += CH) if (pwrite(fd, srcbuf + (o % (256 << 20)), CH, o) != (ssize_t)CH) perror("pwrite");
      printf("pwrite into fallocated (never written) pages, 1T: %.2f GB/s\n", 2 * G / (now() - t0) / 1e9);
      close(fd); unlink(path(0).c_str()); }
    // 2. fallocate whole file then mmap memcpy with T threads
    for (int T : {1, 4, 8}) {
        int fd = open(path(0).c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); double t0 = now();
        if (fallocate(fd, 0, 0, 2 * G)) perror("fallocate"); double ta = now() - t0;
        char* m = (char*)mmap(0, 2 * G, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0); std::atomic<size_t> next{0};
        t0 = now(); par(T, [&](int) { for (;;) { size_t o = next.fetch_add(CH); if (o >= 2 * G) break; memcpy(m + o, srcbuf + (o % (256 << 20)), CH); } });
        double tc = now() - t0;
        printf("fallocate (%.2f GB/s) then mmap memcpy %dT: copy %.2f GB/s, end-to-end %.2f GB/s\n", 2 * G / ta / 1e9, T, 2 * G / tc / 1e9, 2 * G / (ta + tc) / 1e9);
        munmap(m, 2 * G); close(fd); unlink(path(0).c_str());
    }
    // 3. overlapped: allocator thread fallocates 32 MiB steps ahead, T copiers mmap-memcpy behind it (never ahead)
    for (int T : {2, 4}) {
        int fd = open(path(0).c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); if (ftruncate(fd, 2 * G)) return 1;
        char* m = (char*)mmap(0, 2 * G, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        std::atomic<size_t> allocated{0}, next{0}; double t0 = now();
        std::thread alloc([&] { for (size_t o = 0; o < 2 * G; o += 32 << 20) { if (fallocate(fd, 0, o, 32 << 20)) perror("fallocate"); allocated.store(o + (32 << 20)); } });
        par(T, [&](int) { for (;;) { size_t o = next.fetch_add(CH); if (o >= 2 * G) break; while (allocated.load() < o + CH) sched_yield(); memcpy(m + o, srcbuf + (o % (256 << 20)), CH); } });
        alloc.join();
        printf("ONE file: allocator thread + %d mmap copiers overlapped: %.2f GB/s\n", T, 2 * G / (now() - t0) / 1e9);
        munmap(m, 2 * G); close(fd); unlink(path(0).c_str());
    }
    // 4. ten files at once: (a) one pwrite thread per file; (b) allocator + 2 copiers per file
    { const int F = 10; double t0 = now();
      par(F, [&](int f) { int fd = open(path(f).c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); for (size_t o = 0; o < G; o += CH) if (pwrite(fd, srcbuf + (o % (256 << 20)), CH, o) != (ssize_t)CH) perror("pw"); close(fd); });
      printf("10 files x 1 GiB, one pwrite thread per file: %.2f GB/s\n", F * G / (now() - t0) / 1e9);
      for (int f = 0; f < F; f++) unlink(path(f).c_str());
      for (int C : {1, 2, 3}) {
        t0 = now();
        par(F, [&](int f) {
            int fd = open(path(f).c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); if (ftruncate(fd, G)) return;
            char* m = (char*)mmap(0, G, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            std::atomic<size_t> allocated{0}, next{0};
            std::thread alloc([&] { for (size_t o = 0; o < G; o += 32 << 20) { if (fallocate(fd, 0, o, 32 << 20)) perror("fallocate"); allocated.store(o + (32 << 20)); } });
            par(C, [&](int) { for (;;) { size_t o = next.fetch_add(CH); if (o >= G) break; while (allocated.load() < o + CH) sched_yield(); memcpy(m + o, srcbuf + (o % (256 << 20)), CH); } });
            alloc.join(); munmap(m, G); close(fd); });
        printf("10 files x 1 GiB, allocator + %d mmap copiers per file: %.2f GB/s\n", C, F * G / (now() - t0) / 1e9);
        for (int f = 0; f < F; f++) unlink(path(f).c_str());
      }
      // (c) mmap copiers only, no fallocate (page-fault allocation), 3 per file
      t0 = now();
      par(F, [&](int f) { int fd = open(path(f).c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); if (ftruncate(fd, G)) return;
          char* m = (char*)mmap(0, G, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0); std::atomic<size_t> next{0};
          par(3, [&](int) { for (;;) { size_t o = next.fetch_add(CH); if (o >= G) break; memcpy(m + o, srcbuf + (o % (256 << 20)), CH); } });
          munmap(m, G); close(fd); });
      printf("10 files x 1 GiB, 3 mmap copiers per file, no fallocate: %.2f GB/s\n", F * G / (now() - t0) / 1e9);
      for (int f = 0; f < F; f++) unlink(path(f).c_str());
    }
    rmdir(dir.c_str());
    return 0;
}
