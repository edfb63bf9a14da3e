// populate_probe.cpp -- host measurement (not part of libvmig): can NEW tmpfs files be filled faster
// than one pwrite stream per file (inode lock) by splitting page allocation (fallocate, per file) from
// the copy (memcpy into a pre-populated MAP_SHARED mapping, any thread, no lock)?
//   g++ -O2 -pthread -o populate_probe populate_probe.cpp ; ./populate_probe /dev/shm/vmig_pprobe
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
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void par(int T, const std::function<void(int)>& f) { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(f, t); for (auto& x : th) x.join(); }
int main(int argc, char** argv) {
    std::string dir = argc > 1 ? argv[1] : "/dev/shm/vmig_pprobe";
    mkdir(dir.c_str(), 0755);
    const size_t G = 1ull << 30, CH = 4 << 20, WIN = 64 << 20; const int F = 10;
    const size_t SB = 512 << 20; char* srcbuf = (char*)aligned_alloc(4096, SB); memset(srcbuf, 7, SB);
    auto path = [&](int i) { return dir + "/f" + std::to_string(i); };
    auto cleanup = [&] { for (int f = 0; f < F; f++) unlink(path(f).c_str()); };
    // baseline: one pwrite stream per file
    { double t0 = now();
      par(F, [&](int f) { int fd = open(path(f).c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); for (size_t o = 0; o < G; o += CH) if (pwrite(fd, srcbuf + ((f * G + o) % SB), CH, o) != (ssize_t)CH) perror("pw"); close(fd); });
      printf("baseline 10 files, one pwrite stream each: %.2f GB/s\n", F * G / (now() - t0) / 1e9); cleanup(); }
    // single file: fallocate + MAP_POPULATE cost
    { int fd = open(path(0).c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); if (ftruncate(fd, G)) return 1;
      double t0 = now(); if (fallocate(fd, 0, 0, G)) perror("fallocate"); double ta = now() - t0;
      t0 = now(); char* m = (char*)mmap(0, G, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0); double tp = now() - t0;
      t0 = now(); memcpy(m, srcbuf, SB); memcpy(m + SB, srcbuf, SB); double tc = now() - t0;
      t0 = now(); munmap(m, G); double tu = now() - t0;
      printf("1 GiB file: fallocate %.2f GB/s, mmap(MAP_POPULATE) %.2f GB/s, memcpy 1T into it %.2f GB/s, munmap %.2f GB/s\n", G / ta / 1e9, G / tp / 1e9, G / tc / 1e9, G / tu / 1e9);
      close(fd); cleanup(); }
    // pipeline: per file an allocator thread (fallocate + madvise(MADV_POPULATE_WRITE) per 64 MiB window), W shared copiers
    for (int W : {8, 12, 16, 24}) {
        std::vector<int> fds(F); std::vector<char*> maps(F); std::vector<std::atomic<size_t>> ready(F);
        double t0 = now();
        for (int f = 0; f < F; f++) { fds[f] = open(path(f).c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); if (ftruncate(fds[f], G)) return 1;
            maps[f] = (char*)mmap(0, G, PROT_READ | PROT_WRITE, MAP_SHARED, fds[f], 0); ready[f] = 0; }
        std::atomic<size_t> next{0}; const size_t per_file = G / CH, total = per_file * F;
        std::vector<std::thread> alloc;
        for (int f = 0; f < F; f++) alloc.emplace_back([&, f] { for (size_t o = 0; o < G; o += WIN) {
            if (fallocate(fds[f], 0, o, WIN)) perror("fallocate");
            if (madvise(maps[f] + o, WIN, MADV_POPULATE_WRITE)) { static bool w = false; if (!w) { perror("MADV_POPULATE_WRITE"); w = true; } }
            ready[f].store(o + WIN); } });
        par(W, [&](int) { for (;;) { size_t k = next.fetch_add(1); if (k >= total) break; int f = k % F; size_t o = (k / F) * CH;
            while (ready[f].load() < o + CH) sched_yield(); memcpy(maps[f] + o, srcbuf + ((f * G + o) % SB), CH); } });
        for (auto& t : alloc) t.join();
        double tc = now() - t0;
        double t1 = now(); for (int f = 0; f < F; f++) { munmap(maps[f], G); close(fds[f]); } double tu = now() - t1;
        printf("10 files: per-file allocator (fallocate+POPULATE_WRITE, 64 MiB windows) + %2d shared memcpy threads: %.2f GB/s (+ munmap %.0f ms)\n", W, F * G / tc / 1e9, tu * 1e3);
        cleanup();
    }
    // same without fallocate: populate alone allocates
    { int W = 12; std::vector<int> fds(F); std::vector<char*> maps(F); std::vector<std::atomic<size_t>> ready(F);
      double t0 = now();
      for (int f = 0; f < F; f++) { fds[f] = open(path(f).c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); if (ftruncate(fds[f], G)) return 1;
          maps[f] = (char*)mmap(0, G, PROT_READ | PROT_WRITE, MAP_SHARED, fds[f], 0); ready[f] = 0; }
      std::atomic<size_t> next{0}; const size_t per_file = G / CH, total = per_file * F;
      std::vector<std::thread> alloc;
      for (int f = 0; f < F; f++) alloc.emplace_back([&, f] { for (size_t o = 0; o < G; o += WIN) { if (madvise(maps[f] + o, WIN, MADV_POPULATE_WRITE)) perror("populate"); ready[f].store(o + WIN); } });
      par(W, [&](int) { for (;;) { size_t k = next.fetch_add(1); if (k >= total) break; int f = k % F; size_t o = (k / F) * CH;
          while (ready[f].load() < o + CH) sched_yield(); memcpy(maps[f] + o, srcbuf + ((f * G + o) % SB), CH); } });
      for (auto& t : alloc) t.join();
      printf("10 files: POPULATE_WRITE only (no fallocate) + 12 memcpy threads: %.2f GB/s\n", F * G / (now() - t0) / 1e9);
      for (int f = 0; f < F; f++) { munmap(maps[f], G); close(fds[f]); } cleanup(); }
    rmdir(dir.c_str());
    return 0;
}
