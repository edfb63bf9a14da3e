// hostcopy_probe.cpp -- host-path experiment (not part of libvmig; no GPU): the ceiling of ANY copy that goes
// through user space on this box.  T threads, one file pair each: pread a chunk of a tmpfs file into a private
// buffer, pwrite it to (a) a new tmpfs file, (b) a file whose pages already exist.  Chunk 4 MiB (a libvmig
// block; falls out of L2) or 256 KiB (cache resident, what tar's pipe does).  Tells whether libvmig's
// ~20-25 GiB/s end-to-end plateau is the page cache itself or the DMA traffic next to it.
//   g++ -O2 -o hostcopy_probe hostcopy_probe.cpp -lpthread ; ./hostcopy_probe /dev/shm/vmig_hcp [MiB per file=512] [split]
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>
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
    std::string dir = argc > 1 ? argv[1] : "/dev/shm/vmig_hcp";
    const int TMAX = 64; const size_t FB = (argc > 2 ? (size_t)atoi(argv[2]) : 512ull) << 20;
    mkdir(dir.c_str(), 0755);
    par(TMAX, [&](int t) { std::string p = dir + "/s" + std::to_string(t); int fd = open(p.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
        char* b = (char*)malloc(4 << 20); memset(b, t + 1, 4 << 20);
        for (size_t o = 0; o < FB; o += 4 << 20) if (pwrite(fd, b, 4 << 20, o) < 0) exit(1);
        close(fd); free(b); });
    const bool only_split = argc > 3 && std::string(argv[3]) == "split";     // just the processes-x-threads part
    if (!only_split)
    for (size_t chunk : {(size_t)4 << 20, (size_t)256 << 10})
        for (int T : {4, 8, 16, 24, 32, 48, 64})
            for (int pass = 0; pass < 2; pass++) {      // 0: new destination files, 1: overwrite their pages
                if (pass == 0) for (int t = 0; t < TMAX; t++) unlink((dir + "/d" + std::to_string(t)).c_str());
                double t0 = now();
                par(T, [&](int t) {
                    int s = open((dir + "/s" + std::to_string(t)).c_str(), O_RDONLY);
                    int d = open((dir + "/d" + std::to_string(t)).c_str(), O_RDWR | O_CREAT, 0644);
                    char* b = (char*)aligned_alloc(4096, chunk);
                    for (size_t o = 0; o < FB; o += chunk) { if (pread(s, b, chunk, o) != (ssize_t)chunk) exit(2); if (pwrite(d, b, chunk, o) != (ssize_t)chunk) exit(3); }
                    close(s); close(d); free(b); });
                double dt = now() - t0;
                printf("chunk %4zu KiB  %2d threads  %-9s  %6.2f GiB/s\n", chunk >> 10, T, pass ? "overwrite" : "new files", T * (double)FB / dt / (1 << 30));
                fflush(stdout);
            }
    // same copy, 64 workers, split into P processes x (64/P) threads: tells whether the fall beyond ~24 threads is a
    // per-process effect (one address space: TLB shootdowns, automatic NUMA balancing, fd table) or a box-wide one.
    // Round-1 hint: 8 libvmig processes x 13 copy threads reach 41.8 GiB/s, one process with 8 lanes 17.6 GiB/s.
    for (int P : {1, 2, 4, 8, 16}) {
        const size_t chunk = 4 << 20; const int W = 64, T = W / P;
        for (int t = 0; t < TMAX; t++) unlink((dir + "/d" + std::to_string(t)).c_str());
        double t0 = now();
        std::vector<pid_t> kids;
        for (int p = 0; p < P; p++) {
            pid_t k = fork();
            if (k == 0) {
                par(T, [&](int tt) { const int t = p * T + tt;
                    int s = open((dir + "/s" + std::to_string(t)).c_str(), O_RDONLY);
                    int d = open((dir + "/d" + std::to_string(t)).c_str(), O_RDWR | O_CREAT, 0644);
                    char* b = (char*)aligned_alloc(4096, chunk);
                    for (size_t o = 0; o < FB; o += chunk) { if (pread(s, b, chunk, o) != (ssize_t)chunk) _exit(2); if (pwrite(d, b, chunk, o) != (ssize_t)chunk) _exit(3); }
                    close(s); close(d); free(b); });
                _exit(0);
            }
            kids.push_back(k);
        }
        for (pid_t k : kids) { int st; waitpid(k, &st, 0); }
        double dt = now() - t0;
        printf("chunk 4096 KiB  %2d processes x %2d threads  new files  %6.2f GiB/s\n", P, T, W * (double)FB / dt / (1 << 30)); fflush(stdout);
    }
    // in-kernel copy (one memcpy per byte instead of two): copy_file_range tmpfs -> tmpfs
    if (!only_split)
    for (int T : {8, 16, 32, 64}) {
        for (int t = 0; t < TMAX; t++) unlink((dir + "/d" + std::to_string(t)).c_str());
        double t0 = now();
        par(T, [&](int t) {
            int s = open((dir + "/s" + std::to_string(t)).c_str(), O_RDONLY);
            int d = open((dir + "/d" + std::to_string(t)).c_str(), O_RDWR | O_CREAT, 0644);
            size_t left = FB; while (left) { ssize_t n = copy_file_range(s, nullptr, d, nullptr, left, 0); if (n <= 0) { perror("copy_file_range"); exit(4); } left -= n; }
            close(s); close(d); });
        double dt = now() - t0;
        printf("copy_file_range    %2d threads  new files  %6.2f GiB/s\n", T, T * (double)FB / dt / (1 << 30)); fflush(stdout);
    }
    for (int t = 0; t < TMAX; t++) { unlink((dir + "/s" + std::to_string(t)).c_str()); unlink((dir + "/d" + std::to_string(t)).c_str()); }
    rmdir(dir.c_str());
    return 0;
}
