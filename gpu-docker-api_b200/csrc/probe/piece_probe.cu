// piece_probe.cu -- host-path experiment (not part of libvmig): prototype of "piecewise staging".
// hostcopy_probe shows that a user-space copy on the bench box is ~30% faster when its bounce buffer stays
// in cache (256 KiB chunks) than when it falls out (4 MiB chunks).  libvmig's pinned rings (2 x 512 MiB) are
// the second kind.  Here every reader owns NP small pinned pieces: pread a piece, cudaMemcpyAsync it into the
// big HBM slot, reuse the piece when its DMA is done (the DMA engine then reads lines that are still in the
// reader's cache, and the piece is never written back); every writer owns two pieces and pulls: D2H of piece
// p+1 runs while piece p is pwritten.  Same loop with the pieces rotating through a 512 MiB ring for
// comparison.  No kernel runs on the GPU.
//   nvcc -O2 -o piece_probe piece_probe.cu -lpthread ; ./piece_probe /dev/shm/vmig_pp
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <string>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e__)); exit(1);} } while (0)
static void par(int T, const std::function<void(int)>& f) { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(f, t); for (auto& x : th) x.join(); }
static std::vector<int> local_cpus() {
    char bus[64] = {0}; std::vector<int> out;
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, 0) != cudaSuccess) return out;
    for (char* p = bus; *p; p++) *p = (char)tolower(*p);
    std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist"); std::string s;
    if (!f || !std::getline(f, s)) return out;
    size_t i = 0;
    while (i < s.size()) { int a = 0, b; while (i < s.size() && isdigit(s[i])) a = a * 10 + (s[i++] - '0'); b = a;
        if (i < s.size() && s[i] == '-') { i++; b = 0; while (i < s.size() && isdigit(s[i])) b = b * 10 + (s[i++] - '0'); }
        for (int c = a; c <= b; c++) out.push_back(c);
        if (i < s.size() && s[i] == ',') i++; else if (i < s.size() && !isdigit(s[i])) break; }
    return out;
}
static std::vector<int> g_local;
static void bind(bool local) {
    if (g_local.empty()) return;
    cpu_set_t cur, set; CPU_ZERO(&set); sched_getaffinity(0, sizeof cur, &cur);
    if (local) { for (int c : g_local) CPU_SET(c, &set); }
    else { for (int c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &cur)) CPU_SET(c, &set); for (int c : g_local) CPU_CLR(c, &set); }
    if (CPU_COUNT(&set)) sched_setaffinity(0, sizeof set, &set);
}
int main(int argc, char** argv) {
    std::string dir = argc > 1 ? argv[1] : "/dev/shm/vmig_pp";
    const int F = 12; const size_t G = 1ull << 30, BLK = 4 << 20, NB = F * G / BLK;
    CK(cudaSetDevice(0)); g_local = local_cpus();
    mkdir(dir.c_str(), 0755);
    std::vector<int> sfd(F);
    par(F, [&](int f) { char* buf = (char*)malloc(BLK); memset(buf, 5 + f, BLK);
        std::string p = dir + "/s" + std::to_string(f); sfd[f] = open(p.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
        for (size_t o = 0; o < G; o += BLK) if (pwrite(sfd[f], buf, BLK, o) != (ssize_t)BLK) exit(1); free(buf); });
    char* dev; CK(cudaMalloc(&dev, F * G));
    char* ring; const size_t RING = 512ull << 20; CK(cudaHostAlloc(&ring, 2 * RING, cudaHostAllocDefault)); memset(ring, 0, 2 * RING);
    struct Cfg { bool priv; size_t ps; int np; int R; };
    std::vector<Cfg> cfgs = { {false, 4u << 20, 0, 8}, {true, 4u << 20, 2, 8}, {true, 1u << 20, 4, 8}, {true, 1u << 20, 2, 8}, {true, 512u << 10, 4, 8},
                              {true, 256u << 10, 4, 8}, {true, 1u << 20, 4, 12}, {true, 512u << 10, 4, 12}, {false, 1u << 20, 0, 8} };
    for (int rep = 0; rep < 2; rep++)
    for (const Cfg& c : cfgs) {
        const int R = c.R, W = F; const size_t PS = c.ps;
        std::vector<int> dfd(F);
        for (int f = 0; f < F; f++) { std::string p = dir + "/d" + std::to_string(f); unlink(p.c_str()); dfd[f] = open(p.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644); }
        std::atomic<size_t> next{0}; std::vector<std::atomic<int>> landed(NB); for (auto& x : landed) x = 0;
        const size_t rregion = RING / R, wregion = RING / W;
        double t0 = now();
        std::vector<std::thread> th;
        for (int r = 0; r < R; r++) th.emplace_back([&, r] {
            bind(true);
            cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            const int NPI = c.priv ? c.np : (int)(rregion / PS);          // pieces this reader cycles through
            std::vector<cudaEvent_t> pev(NPI); std::vector<char> used(NPI, 0);
            for (auto& e : pev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            std::deque<std::pair<size_t, cudaEvent_t>> inflight; std::vector<cudaEvent_t> evpool;
            char* base = ring + (size_t)r * rregion; size_t rot = 0;
            auto reap = [&](bool all) { while (!inflight.empty()) { auto& fr = inflight.front();
                    if (all) CK(cudaEventSynchronize(fr.second)); else if (cudaEventQuery(fr.second) != cudaSuccess) break;
                    landed[fr.first].store(1); evpool.push_back(fr.second); inflight.pop_front(); } };
            for (;;) { size_t k = next.fetch_add(1); if (k >= NB) break; int f = k % F; size_t o = (k / F) * BLK;
                for (size_t s = 0; s < BLK; s += PS) { int pi = (int)(rot++ % NPI); char* b = base + (size_t)pi * PS;
                    if (used[pi]) CK(cudaEventSynchronize(pev[pi]));
                    if (pread(sfd[f], b, PS, o + s) != (ssize_t)PS) exit(1);
                    CK(cudaMemcpyAsync(dev + f * G + o + s, b, PS, cudaMemcpyHostToDevice, st)); CK(cudaEventRecord(pev[pi], st)); used[pi] = 1; }
                cudaEvent_t e; if (evpool.empty()) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); else { e = evpool.back(); evpool.pop_back(); }
                CK(cudaEventRecord(e, st)); inflight.push_back({k, e}); reap(false); }
            reap(true);
        });
        for (int w = 0; w < W; w++) th.emplace_back([&, w] {
            bind(false);
            cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            const int NPI = c.priv ? 2 : (int)(wregion / PS);
            std::vector<cudaEvent_t> pev(NPI); for (auto& e : pev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            char* base = ring + RING + (size_t)w * wregion; size_t rot = 0;
            const size_t NP = G / PS;                              // pieces in this writer's file
            auto blk_of = [&](size_t p) { return (p * PS / BLK) * F + w; };
            auto issue = [&](size_t p) { size_t k = blk_of(p); while (!landed[k].load()) sched_yield();
                int pi = (int)(p % NPI); CK(cudaMemcpyAsync(base + (size_t)pi * PS, dev + w * G + p * PS, PS, cudaMemcpyDeviceToHost, st)); CK(cudaEventRecord(pev[pi], st)); };
            issue(0);
            for (size_t p = 0; p < NP; p++) { if (p + 1 < NP) issue(p + 1);
                int pi = (int)(p % NPI); CK(cudaEventSynchronize(pev[pi]));
                if (pwrite(dfd[w], base + (size_t)pi * PS, PS, p * PS) != (ssize_t)PS) exit(1); }
            (void)rot;
        });
        for (auto& t : th) t.join();
        double dt = now() - t0;
        printf("%-34s piece %4zu KiB x%d/reader, %2d readers + %d writers: %6.2f GiB/s\n", c.priv ? "private cache-sized pieces" : "pieces rotate through 512 MiB ring", PS >> 10, c.priv ? c.np : (int)(rregion / PS), R, W, F * (double)G / dt / (1 << 30));
        fflush(stdout);
        for (int f = 0; f < F; f++) close(dfd[f]);
    }
    std::string rm = "rm -rf " + dir; return system(rm.c_str());
}
