// tools/vmig_datagen.cpp -- deterministic synthetic trees of BASELINE.md §3 for bench.py (BOTH arms) and the
// profile scripts.  Stand-alone on purpose: links neither libvmig nor oracle/, so the reference arm of the bench
// can build its input without loading the product, and the product's arm without touching the checker.
//
//   vmig_datagen files DIR SEED N_FILES FILE_BYTES [THREADS]
//       N files f%05u.bin of FILE_BYTES each; bytes = SplitMix64 stream seeded SEED ^ fnv1a64(relative path),
//       little-endian 8-byte words (incompressible, no accidental duplicate blocks).  Configs 1, 2A, 3, 4, 5.
//   vmig_datagen layer DIR SEED TOTAL_BYTES N_FILES [THREADS]
//       config 2B, a "realistic" overlay2 diff layer: N_FILES regular files whose sizes are log-uniform over
//       1 KiB..64 MiB and then scaled so that they sum to TOTAL_BYTES (the two figures of SURVEY.md §8d cannot
//       both hold unscaled: the unscaled mean is 5.8 MiB), in a depth-4 tree a*/b*/c*/d*, 1 % symlinks, 0.5 %
//       empty files, one hard-link pair.
// Prints one line: files=<n> bytes=<sum>.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static uint64_t fnv1a64(const char* s, size_t n) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 0x100000001b3ULL; }
    return h;
}
static inline uint64_t splitmix_at(uint64_t seed, uint64_t j) {
    uint64_t z = seed + (j + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
struct Item { std::string rel; uint64_t size; uint64_t seed; };

static int mkdirs(const std::string& path) {
    for (size_t i = 1; i <= path.size(); i++)
        if (i == path.size() || path[i] == '/') {
            const std::string p = path.substr(0, i);
            if (mkdir(p.c_str(), 0755) != 0 && errno != EEXIST) { fprintf(stderr, "mkdir %s: %s\n", p.c_str(), strerror(errno)); return 1; }
        }
    return 0;
}

// chunk-major work order: consecutive work items hit different files (a tmpfs file takes writes one thread at a time)
static int write_items(const std::string& dir, const std::vector<Item>& items, unsigned threads) {
    const uint64_t kChunk = 4ull << 20;
    std::vector<int> fds(items.size(), -1);
    std::vector<uint64_t> first(items.size() + 1, 0);
    const bool keep_open = items.size() <= 512;
    for (size_t f = 0; f < items.size(); f++) {
        first[f + 1] = first[f] + (items[f].size + kChunk - 1) / kChunk;
        const std::string p = dir + "/" + items[f].rel;
        int fd = open(p.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
        if (fd < 0) { fprintf(stderr, "create %s: %s\n", p.c_str(), strerror(errno)); return 1; }
        if (keep_open) fds[f] = fd; else close(fd);
    }
    std::atomic<uint64_t> next{0}; std::atomic<int> bad{0};
    uint64_t max_chunks = 0;
    for (auto& it : items) max_chunks = std::max<uint64_t>(max_chunks, (it.size + kChunk - 1) / kChunk);
    std::vector<std::thread> th;
    if (keep_open) {
        const uint64_t total = (uint64_t)items.size() * max_chunks;
        for (unsigned t = 0; t < threads; t++)
            th.emplace_back([&] {
                std::vector<uint64_t> buf(kChunk / 8);
                for (;;) {
                    const uint64_t k = next.fetch_add(1);
                    if (k >= total || bad.load()) break;
                    const size_t f = (size_t)(k % items.size()); const uint64_t c = k / items.size();
                    const uint64_t off = c * kChunk;
                    if (off >= items[f].size) continue;
                    const uint64_t len = std::min<uint64_t>(kChunk, items[f].size - off), w0 = off / 8, nw = (len + 7) / 8;
                    for (uint64_t j = 0; j < nw; j++) buf[j] = splitmix_at(items[f].seed, w0 + j);
                    uint64_t put = 0;
                    while (put < len) {
                        ssize_t w = pwrite(fds[f], (const char*)buf.data() + put, len - put, (off_t)(off + put));
                        if (w <= 0) { bad.store(errno ? errno : EIO); break; }
                        put += (uint64_t)w;
                    }
                }
            });
    } else {
        for (unsigned t = 0; t < threads; t++)
            th.emplace_back([&] {
                std::vector<uint64_t> buf(kChunk / 8);
                for (;;) {
                    const uint64_t f = next.fetch_add(1);
                    if (f >= items.size() || bad.load()) break;
                    if (!items[f].size) continue;
                    const std::string p = dir + "/" + items[f].rel;
                    int fd = open(p.c_str(), O_WRONLY | O_CLOEXEC);
                    if (fd < 0) { bad.store(errno); break; }
                    for (uint64_t off = 0; off < items[f].size; off += kChunk) {
                        const uint64_t len = std::min<uint64_t>(kChunk, items[f].size - off), w0 = off / 8, nw = (len + 7) / 8;
                        for (uint64_t j = 0; j < nw; j++) buf[j] = splitmix_at(items[f].seed, w0 + j);
                        uint64_t put = 0;
                        while (put < len) {
                            ssize_t w = pwrite(fd, (const char*)buf.data() + put, len - put, (off_t)(off + put));
                            if (w <= 0) { bad.store(errno ? errno : EIO); break; }
                            put += (uint64_t)w;
                        }
                    }
                    close(fd);
                }
            });
    }
    for (auto& t : th) t.join();
    for (int x : fds) if (x >= 0) close(x);
    if (bad.load()) { fprintf(stderr, "write in %s: %s\n", dir.c_str(), strerror(bad.load())); return 1; }
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 6) { fprintf(stderr, "usage: %s files DIR SEED N BYTES [THREADS] | layer DIR SEED TOTAL N [THREADS]\n", argv[0]); return 2; }
    const std::string mode = argv[1], dir = argv[2];
    const uint64_t seed = strtoull(argv[3], nullptr, 0);
    const unsigned threads = argc > 6 ? (unsigned)atoi(argv[6]) : 16;
    if (mkdirs(dir)) return 1;
    std::vector<Item> items;
    uint64_t total = 0;
    if (mode == "files") {
        const uint64_t n = strtoull(argv[4], nullptr, 0), bytes = strtoull(argv[5], nullptr, 0);
        for (uint64_t f = 0; f < n; f++) {
            char name[64]; snprintf(name, sizeof name, "f%05u.bin", (unsigned)f);
            items.push_back({name, bytes, seed ^ fnv1a64(name, strlen(name))});
            total += bytes;
        }
        if (write_items(dir, items, threads)) return 1;
    } else if (mode == "layer") {
        const uint64_t want = strtoull(argv[4], nullptr, 0), n = strtoull(argv[5], nullptr, 0);
        std::vector<double> raw(n); double sum = 0;
        for (uint64_t i = 0; i < n; i++) {
            const double u = (double)(splitmix_at(seed ^ 0x5151ull, i) >> 11) / 9007199254740992.0;      // [0,1)
            raw[i] = std::exp(std::log(1024.0) + u * (std::log(64.0 * 1048576.0) - std::log(1024.0)));
            sum += raw[i];
        }
        for (uint64_t i = 0; i < n; i++) {
            uint64_t sz = (uint64_t)(raw[i] * (double)want / sum);
            if (i % 200 == 13) sz = 0;                                                                   // 0.5 % empty files
            char rel[128];
            snprintf(rel, sizeof rel, "a%u/b%u/c%u/d%u/f%06u.bin", (unsigned)(i % 8), (unsigned)((i / 8) % 8), (unsigned)((i / 64) % 8),
                     (unsigned)((i / 512) % 4), (unsigned)i);
            if (i < 2048) { std::string d = dir + "/" + rel; d.resize(d.rfind('/')); if (mkdirs(d)) return 1; }
            items.push_back({rel, sz, seed ^ fnv1a64(rel, strlen(rel))});
            total += sz;
        }
        if (write_items(dir, items, threads)) return 1;
        for (uint64_t i = 7; i < n; i += 100) {                                                           // 1 % symlinks
            std::string p = dir + "/" + items[i].rel; const std::string name = p.substr(p.rfind('/') + 1);
            p.resize(p.rfind('/')); p += "/l" + std::to_string(i);
            if (symlink(name.c_str(), p.c_str()) != 0 && errno != EEXIST) { fprintf(stderr, "symlink %s: %s\n", p.c_str(), strerror(errno)); return 1; }
        }
        const std::string a = dir + "/" + items[0].rel, b = dir + "/a1/hard";
        if (link(a.c_str(), b.c_str()) != 0 && errno != EEXIST) { fprintf(stderr, "link: %s\n", strerror(errno)); return 1; }
    } else {
        fprintf(stderr, "unknown mode %s\n", mode.c_str()); return 2;
    }
    printf("files=%zu bytes=%llu\n", items.size() + (mode == "layer" ? 1 : 0), (unsigned long long)total);
    return 0;
}
