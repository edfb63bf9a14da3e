// vmig_util.cpp -- host-side helpers of the C ABI that need no GPU: block-table queries, the
// Go utils/file.go helpers the volume-resize caller uses next to the copy (reference
// internal/services/volume.go:117-140), and the deterministic synthetic-tree generator of
// BASELINE.md §3.
#include "vmig_common.h"
#include "vmig_table.h"
#include "vmig_tree.h"

#include <cctype>
#include <dirent.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include <atomic>
#include <thread>
#include <vector>
#include <cstdlib>
#include <cmath>

using namespace vmig;

static int dir_size_rec(const std::string& dir, int64_t* bytes, uint64_t* nfiles)
{
    DIR* d = opendir(dir.c_str());
    if (!d) return fail(VMIG_EIO, "opendir %s: %s", dir.c_str(), errno_str(errno).c_str());
    std::vector<std::string> subdirs;
    while (struct dirent* de = readdir(d)) {
        const char* n = de->d_name;
        if (n[0] == '.' && (n[1] == 0 || (n[1] == '.' && n[2] == 0))) continue;
        struct stat st;
        if (fstatat(dirfd(d), n, &st, AT_SYMLINK_NOFOLLOW) != 0) { int e = errno; closedir(d); return fail(VMIG_EIO, "lstat %s/%s: %s", dir.c_str(), n, errno_str(e).c_str()); }
        if (S_ISDIR(st.st_mode)) subdirs.push_back(dir + "/" + n);
        else { *bytes += st.st_size; (*nfiles)++; }       // filepath.Walk: every non-directory's Size()
    }
    closedir(d);
    for (auto& s : subdirs) { int rc = dir_size_rec(s, bytes, nfiles); if (rc) return rc; }
    return VMIG_OK;
}

static inline uint64_t splitmix_at(uint64_t seed, uint64_t j) {
    uint64_t z = seed + (j + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t fnv1a64(const char* s, size_t n) {
    uint64_t h = 0xCBF29CE484222325ULL;
    for (size_t i = 0; i < n; i++) { h ^= (uint8_t)s[i]; h *= 0x100000001B3ULL; }
    return h;
}

extern "C" {

int vmig_table_info_read(const char* path, vmig_table_info* out)
{
    if (!path || !out) return fail(VMIG_EINVAL, "null argument");
    BlockTable t; int rc = table_load(path, &t); if (rc) return rc;
    out->block_bytes = t.block_bytes; out->algo = t.algo; out->n_files = t.files.size(); out->n_blocks = t.hashes.size();
    out->bytes_total = t.bytes_total();
    return VMIG_OK;
}
int vmig_table_hashes(const char* path, uint64_t* out, uint64_t cap)
{
    if (!path || (!out && cap)) return fail(VMIG_EINVAL, "null argument");
    BlockTable t; int rc = table_load(path, &t); if (rc) return rc;
    const uint64_t n = std::min<uint64_t>(cap, t.hashes.size());
    if (n) memcpy(out, t.hashes.data(), n * 8);
    return VMIG_OK;
}

int vmig_manifest(const char* src_dir, uint32_t flags, uint32_t block_bytes, const char* out_table, vmig_stats* stats)
{
    if (!src_dir || !*src_dir) return fail(VMIG_EINVAL, "src_dir is NULL/empty");
    if (!block_bytes) block_bytes = 4u << 20;
    if (block_bytes & 4095u) return fail(VMIG_EINVAL, "block_bytes %u is not a multiple of 4096", block_bytes);
    Manifest man;
    const uint64_t t0 = now_ns();
    int rc = walk_tree(src_dir, block_bytes, (flags & VMIG_F_SKIP_HIDDEN_TOPDIRS) != 0, &man);
    if (rc) return rc;
    if (out_table && *out_table) {
        BlockTable t; t.block_bytes = block_bytes; t.algo = 1;
        for (auto& e : man.files) t.files.push_back({e.rel, e.size, e.first_block});
        t.hashes.assign(man.n_blocks, 0);
        rc = table_store(out_table, t);
        if (rc) return rc;
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->bytes_total = man.bytes_total; stats->blocks_total = man.n_blocks;
        stats->files = man.files.size(); stats->dirs = man.dirs.size(); stats->symlinks = man.symlinks.size();
        stats->specials = man.specials.size();
        for (auto& e : man.files) if (e.hardlink_of >= 0) stats->hardlinks++;
        stats->ns_walk = stats->ns_total = now_ns() - t0;
    }
    return VMIG_OK;
}

/* reference utils/file.go:13-22 (DirSize): sum of info.Size() over every non-directory entry */
int vmig_dir_size(const char* dir, int64_t* bytes, uint64_t* n_files)
{
    if (!dir || !bytes) return fail(VMIG_EINVAL, "null argument");
    struct stat st;
    if (lstat(dir, &st) != 0) return fail(VMIG_EIO, "lstat %s: %s", dir, errno_str(errno).c_str());
    int64_t b = 0; uint64_t n = 0;
    if (!S_ISDIR(st.st_mode)) { b = st.st_size; n = 1; }
    else { int rc = dir_size_rec(dir, &b, &n); if (rc) return rc; }
    *bytes = b; if (n_files) *n_files = n;
    return VMIG_OK;
}

/* reference utils/file.go:24-48 (ToBytes): last two chars are the unit, the rest ParseFloat */
int vmig_to_bytes(const char* s, int64_t* out)
{
    if (!s || !out) return fail(VMIG_EINVAL, "null argument");
    const size_t n = strlen(s);
    if (n < 3) return fail(VMIG_EINVAL, "cannot parse size '%s'", s);
    const std::string num(s, n - 2), unit(s + n - 2);
    // strconv.ParseFloat takes no leading white space (strtod would skip it)
    if (num.empty() || isspace((unsigned char)num[0])) return fail(VMIG_EINVAL, "cannot parse size '%s'", s);
    char* end = nullptr; errno = 0;
    const double v = strtod(num.c_str(), &end);
    if (end == num.c_str() || *end != 0 || errno == ERANGE) return fail(VMIG_EINVAL, "cannot parse size '%s'", s);
    int64_t mult;
    if (unit == "KB") mult = 1ll << 10; else if (unit == "MB") mult = 1ll << 20;
    else if (unit == "GB") mult = 1ll << 30; else if (unit == "TB") mult = 1ll << 40;
    else return fail(VMIG_EINVAL, "unsupported unit: %s", unit.c_str());
    *out = (int64_t)(v * (double)mult);
    return VMIG_OK;
}

int vmig_datagen_files(const char* dir, uint64_t seed, uint32_t n_files, uint64_t file_bytes, uint32_t threads)
{
    if (!dir || !n_files) return fail(VMIG_EINVAL, "bad datagen arguments");
    if (mkdir(dir, 0755) != 0 && errno != EEXIST) return fail(VMIG_EIO, "mkdir %s: %s", dir, errno_str(errno).c_str());
    if (!threads) threads = 16;
    const uint64_t kChunk = 4ull << 20;
    const uint64_t chunks_per_file = (file_bytes + kChunk - 1) / kChunk;
    std::vector<int> fds(n_files, -1);
    std::vector<uint64_t> seeds(n_files);
    for (uint32_t f = 0; f < n_files; f++) {
        char name[64]; snprintf(name, sizeof name, "f%05u.bin", f);
        seeds[f] = seed ^ fnv1a64(name, strlen(name));
        const std::string p = std::string(dir) + "/" + name;
        fds[f] = open(p.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
        if (fds[f] < 0) { int e = errno; for (int x : fds) if (x >= 0) close(x); return fail(VMIG_EIO, "create %s: %s", p.c_str(), errno_str(e).c_str()); }
    }
    std::atomic<uint64_t> next{0}; std::atomic<int> bad{0};
    const uint64_t total = (uint64_t)n_files * chunks_per_file;
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < threads; t++)
        th.emplace_back([&] {
            std::vector<uint64_t> buf(kChunk / 8);
            for (;;) {
                const uint64_t k = next.fetch_add(1);
                if (k >= total || bad.load()) break;
                // chunk-major order: consecutive work items hit different files
                const uint32_t f = (uint32_t)(k % n_files); const uint64_t c = k / n_files;
                const uint64_t off = c * kChunk, len = std::min<uint64_t>(kChunk, file_bytes - off);
                const uint64_t w0 = off / 8, nw = (len + 7) / 8;
                for (uint64_t j = 0; j < nw; j++) buf[j] = splitmix_at(seeds[f], w0 + j);
                uint64_t put = 0;
                while (put < len) {
                    ssize_t w = pwrite(fds[f], (const char*)buf.data() + put, len - put, (off_t)(off + put));
                    if (w <= 0) { bad.store(errno ? errno : EIO); break; }
                    put += (uint64_t)w;
                }
            }
        });
    for (auto& t : th) t.join();
    for (int x : fds) close(x);
    if (bad.load()) return fail(VMIG_EIO, "datagen write in %s: %s", dir, errno_str(bad.load()).c_str());
    return VMIG_OK;
}

}  // extern "C"
