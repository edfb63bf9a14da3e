// vmig_tree.cpp -- source-tree walk and destination metadata replay (see vmig_tree.h).
#include "vmig_tree.h"
#include "vmig_common.h"

#include <dirent.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/sysmacros.h>
#include <algorithm>
#include <map>
#include <utility>

namespace vmig {

static inline std::string join(const std::string& a, const std::string& b) {
    if (b == ".") return a;
    return a + "/" + b;
}

static int walk_dir(const std::string& root, const std::string& rel, uint32_t block_bytes, bool skip_hidden_topdirs,
                    Manifest* m, std::map<std::pair<dev_t, ino_t>, std::string>* inode_first, int depth)
{
    const std::string abs = join(root, rel);
    DIR* d = opendir(abs.c_str());
    if (!d) return fail(VMIG_EIO, "opendir %s: %s", abs.c_str(), errno_str(errno).c_str());
    std::vector<std::string> names;
    errno = 0;
    while (struct dirent* de = readdir(d)) {
        const char* n = de->d_name;
        if (n[0] == '.' && (n[1] == 0 || (n[1] == '.' && n[2] == 0))) continue;
        names.emplace_back(n);
    }
    const int rd_errno = errno;
    const int dfd_dup = dup(dirfd(d));
    closedir(d);
    if (rd_errno) { if (dfd_dup >= 0) close(dfd_dup); return fail(VMIG_EIO, "readdir %s: %s", abs.c_str(), errno_str(rd_errno).c_str()); }
    if (dfd_dup < 0) return fail(VMIG_EIO, "dup dirfd %s: %s", abs.c_str(), errno_str(errno).c_str());
    std::sort(names.begin(), names.end());

    std::vector<std::string> subdirs;
    for (const auto& n : names) {
        struct stat st;
        if (fstatat(dfd_dup, n.c_str(), &st, AT_SYMLINK_NOFOLLOW) != 0) {
            const int e = errno; close(dfd_dup);
            return fail(VMIG_EIO, "lstat %s/%s: %s", abs.c_str(), n.c_str(), errno_str(e).c_str());
        }
        Entry e;
        e.rel = rel == "." ? n : rel + "/" + n;
        e.mode = st.st_mode; e.uid = st.st_uid; e.gid = st.st_gid; e.mtime = st.st_mtim; e.atime = st.st_atim;
        switch (st.st_mode & S_IFMT) {
        case S_IFDIR:
            if (skip_hidden_topdirs && depth == 0 && n[0] == '.') continue;   // `mv /root/src/*` misses these
            e.type = kDir; m->dirs.push_back(e); subdirs.push_back(e.rel);
            break;
        case S_IFREG: {
            e.type = kFile; e.size = (uint64_t)st.st_size;
            e.n_blocks = (e.size + block_bytes - 1) / block_bytes;
            if (st.st_nlink > 1) {
                auto key = std::make_pair(st.st_dev, st.st_ino);
                auto it = inode_first->find(key);
                if (it == inode_first->end()) { inode_first->emplace(key, e.rel); e.target = e.rel; }
                else e.target = it->second;          // group key; resolved to an index after sorting
            }
            m->files.push_back(e);
            break;
        }
        case S_IFLNK: {
            e.type = kSymlink;
            std::string buf((size_t)st.st_size + 256, '\0');
            ssize_t k = readlinkat(dfd_dup, n.c_str(), &buf[0], buf.size());
            if (k < 0) { const int er = errno; close(dfd_dup); return fail(VMIG_EIO, "readlink %s: %s", e.rel.c_str(), errno_str(er).c_str()); }
            buf.resize((size_t)k); e.target = buf;
            m->symlinks.push_back(e);
            break;
        }
        case S_IFCHR: case S_IFBLK: case S_IFIFO:
            e.type = kSpecial; e.rdev = st.st_rdev; m->specials.push_back(e);
            break;
        case S_IFSOCK:
            m->sockets_skipped++;        // GNU tar: "socket ignored"
            break;
        default: break;
        }
    }
    close(dfd_dup);
    for (const auto& s : subdirs) {
        int rc = walk_dir(root, s, block_bytes, skip_hidden_topdirs, m, inode_first, depth + 1);
        if (rc) return rc;
    }
    return VMIG_OK;
}

int walk_tree(const std::string& src_root, uint32_t block_bytes, bool skip_hidden_topdirs, Manifest* m)
{
    *m = Manifest();
    struct stat st;
    if (stat(src_root.c_str(), &st) != 0) return fail(VMIG_EIO, "stat %s: %s", src_root.c_str(), errno_str(errno).c_str());
    if (!S_ISDIR(st.st_mode)) return fail(VMIG_ENOTDIR, "%s is not a directory", src_root.c_str());
    Entry root;
    root.rel = "."; root.type = kDir; root.mode = st.st_mode; root.uid = st.st_uid; root.gid = st.st_gid;
    root.mtime = st.st_mtim; root.atime = st.st_atim;
    m->dirs.push_back(root);
    std::map<std::pair<dev_t, ino_t>, std::string> inode_first;
    int rc = walk_dir(src_root, ".", block_bytes, skip_hidden_topdirs, m, &inode_first, 0);
    if (rc) return rc;

    std::sort(m->files.begin(), m->files.end(), [](const Entry& a, const Entry& b) { return a.rel < b.rel; });
    // resolve hard links to indices; the primary of a group is its bytewise-smallest path
    // (a group with a single member inside the tree is just a regular file)
    std::map<std::string, int64_t> group_min;
    for (size_t i = 0; i < m->files.size(); i++) {
        const Entry& e = m->files[i];
        if (e.target.empty()) continue;
        auto it = group_min.find(e.target);
        if (it == group_min.end()) group_min[e.target] = (int64_t)i;     // files are sorted: first seen = smallest
    }
    for (size_t i = 0; i < m->files.size(); i++) {
        Entry& e = m->files[i];
        if (e.target.empty()) continue;
        const int64_t p = group_min[e.target];
        if (p != (int64_t)i) e.hardlink_of = p;
        e.target.clear();
    }
    uint64_t fb = 0;
    for (auto& e : m->files) { e.first_block = fb; fb += e.n_blocks; m->bytes_total += e.size; }
    m->n_blocks = fb;
    return VMIG_OK;
}

MetaPolicy default_meta_policy(uint32_t flags)
{
    MetaPolicy p;
    p.is_root = geteuid() == 0;
    mode_t um = umask(0); umask(um);
    p.umask_bits = um;
    p.mtime_ns = (flags & VMIG_F_MTIME_NS) != 0 || (flags & VMIG_F_MOVE_SRC) != 0;
    p.no_metadata = (flags & VMIG_F_NO_METADATA) != 0;
    p.keep_atime = (flags & VMIG_F_MOVE_SRC) != 0;
    return p;
}

static inline mode_t eff_mode(const Entry& e, const MetaPolicy& pol) {
    mode_t m = e.mode & 07777;
    if (!pol.is_root) m &= ~pol.umask_bits & ~(mode_t)(S_ISUID | S_ISGID);   // tar without -p
    return m;
}
static inline void fill_times(const Entry& e, const MetaPolicy& pol, struct timespec ts[2]) {
    if (pol.keep_atime) ts[0] = e.atime; else { ts[0].tv_sec = 0; ts[0].tv_nsec = UTIME_OMIT; }
    ts[1] = e.mtime;
    if (!pol.mtime_ns) ts[1].tv_nsec = 0;   // tar's gnu format stores whole seconds
}

int unlink_if_exists(const std::string& path, bool* was_dir)
{
    struct stat st;
    if (was_dir) *was_dir = false;
    if (lstat(path.c_str(), &st) != 0) {
        if (errno == ENOENT) return VMIG_OK;
        return fail(VMIG_EIO, "lstat %s: %s", path.c_str(), errno_str(errno).c_str());
    }
    if (S_ISDIR(st.st_mode)) { if (was_dir) *was_dir = true; return VMIG_OK; }
    if (unlink(path.c_str()) != 0 && errno != ENOENT)
        return fail(VMIG_EIO, "unlink %s: %s", path.c_str(), errno_str(errno).c_str());
    return VMIG_OK;
}

int make_dirs(const std::string& dst_root, const Manifest& m)
{
    for (size_t i = 1; i < m.dirs.size(); i++) {
        const std::string p = join(dst_root, m.dirs[i].rel);
        if (mkdir(p.c_str(), 0700) == 0) continue;
        if (errno != EEXIST) return fail(VMIG_EIO, "mkdir %s: %s", p.c_str(), errno_str(errno).c_str());
        struct stat st;
        if (lstat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) continue;
        // a non-directory is in the way: tar replaces it
        if (unlink(p.c_str()) != 0 || mkdir(p.c_str(), 0700) != 0)
            return fail(VMIG_EIO, "replace %s by directory: %s", p.c_str(), errno_str(errno).c_str());
    }
    return VMIG_OK;
}

int apply_file_meta(int fd, const std::string& path, const Entry& e, const MetaPolicy& pol)
{
    if (pol.no_metadata) return VMIG_OK;
    struct timespec ts[2];
    fill_times(e, pol, ts);
    if (fd >= 0) {
        if (pol.is_root && fchown(fd, e.uid, e.gid) != 0) return fail(VMIG_EIO, "fchown %s: %s", path.c_str(), errno_str(errno).c_str());
        if (fchmod(fd, eff_mode(e, pol)) != 0) return fail(VMIG_EIO, "fchmod %s: %s", path.c_str(), errno_str(errno).c_str());
        if (futimens(fd, ts) != 0) return fail(VMIG_EIO, "futimens %s: %s", path.c_str(), errno_str(errno).c_str());
    } else {
        if (pol.is_root && lchown(path.c_str(), e.uid, e.gid) != 0) return fail(VMIG_EIO, "lchown %s: %s", path.c_str(), errno_str(errno).c_str());
        if (chmod(path.c_str(), eff_mode(e, pol)) != 0) return fail(VMIG_EIO, "chmod %s: %s", path.c_str(), errno_str(errno).c_str());
        if (utimensat(AT_FDCWD, path.c_str(), ts, 0) != 0) return fail(VMIG_EIO, "utimensat %s: %s", path.c_str(), errno_str(errno).c_str());
    }
    return VMIG_OK;
}

int replay_metadata(const std::string& dst_root, const Manifest& m, const MetaPolicy& pol,
                    uint64_t* n_symlinks, uint64_t* n_hardlinks, uint64_t* n_specials)
{
    struct timespec ts[2];
    for (const auto& e : m.symlinks) {
        const std::string p = join(dst_root, e.rel);
        bool was_dir = false;
        int rc = unlink_if_exists(p, &was_dir);
        if (rc) return rc;
        if (was_dir && rmdir(p.c_str()) != 0) return fail(VMIG_EIO, "rmdir %s (symlink over directory): %s", p.c_str(), errno_str(errno).c_str());
        if (symlink(e.target.c_str(), p.c_str()) != 0) return fail(VMIG_EIO, "symlink %s: %s", p.c_str(), errno_str(errno).c_str());
        if (!pol.no_metadata) {
            if (pol.is_root && lchown(p.c_str(), e.uid, e.gid) != 0) return fail(VMIG_EIO, "lchown %s: %s", p.c_str(), errno_str(errno).c_str());
            fill_times(e, pol, ts);
            if (utimensat(AT_FDCWD, p.c_str(), ts, AT_SYMLINK_NOFOLLOW) != 0) return fail(VMIG_EIO, "utimensat %s: %s", p.c_str(), errno_str(errno).c_str());
        }
        if (n_symlinks) (*n_symlinks)++;
    }
    for (const auto& e : m.specials) {
        const std::string p = join(dst_root, e.rel);
        bool was_dir = false;
        int rc = unlink_if_exists(p, &was_dir);
        if (rc) return rc;
        if (was_dir && rmdir(p.c_str()) != 0) return fail(VMIG_EIO, "rmdir %s: %s", p.c_str(), errno_str(errno).c_str());
        if (mknod(p.c_str(), (e.mode & S_IFMT) | 0600, e.rdev) != 0) return fail(VMIG_EIO, "mknod %s: %s", p.c_str(), errno_str(errno).c_str());
        rc = apply_file_meta(-1, p, e, pol);
        if (rc) return rc;
        if (n_specials) (*n_specials)++;
    }
    for (const auto& e : m.files) {
        if (e.hardlink_of < 0) continue;
        const std::string p = join(dst_root, e.rel);
        const std::string prim = join(dst_root, m.files[(size_t)e.hardlink_of].rel);
        bool was_dir = false;
        int rc = unlink_if_exists(p, &was_dir);
        if (rc) return rc;
        if (was_dir) return fail(VMIG_EIO, "%s: directory in the way of a hard link", p.c_str());
        if (link(prim.c_str(), p.c_str()) != 0) return fail(VMIG_EIO, "link %s -> %s: %s", p.c_str(), prim.c_str(), errno_str(errno).c_str());
        if (n_hardlinks) (*n_hardlinks)++;
    }
    if (!pol.no_metadata) {
        for (size_t i = m.dirs.size(); i-- > 0;) {     // children before parents
            const Entry& e = m.dirs[i];
            const std::string p = join(dst_root, e.rel);
            if (pol.is_root && chown(p.c_str(), e.uid, e.gid) != 0) return fail(VMIG_EIO, "chown %s: %s", p.c_str(), errno_str(errno).c_str());
            if (chmod(p.c_str(), eff_mode(e, pol)) != 0) return fail(VMIG_EIO, "chmod %s: %s", p.c_str(), errno_str(errno).c_str());
            fill_times(e, pol, ts);
            if (utimensat(AT_FDCWD, p.c_str(), ts, 0) != 0) return fail(VMIG_EIO, "utimensat %s: %s", p.c_str(), errno_str(errno).c_str());
        }
    }
    return VMIG_OK;
}

int remove_source(const std::string& src_root, const Manifest& m)
{
    auto rm = [&](const std::vector<Entry>& v) -> int {
        for (const auto& e : v) {
            const std::string p = join(src_root, e.rel);
            if (unlink(p.c_str()) != 0 && errno != ENOENT) return fail(VMIG_EIO, "unlink %s: %s", p.c_str(), errno_str(errno).c_str());
        }
        return VMIG_OK;
    };
    int rc;
    if ((rc = rm(m.files)) || (rc = rm(m.symlinks)) || (rc = rm(m.specials))) return rc;
    for (size_t i = m.dirs.size(); i-- > 1;) {
        const std::string p = join(src_root, m.dirs[i].rel);
        if (rmdir(p.c_str()) != 0 && errno != ENOENT && errno != ENOTEMPTY)
            return fail(VMIG_EIO, "rmdir %s: %s", p.c_str(), errno_str(errno).c_str());
    }
    return VMIG_OK;
}

}  // namespace vmig
