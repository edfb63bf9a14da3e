// vmig_tree.h -- source-tree manifest (walk) and destination metadata replay.
//
// What "the same result as the reference" means is fixed by what GNU tar restores when
// `(cd src; tar c .) | (cd dst; tar x)` runs as root (reference utils/copy.go:17-27): entry
// type, permission bits, uid/gid, mtime (whole seconds: tar's default gnu format), symlink
// targets, hard-link grouping inside the tree, char/block device numbers, FIFOs; sockets are
// skipped ("socket ignored"); xattrs/ACLs are dropped; existing destination entries are
// replaced, extras are kept.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include <sys/stat.h>
#include <sys/types.h>

namespace vmig {

enum EntryType : uint8_t { kDir = 1, kFile, kSymlink, kSpecial /* chr, blk, fifo */ };

struct Entry {
    std::string rel;          // relative path, no leading "./"; "." for the root itself
    EntryType   type;
    mode_t      mode;         // full st_mode
    uid_t       uid;
    gid_t       gid;
    struct timespec mtime;
    struct timespec atime;
    uint64_t    size = 0;     // regular files
    dev_t       rdev = 0;     // specials
    std::string target;       // symlinks
    int64_t     hardlink_of = -1;   // files: index (in Manifest::files) of the first path to the same inode
    uint64_t    first_block = 0;    // files: index of this file's first hash in the block table
    uint64_t    n_blocks = 0;
};

struct Manifest {
    std::vector<Entry> dirs;      // parents before children; dirs[0] is "."
    std::vector<Entry> files;     // sorted bytewise by rel (block-table order)
    std::vector<Entry> symlinks;
    std::vector<Entry> specials;
    uint64_t bytes_total = 0;     // sum of regular-file sizes (hard-linked paths counted once per path)
    uint64_t n_blocks = 0;
    uint64_t sockets_skipped = 0;
};

// Open `rel` (a manifest path) beneath root_fd with `flags`, refusing every symlink on the way and every
// escape from the root (openat2 RESOLVE_BENEATH|RESOLVE_NO_SYMLINKS; a component-wise O_NOFOLLOW walk on
// kernels without it).  A path that stopped being plain since the walk fails with VMIG_ESRCCHANGED.
int open_beneath(int root_fd, const std::string& rel, int flags, int* out_fd);
int open_beneath_walk(int root_fd, const std::string& rel, int flags, int* out_fd);   // the fallback, exposed for tests/tree_unit.cpp

// Walk src_root.  skip_hidden_topdirs reproduces `mv /root/src/*` (reference utils/copy.go:116).
int walk_tree(const std::string& src_root, uint32_t block_bytes, bool skip_hidden_topdirs, Manifest* out);

struct MetaPolicy {
    bool is_root;        // euid == 0: chown + exact modes (tar -p --same-owner defaults for root)
    mode_t umask_bits;   // applied when !is_root
    bool mtime_ns;       // VMIG_F_MTIME_NS
    bool no_metadata;    // VMIG_F_NO_METADATA
    bool keep_atime;     // move semantics (mv preserves atime too)
};
MetaPolicy default_meta_policy(uint32_t flags);

// Create every directory of the manifest under dst_root (parents first, traversable by us).
int make_dirs(const std::string& dst_root, const Manifest& m);
// Symlinks, specials, hard links and empty regular files; then directory metadata bottom-up.
int replay_metadata(const std::string& dst_root, const Manifest& m, const MetaPolicy& pol,
                    uint64_t* n_symlinks, uint64_t* n_hardlinks, uint64_t* n_specials);
// True if the open file already carries the owner/mode/mtime apply_file_meta would give it (then it is left alone: a
// diff pass that skips every block of a file must not move its ctime, the block table identifies files by it).
bool file_meta_matches(int fd, const Entry& e, const MetaPolicy& pol);
// Apply owner/mode/mtime to an open regular file (fd) or to a path (fd < 0).
int apply_file_meta(int fd, const std::string& path, const Entry& e, const MetaPolicy& pol);
// VMIG_F_MOVE_SRC: unlink migrated source entries, children before parents; the root stays.
int remove_source(const std::string& src_root, const Manifest& m);
// VMIG_F_PRUNE: remove every destination entry the manifest does not list (directories recursively); entries are
// handled by descriptor, symlinks are never followed.  dry_run only counts them (the verify pass uses that).
int prune_extras(const std::string& dst_root, const Manifest& m, bool dry_run, uint64_t* n_extras);
// Remove whatever non-directory sits at path (tar replaces existing entries).  ENOENT is fine.
int unlink_if_exists(const std::string& path, bool* was_dir);

}  // namespace vmig
