// vmig_table.h -- per-(resource, version) block table: file manifest + one XXH64 per 4 MiB block.
// File format: include/vmig.h ("Block-table file").  No reference symbol exists for it (the
// reference has no hashing); its home is the per-version directory that
// internal/services/replicaset.go:681-704 (setToMergeMap) creates and
// internal/version/merge.go:16 (ContainerMergeMap) records.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include <unordered_map>

namespace vmig {

// ino / ctime_ns identify the on-disk file the hashes were taken FOR: the destination file the migration that
// wrote the table left behind (the source file for a VMIG_F_HASH_ONLY table).  0/0 = unknown (format 01, hard-linked
// paths).  The in-place diff path trusts a prior table for a destination file only while both still match
// (ctime cannot be set from user space: any out-of-band write, truncate, chmod or rename-over bumps it).
struct TableFile { std::string rel; uint64_t size; uint64_t first_block; uint64_t ino = 0; int64_t ctime_ns = 0; };

struct BlockTable {
    uint32_t block_bytes = 0;
    uint32_t algo = 1;                       // 1 = XXH64, seed 0
    std::vector<TableFile> files;            // sorted bytewise by rel
    std::vector<uint64_t> hashes;
    std::unordered_map<std::string, size_t> index;   // rel -> files[] index (built by load)
    uint64_t bytes_total() const { uint64_t s = 0; for (auto& f : files) s += f.size; return s; }
    uint64_t blocks_of(const TableFile& f) const { return (f.size + block_bytes - 1) / block_bytes; }
};

int table_load(const std::string& path, BlockTable* out);
// Atomic: writes path.tmp.<pid>.<tid>, fsyncs, renames over path.
int table_store(const std::string& path, const BlockTable& t);

}  // namespace vmig
