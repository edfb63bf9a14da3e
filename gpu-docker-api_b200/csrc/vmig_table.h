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

struct TableFile { std::string rel; uint64_t size; uint64_t first_block; };

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
