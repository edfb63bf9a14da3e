// vmig_kernels.cuh -- launch interface of the sm_100a kernels of libvmig (see vmig_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vmig {

// Every block handed to xxh64_blocks must start at a 16-byte aligned device address and be
// followed by >= 32 readable bytes (the engine lays blocks out at 512-byte aligned slot
// offsets inside padded allocations, so both hold by construction).
constexpr uint32_t kBlockAlign = 512;
constexpr uint32_t kTailPad    = 64;

struct HashLaunch {
    const uint8_t*  base;        // device base pointer of the staged bytes
    const uint64_t* offs;        // [n] device: byte offset of block i from base (16-B aligned)
    const uint32_t* lens;        // [n] device: length of block i in bytes (may be 0)
    uint32_t        n;
    uint64_t*       hashes;      // [n] device out: XXH64(seed 0)
    const uint64_t* prior;       // [n] device, nullable: prior version's hash of the same block
    const uint8_t*  prior_valid; // [n] device, nullable: 1 where prior[i] is meaningful
    uint8_t*        changed;     // [n] device out, nullable: 1 = must be copied back / written
    uint32_t*       work_counter;// device u32; zeroed by the launcher on the same stream
};

// K1: canonical XXH64 per block (+ fused compare against the prior table).
cudaError_t launch_xxh64_blocks(const HashLaunch& a, int sm_count, cudaStream_t st);
// K2: ordered compaction of changed[] into survivors[] (ascending block index) and *n_survivors.
cudaError_t launch_diff_select(const uint8_t* changed, uint32_t n, uint32_t* survivors,
                               uint32_t* n_survivors, cudaStream_t st);
// Resident-batch helpers (bench / tests): SplitMix64 fill and first-8-bytes flip.
cudaError_t launch_splitmix_fill(uint8_t* base, uint64_t nbytes, uint64_t seed, cudaStream_t st);
cudaError_t launch_flip_first8(uint8_t* base, const uint64_t* offs, const uint64_t* blocks, uint64_t n,
                               cudaStream_t st);
// Dynamic shared memory the hash kernel needs (for diagnostics).
size_t xxh64_blocks_smem_bytes();

}  // namespace vmig
