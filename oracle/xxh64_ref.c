/*
 * oracle/xxh64_ref.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of canonical XXH64 (seed-parameterised, the engine
 * uses seed 0) and of the per-file 4 MiB block-table the migration engine
 * produces.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load this library; the product (libvmig.so) never links it.
 *
 * Provenance / pinning
 * --------------------
 * The reference (XShengTech/gpu-docker-api) performs no hashing at all
 * (no xxhash/sha/crc import in go.mod:5-19, utils/copy.go:1-14),
 * so block hashes are "parity unpinned" BY THE REFERENCE.  They are pinned
 * instead to the public xxHash specification (XXH64, xxHash 0.8.2):
 *   - the known-answer vectors of SURVEY.md Appendix A (tests/golden/xxh64_kat.json)
 *   - libxxhash.so.0.8.2 / python-xxhash 3.7.0 when present on the box
 * (tests/test_oracle.py checks all three agree).
 *
 * Algorithm (public spec, little-endian reads, arithmetic mod 2^64):
 *   round(a,x) = rotl(a + x*P2, 31) * P1
 *   merge(h,v) = (h ^ round(0,v)) * P1 + P4
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>

#define P1 0x9E3779B185EBCA87ULL
#define P2 0xC2B2AE3D27D4EB4FULL
#define P3 0x165667B19E3779F9ULL
#define P4 0x85EBCA77C2B2AE63ULL
#define P5 0x27D4EB2F165667C5ULL

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; } /* x86: LE */
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t xround(uint64_t acc, uint64_t x) { return rotl64(acc + x * P2, 31) * P1; }
static inline uint64_t xmerge(uint64_t h, uint64_t v) { return (h ^ xround(0, v)) * P1 + P4; }

uint64_t oracle_xxh64(const void *data, uint64_t len, uint64_t seed)
{
    const uint8_t *p = (const uint8_t *)data;
    const uint8_t *end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        const uint8_t *limit = end - 32;
        do {
            v1 = xround(v1, rd64(p));
            v2 = xround(v2, rd64(p + 8));
            v3 = xround(v3, rd64(p + 16));
            v4 = xround(v4, rd64(p + 24));
            p += 32;
        } while (p <= limit);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
    } else {
        h = seed + P5;
    }
    h += len;
    while (p + 8 <= end) { h = rotl64(h ^ xround(0, rd64(p)), 27) * P1 + P4; p += 8; }
    if (p + 4 <= end)    { h = rotl64(h ^ ((uint64_t)rd32(p) * P1), 23) * P2 + P3; p += 4; }
    while (p < end)      { h = rotl64(h ^ ((uint64_t)(*p) * P5), 11) * P1; p++; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

/* Hash n blocks described by (offs[i], lens[i]) inside one buffer. */
void oracle_hash_blocks(const void *buf, const uint64_t *offs, const uint32_t *lens,
                        uint64_t n, uint64_t *out)
{
    for (uint64_t i = 0; i < n; i++)
        out[i] = oracle_xxh64((const uint8_t *)buf + offs[i], lens[i], 0);
}

/* Block hashes of one regular file, block_bytes-strided from offset 0 (the
 * engine's block-table rule: SURVEY.md §8 a8; an empty file has 0 blocks).
 * Returns number of blocks written, or -1 on I/O error. */
int64_t oracle_hash_file(const char *path, uint32_t block_bytes, uint64_t *out, uint64_t max_out)
{
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    uint8_t *buf = (uint8_t *)malloc(block_bytes);
    if (!buf) { close(fd); return -1; }
    int64_t nb = 0;
    for (;;) {
        size_t got = 0;
        while (got < block_bytes) {
            ssize_t r = read(fd, buf + got, block_bytes - got);
            if (r < 0) { free(buf); close(fd); return -1; }
            if (r == 0) break;
            got += (size_t)r;
        }
        if (got == 0) break;
        if ((uint64_t)nb < max_out) out[nb] = oracle_xxh64(buf, got, 0);
        nb++;
        if (got < block_bytes) break;
    }
    free(buf); close(fd);
    return nb;
}

/* SplitMix64 stream: the deterministic synthetic-data generator of BASELINE.md §3.
 * word j of the stream (0-based) with state seed: z = seed + (j+1)*GAMMA; mix. */
static inline uint64_t splitmix_at(uint64_t seed, uint64_t j)
{
    uint64_t z = seed + (j + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
void oracle_splitmix_fill(uint64_t seed, uint64_t first_word, uint64_t nbytes, void *dst)
{
    uint8_t *d = (uint8_t *)dst;
    uint64_t nw = nbytes / 8, j;
    for (j = 0; j < nw; j++) { uint64_t v = splitmix_at(seed, first_word + j); memcpy(d + 8 * j, &v, 8); }
    if (nbytes & 7) { uint64_t v = splitmix_at(seed, first_word + nw); memcpy(d + 8 * nw, &v, nbytes & 7); }
}
uint64_t oracle_fnv1a64(const void *s, uint64_t n)
{
    const uint8_t *p = (const uint8_t *)s; uint64_t h = 0xCBF29CE484222325ULL;
    for (uint64_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001B3ULL; }
    return h;
}
