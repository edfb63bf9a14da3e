// vmig_kernels.cu -- hand-written sm_100a kernels of the volume-migration engine.
//
// K1  xxh64_blocks : canonical XXH64 (seed 0) of every staged file block, fused with the compare
//                    against the prior version's block table (north star: "per-block xxHash64 ...
//                    compare against the prior version's block table").  The reference has no
//                    hashing; the algorithm is the public xxHash spec restated in
//                    SURVEY.md Appendix A and checked bit-for-bit against oracle/xxh64_ref.c.
// K2  diff_select  : ordered compaction of the changed flags into a survivor index list.
//
// Why K1 looks the way it does:
//  * XXH64 is four dependent chains per block (acc = rotl(acc + x*P2, 31) * P1 per 8 input
//    bytes per chain; rotl breaks associativity) so the only parallelism inside a block is 4.
//    Parallelism comes from hashing many blocks at once: one QUAD (4 lanes) per block, 8 blocks
//    per warp, kWarps warps per CTA, one persistent CTA per SM.
//  * Each lane walks an 8-byte column with a 32-byte stride -- poison for global loads -- so the
//    bytes are staged through shared memory by the TMA engine: one elected lane per quad issues
//    1-D bulk copies (cp.async.bulk global->shared, SASS UBLKCP) of kChunk contiguous bytes of
//    its block into a kStages-deep ring and the quad consumes them with conflict-free LDS.64
//    (quad slots are padded by 32 B so the 4 quads of a half-warp hit disjoint bank groups).
//    Completion is tracked by one mbarrier per (warp, stage); no CTA-wide barrier is used in
//    the steady state and no LSU bandwidth is spent on the copy.
//  * No tensor cores: there is no contraction here, only 64-bit integer mul/add/rotate.
//  * Work distribution: quad (cta c, warp w, quad q) starts on block c + G*(w + kWarps*q) so a
//    small batch spreads over all SMs first, then over the 4 SM sub-partitions; afterwards
//    quads pull block indices from a global atomic counter (ragged block lengths balance).
#include "vmig_kernels.cuh"

namespace vmig {

namespace {

constexpr uint64_t P1 = 0x9E3779B185EBCA87ULL;
constexpr uint64_t P2 = 0xC2B2AE3D27D4EB4FULL;
constexpr uint64_t P3 = 0x165667B19E3779F9ULL;
constexpr uint64_t P4 = 0x85EBCA77C2B2AE63ULL;
constexpr uint64_t P5 = 0x27D4EB2F165667C5ULL;

constexpr int kWarps  = 4;      // one per SM sub-partition
constexpr int kQuads  = 8;      // blocks in flight per warp
constexpr int kChunk  = 2048;   // bytes per bulk copy (64 stripes)
constexpr int kStages = 3;      // ring depth per quad
constexpr int kQuadStride  = kChunk + 32;              // +32 B: bank-group skew between quads
constexpr int kStageStride = kQuads * kQuadStride;
constexpr int kWarpData    = kStages * kStageStride;
constexpr int kWarpDesc    = kStages * kQuads * 16;     // uint4 chunk descriptors
constexpr int kWarpBars    = kStages * 8;               // mbarriers
constexpr int kWarpSmem    = kWarpData + kWarpDesc + ((kWarpBars + 15) & ~15);
constexpr int kSmemBytes   = kWarps * kWarpSmem;
static_assert(kChunk % 256 == 0, "chunk must be a multiple of 8 stripes");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr uint32_t kFlagFirst = 1u, kFlagLast = 2u, kFlagNone = 4u;

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t xround(uint64_t acc, uint64_t x) { return rotl64(acc + x * P2, 31) * P1; }
__device__ __forceinline__ uint64_t xmerge(uint64_t h, uint64_t v) { return (h ^ xround(0, v)) * P1 + P4; }


// The XXH64 round acc' = rotl(acc + x*P2, 31) * P1 arranged as a 4-level dependent chain
// (nvcc's own lowering of the 64-bit expression is 7 levels deep: it emulates the rotate with
// SHF+SHF+LOP3 and completes the 64-bit product before the next add).
// Carried state: acc = w + (v << 32) with w = r.lo*P1.lo (64-bit) and v = r.lo*P1.hi + r.hi*P1.lo
// (the cross terms, only needed in the high word).  One round with input word x:
//   t    = x.lo * P2.lo + w                 IMAD.WIDE.U32 (64-bit addend = previous wide)  level 1
//   t.hi = t.hi + v + (x.lo*P2.hi + x.hi*P2.lo)   IADD3; the x terms are off the chain     level 2
//   r    = rotl(t, 31)                      2 x SHF.L.W (funnel shifts)                     level 3
//   w'   = r.lo * P1.lo                     IMAD.WIDE.U32                                   level 4
//   v'   = r.lo*P1.hi + r.hi*P1.lo          2 x IMAD, consumed only at level 2 of the next round
struct Chain {
    uint64_t w; uint32_t v;
    static constexpr uint32_t P1lo = (uint32_t)P1, P1hi = (uint32_t)(P1 >> 32);
    static constexpr uint32_t P2lo = (uint32_t)P2, P2hi = (uint32_t)(P2 >> 32);
    __device__ __forceinline__ void begin(uint64_t acc) { w = acc; v = 0; }
    // Written in PTX so that neither NVVM nor ptxas re-associates the sums into a serial IMAD
    // chain (both minimise instruction count, which here lengthens the dependent chain).
    __device__ __forceinline__ void step(uint64_t x) {
        uint32_t xl, xh, mh, tlo, thi, rl, rh;
        asm("mov.b64 {%0, %1}, %2;" : "=r"(xl), "=r"(xh) : "l"(x));
        asm("{\n\t.reg .u32 a;\n\tmul.lo.u32 a, %1, %3;\n\tmad.lo.u32 %0, %2, %4, a;\n\t}"
            : "=r"(mh) : "r"(xl), "r"(xh), "r"(P2hi), "r"(P2lo));                 // off-chain
        asm("{\n\t.reg .u64 t;\n\tmad.wide.u32 t, %2, %3, %4;\n\tmov.b64 {%0, %1}, t;\n\t}"
            : "=r"(tlo), "=r"(thi) : "r"(xl), "r"(P2lo), "l"(w));                 // level 1
        thi = thi + v + mh;                                                       // level 2 (IADD3)
        rl = __funnelshift_l(thi, tlo, 31);                                       // level 3
        rh = __funnelshift_l(tlo, thi, 31);
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(rl), "r"(P1lo));           // level 4
        asm("{\n\t.reg .u32 a;\n\tmul.lo.u32 a, %2, %3;\n\tmad.lo.u32 %0, %1, %4, a;\n\t}"
            : "=r"(v) : "r"(rl), "r"(rh), "r"(P1lo), "r"(P1hi));
    }
    __device__ __forceinline__ uint64_t end() const { return w + ((uint64_t)v << 32); }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP.S.G).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
    uint64_t v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint64_t finish_hash(uint64_t v1, uint64_t v2, uint64_t v3, uint64_t v4, uint32_t len,
                                                const uint8_t* tail) {
    uint64_t h;
    if (len >= 32) {
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
    } else {
        h = P5;  // seed 0
    }
    h += (uint64_t)len;
    // tail = the (len & 31) bytes after the last full stripe; its address is 16-B aligned and
    // the allocation is padded, so two 16-byte loads are always in bounds.
    uint32_t rem = len & 31u;
    if (rem) {
        const uint4 t0 = __ldg(reinterpret_cast<const uint4*>(tail));
        const uint4 t1 = __ldg(reinterpret_cast<const uint4*>(tail) + 1);
        uint32_t w[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
        uint32_t i = 0;  // index in 32-bit words
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (rem >= 8) {
                uint64_t x = (uint64_t)w[2 * k] | ((uint64_t)w[2 * k + 1] << 32);
                h = rotl64(h ^ xround(0, x), 27) * P1 + P4;
                rem -= 8; i += 2;
            }
        }
        uint32_t cur = 0, nxt = 0;  // select w[i], w[i+1] without dynamic register indexing
#pragma unroll
        for (int k = 0; k < 8; k++) { if ((uint32_t)k == i) cur = w[k]; if ((uint32_t)k == i + 1) nxt = w[k]; }
        if (rem >= 4) {
            h = rotl64(h ^ ((uint64_t)cur * P1), 23) * P2 + P3;
            rem -= 4; cur = nxt;
        }
        for (; rem; rem--) {
            h = rotl64(h ^ ((uint64_t)(cur & 0xFFu) * P5), 11) * P1;
            cur >>= 8;
        }
    }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

__global__ void __launch_bounds__(kWarps * 32, 1)
xxh64_blocks_kernel(const uint8_t* __restrict__ base, const uint64_t* __restrict__ offs,
                    const uint32_t* __restrict__ lens, uint32_t n, uint64_t* __restrict__ hashes,
                    const uint64_t* __restrict__ prior, const uint8_t* __restrict__ prior_valid,
                    uint8_t* __restrict__ changed, uint32_t* __restrict__ work_counter)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t quad = lane >> 2, k = lane & 3u;
    const bool leader = (k == 0);
    const uint32_t quad_mask = 0xFu << (lane & ~3u);

    uint8_t* wbase = smem + warp * kWarpSmem;
    const uint32_t data_s = smem_u32(wbase);
    uint4* descs = reinterpret_cast<uint4*>(wbase + kWarpData);
    const uint32_t bars_s = smem_u32(wbase + kWarpData + kWarpDesc);

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) mbar_init(bars_s + 8 * s, kQuads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // ---- issue-side state (quad-uniform; every lane tracks it, the leader acts on it)
    const uint32_t total_quads = gridDim.x * kWarps * kQuads;
    uint32_t iss_blk = blockIdx.x + gridDim.x * (warp + kWarps * quad);  // first assignment
    bool     iss_need_fetch = false, iss_done = false, iss_first = true;
    uint64_t iss_off = 0;
    uint32_t iss_rem = 0, iss_len = 0;
    if (iss_blk < n) {
        iss_off = offs[iss_blk]; iss_len = lens[iss_blk]; iss_rem = iss_len & ~31u;
    } else {
        iss_done = true;
    }

    auto issue = [&](int stage) {
        if (iss_need_fetch && !iss_done) {
            uint32_t idx = 0;
            if (leader) idx = total_quads + atomicAdd(work_counter, 1u);
            idx = __shfl_sync(quad_mask, idx, lane & ~3u);
            iss_need_fetch = false;
            if (idx < n) {
                iss_blk = idx; iss_off = offs[idx]; iss_len = lens[idx]; iss_rem = iss_len & ~31u; iss_first = true;
            } else {
                iss_done = true;
            }
        }
        const uint32_t bar = bars_s + 8 * stage;
        if (iss_done) {
            if (leader) { descs[stage * kQuads + quad] = make_uint4(0, 0, kFlagNone, 0); mbar_arrive(bar); }
            return;
        }
        const uint32_t nb = iss_rem < (uint32_t)kChunk ? iss_rem : (uint32_t)kChunk;
        const uint32_t flags = (iss_first ? kFlagFirst : 0u) | (iss_rem == nb ? kFlagLast : 0u);
        if (leader) {
            descs[stage * kQuads + quad] = make_uint4(nb, iss_blk, flags, iss_len);
            if (nb) {
                mbar_arrive_expect_tx(bar, nb);
                bulk_g2s(data_s + stage * kStageStride + quad * kQuadStride, base + iss_off, nb, bar);
            } else {
                mbar_arrive(bar);
            }
        }
        iss_off += nb; iss_rem -= nb; iss_first = false;
        if (flags & kFlagLast) iss_need_fetch = true;
    };

#pragma unroll
    for (int s = 0; s < kStages - 1; s++) issue(s);

    uint64_t acc = 0;
    for (uint32_t it = 0;; it++) {
        const int stage = it % kStages;
        issue((it + kStages - 1) % kStages);   // that stage was drained in iteration it-1
        mbar_wait(bars_s + 8 * stage, (it / kStages) & 1u);
        const uint4 d = descs[stage * kQuads + quad];
        if (__all_sync(0xFFFFFFFFu, (d.z & kFlagNone) != 0)) break;   // nothing left anywhere in this warp

        if (!(d.z & kFlagNone)) {
            if (d.z & kFlagFirst) {
                // seed 0: v1 = P1+P2, v2 = P2, v3 = 0, v4 = -P1
                acc = (k == 0) ? (P1 + P2) : (k == 1) ? P2 : (k == 2) ? 0ULL : (0ULL - P1);
            }
            uint32_t sp = data_s + stage * kStageStride + quad * kQuadStride + k * 8;
            uint32_t rounds = d.x >> 5;
            // 8 rounds per group; the next group's LDS.64 are issued before the current chain.
            uint64_t x[8];
            if (rounds >= 8) {
#pragma unroll
                for (int j = 0; j < 8; j++) x[j] = lds64(sp + 32 * j);
                sp += 256;
                Chain c;
                c.begin(acc);
                while (rounds >= 16) {
                    uint64_t y[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) y[j] = lds64(sp + 32 * j);
                    sp += 256;
#pragma unroll
                    for (int j = 0; j < 8; j++) c.step(x[j]);
#pragma unroll
                    for (int j = 0; j < 8; j++) x[j] = y[j];
                    rounds -= 8;
                }
#pragma unroll
                for (int j = 0; j < 8; j++) c.step(x[j]);
                acc = c.end();
                rounds -= 8;
            }
            for (; rounds; rounds--) { acc = xround(acc, lds64(sp)); sp += 32; }

            if (d.z & kFlagLast) {
                const uint32_t qb = lane & ~3u;
                const uint64_t v1 = __shfl_sync(quad_mask, acc, qb + 0);
                const uint64_t v2 = __shfl_sync(quad_mask, acc, qb + 1);
                const uint64_t v3 = __shfl_sync(quad_mask, acc, qb + 2);
                const uint64_t v4 = __shfl_sync(quad_mask, acc, qb + 3);
                if (leader) {
                    const uint32_t blk = d.y, len = d.w;
                    const uint64_t h = finish_hash(v1, v2, v3, v4, len, base + offs[blk] + (len & ~31u));
                    hashes[blk] = h;
                    if (changed) {
                        const bool same = prior && prior_valid && prior_valid[blk] && prior[blk] == h;
                        changed[blk] = same ? 0 : 1;
                    }
                }
            }
        }
        __syncwarp();   // all lanes are done reading this stage before it is refilled
    }
}

// ---- K2: ordered compaction of changed[] (single CTA; n is O(10^4..10^6), 1 byte per block)
__global__ void __launch_bounds__(1024, 1)
diff_select_kernel(const uint8_t* __restrict__ changed, uint32_t n, uint32_t* __restrict__ survivors,
                   uint32_t* __restrict__ n_survivors)
{
    __shared__ uint32_t warp_counts[32];
    __shared__ uint32_t running;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (uint32_t start = 0; start < n; start += 1024) {
        const uint32_t i = start + threadIdx.x;
        const bool keep = i < n && changed[i] != 0;
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) warp_counts[warp] = __popc(ballot);
        __syncthreads();
        uint32_t woff = 0, total = 0;
        {
            uint32_t c = warp_counts[lane];
            uint32_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += t; }
            woff  = __shfl_sync(0xFFFFFFFFu, incl - c, warp);
            total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        const uint32_t base = running;
        if (keep) survivors[base + woff + __popc(ballot & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) running = base + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_survivors = running;
}

__device__ __forceinline__ uint64_t splitmix_at(uint64_t seed, uint64_t j) {
    uint64_t z = seed + (j + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__global__ void splitmix_fill_kernel(uint64_t* __restrict__ out, uint64_t nwords, uint64_t seed) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 2;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < nwords; i += stride) {
        ulonglong2 v;
        v.x = splitmix_at(seed, i);
        v.y = splitmix_at(seed, i + 1);
        if (i + 1 < nwords) *reinterpret_cast<ulonglong2*>(out + i) = v; else out[i] = v.x;
    }
}
__global__ void flip_first8_kernel(uint8_t* base, const uint64_t* offs, const uint64_t* blocks, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { uint64_t* p = reinterpret_cast<uint64_t*>(base + offs[blocks[i]]); *p ^= ~0ULL; }
}

}  // namespace

size_t xxh64_blocks_smem_bytes() { return kSmemBytes; }

cudaError_t launch_xxh64_blocks(const HashLaunch& a, int sm_count, cudaStream_t st)
{
    if (a.n == 0) return cudaSuccess;
    static thread_local bool attr_set[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && !attr_set[dev]) {
        e = cudaFuncSetAttribute(xxh64_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    e = cudaMemsetAsync(a.work_counter, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    // persistent grid: one CTA per SM, never more CTAs than blocks
    uint32_t grid = (uint32_t)sm_count;
    if (a.n < grid) grid = a.n;
    xxh64_blocks_kernel<<<grid, kWarps * 32, kSmemBytes, st>>>(a.base, a.offs, a.lens, a.n, a.hashes, a.prior,
                                                              a.prior_valid, a.changed, a.work_counter);
    return cudaGetLastError();
}

cudaError_t launch_diff_select(const uint8_t* changed, uint32_t n, uint32_t* survivors, uint32_t* n_survivors,
                               cudaStream_t st)
{
    diff_select_kernel<<<1, 1024, 0, st>>>(changed, n, survivors, n_survivors);
    return cudaGetLastError();
}

cudaError_t launch_splitmix_fill(uint8_t* base, uint64_t nbytes, uint64_t seed, cudaStream_t st)
{
    const uint64_t nwords = nbytes / 8;
    if (!nwords) return cudaSuccess;
    splitmix_fill_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<uint64_t*>(base), nwords, seed);
    return cudaGetLastError();
}

cudaError_t launch_flip_first8(uint8_t* base, const uint64_t* offs, const uint64_t* blocks, uint64_t n, cudaStream_t st)
{
    if (!n) return cudaSuccess;
    flip_first8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(base, offs, blocks, n);
    return cudaGetLastError();
}

}  // namespace vmig
