// vmig_kernels.cu -- hand-written sm_100a kernels of the volume-migration engine.
//
// K1  xxh64_blocks : canonical XXH64 (seed 0) of every staged file block, fused with the compare
//                    against the prior version's block table (north star: "per-block xxHash64 ...
//                    compare against the prior version's block table").  The reference has no
//                    hashing; the algorithm is the public xxHash spec restated in
//                    SURVEY.md Appendix A and checked bit-for-bit against oracle/xxh64_ref.c.
// K2  diff_select  : ordered compaction of the changed flags into a survivor index list.
//
// Why K1 looks the way it does (measurements in profiles/r01_chain_probe.txt, r01_k1_variants.txt):
//  * XXH64 is four dependent chains per block (acc = rotl(acc + x*P2, 31) * P1 per 8 input bytes per chain; rotl
//    breaks associativity) so the only parallelism inside a block is 4.  Parallelism comes from hashing many blocks at
//    once: one QUAD (4 lanes) per block, 8 blocks per chain warp, 4 chain warps per CTA (one per SM sub-partition), one
//    persistent CTA per SM -> 32 blocks in flight per SM, 4 736 per GPU.
//  * With ~17 blocks per SM at the 10 GiB bench size the kernel is bound by the LATENCY of that chain, so the chain is
//    what is engineered.  On B200 IMAD/IADD3/SHF have 4.4-cycle dependent latency but IMAD.WIDE / IMAD.HI (the 32x32->64
//    product a 64-bit multiply needs) take ~12.5 cycles and occupy their unit ~7 cycles.  A round needs two 64-bit
//    multiplies (x*P2 and rotl(..)*P1); only the second depends on the chain.  So the work is split over THREE warp
//    roles per ring (a ring = the 8 quad slots of one chain warp; 12 warps per CTA):
//      - the TMA warp pulls block indices (static first assignment, then a global atomic counter) and refills drained
//        stages with 1-D TMA bulk copies (cp.async.bulk global->shared, SASS UBLKCP) of 1 KiB chunks of each block into
//        a 6-STAGE shared-memory ring, one mbarrier arrive.expect_tx per stage, plus the stage's descriptors;
//      - the PRE-MULTIPLY warp waits for the stage's mbarrier and replaces x by x*P2 in place (LDS.128 / IMAD.WIDE +
//        2 IMAD / STS.128, fully parallel, off every chain);
//      - the CHAIN warp (highest warp id = highest issue priority) runs the chains: per 8 bytes one LDS.64, two funnel
//        shifts, two IMADs and ONE IMAD.WIDE whose 64-bit addend carries x*P2 -- a 24-cycle round in isolation (nvcc's
//        own lowering of the 64-bit expression: 42 cycles), one predicate-free basic block per 1 KiB chunk.
//    Hand-offs: "TMA landed" is an mbarrier (complete_tx needs one); "pre-multiplied" and "drained" are plain
//    shared-memory counters written by one lane and polled with ld.volatile -- an mbarrier arrive/test/wait costs the
//    issuing warp ~150 cycles of blocked issue (profiles/r01_k1_timeline_mbarrier.txt).  Ordering argument for the
//    counters: the writer publishes with __syncwarp + __threadfence_block + st.volatile, the reader polls, then reads
//    the data it guards through the same (generic) proxy; the async-proxy refill of a stage is issued only after the
//    chain warp's hand-back has been observed, i.e. after every generic read of that stage has completed.  Evidence
//    that this holds in practice: tests/test_gpu.py::test_k1_repeated_under_load_from_another_stream (20 launches under
//    SM contention) and ::test_full_size_resident_pass_properties (all 2 560 hashes of the bench batch vs the oracle).
//    There is one __syncthreads (after barrier init) and no CTA-wide barrier afterwards.
//  * Each chain lane walks an 8-byte column with a 32-byte stride; staging through shared memory turns that into
//    contiguous 1 KiB HBM reads (ncu: DRAM bytes == algorithmic bytes to 4 digits) and conflict-free LDS.64 (quad slots
//    are padded by 32 B so the 4 quads of a half-warp hit disjoint bank groups).
//  * No tensor cores: there is no contraction here, only 64-bit integer mul/add/rotate.
//  * Work distribution: quad (cta c, ring w, quad q) starts on block c + G*(w + 4*q) so a small batch spreads over all
//    SMs first, then over the 4 sub-partitions; afterwards quads pull indices from the atomic counter (ragged block
//    lengths balance).
#include "vmig_kernels.cuh"
#include <atomic>
#include <cstdlib>

namespace vmig {

namespace {

constexpr uint64_t P1 = 0x9E3779B185EBCA87ULL;
constexpr uint64_t P2 = 0xC2B2AE3D27D4EB4FULL;
constexpr uint64_t P3 = 0x165667B19E3779F9ULL;
constexpr uint64_t P4 = 0x85EBCA77C2B2AE63ULL;
constexpr uint64_t P5 = 0x27D4EB2F165667C5ULL;

constexpr int kWarps  = 4;      // consumer warps, one per SM sub-partition (+ as many producer warps)
constexpr int kQuads  = 8;      // blocks in flight per warp
constexpr int kChunk  = 1024;   // bytes per bulk copy (32 stripes)
constexpr int kStages = 6;      // ring depth per quad: 1 draining, 1 pre-multiplied, 4 in flight from HBM
constexpr int kQuadStride  = kChunk + 32;              // +32 B: bank-group skew between quads
constexpr int kStageStride = kQuads * kQuadStride;
constexpr int kWarpData    = kStages * kStageStride;
constexpr int kWarpDesc    = kStages * kQuads * 16;     // uint4 chunk descriptors
constexpr int kWarpBars    = kStages * 8 + 16;          // mbarriers full[stage] + the ready / drained counters
constexpr int kWarpSmem    = kWarpData + kWarpDesc + ((kWarpBars + 15) & ~15);
constexpr int kSmemBytes   = kWarps * kWarpSmem;
static_assert(kChunk % 256 == 0, "chunk must be a multiple of 8 stripes");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr uint32_t kFlagFirst = 1u, kFlagLast = 2u, kFlagNone = 4u;
// stage-wide bits, replicated by the TMA warp into all 8 descriptors of a stage so that the chain
// warp can branch on them without a vote: kFlagSlow = some quad has a ragged or last chunk,
// kFlagEnd = every quad is out of work (the ring shuts down on this stage)
constexpr uint32_t kFlagSlow = 8u, kFlagEnd = 16u;

// Timeline instrumentation for probe/k1_trace.cu only (compiled out of libvmig).
#ifdef VMIG_K1_TRACE
__device__ long long* g_k1_trace;      // [3 roles][kTraceChunks][8] clock64 stamps of CTA 0, ring 0
constexpr int kTraceChunks = 256;
#define K1_TRACE_INIT long long* const k1_trace_p = (blockIdx.x == 0 && warp == 0 && lane == 0) ? g_k1_trace : nullptr
#define K1_TRACE(role_, it_, slot_)                                                                  \
    do {                                                                                             \
        if (k1_trace_p && (it_) < (uint32_t)kTraceChunks)                                             \
            k1_trace_p[((role_) * kTraceChunks + (it_)) * 8 + (slot_)] = clock64();                   \
    } while (0)
#else
#define K1_TRACE_INIT do { } while (0)
#define K1_TRACE(role_, it_, slot_) do { } while (0)
#endif

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t xround(uint64_t acc, uint64_t x) { return rotl64(acc + x * P2, 31) * P1; }
__device__ __forceinline__ uint64_t xmerge(uint64_t h, uint64_t v) { return (h ^ xround(0, v)) * P1 + P4; }


// The XXH64 round on pre-multiplied input m = x*P2 (see the header comment).
// Carried state is t = acc + m (pre-rotation).  One step, given the NEXT pre-multiplied word:
//   r       = rotl(t, 31)                          2 x SHF.L.W (funnel shifts)
//   mhv     = m.hi + r.hi*P1.lo + r.lo*P1.hi       2 x IMAD (cross terms; high word only)
//   t'      = r.lo*P1.lo + {m.lo, mhv}             1 x IMAD.WIDE.U32 with 64-bit addend
// Written as mul.lo/mul.hi + add.cc/addc, which ptxas fuses into exactly that IMAD.WIDE
// (23.9 cycles/round measured; every other spelling tried made ptxas split the addend off into
// an IADD3/IADD3.X carry chain or serialise two IMAD.WIDEs: profiles/r01_chain_probe.txt).
struct Chain {
    uint32_t tlo, thi;
    static constexpr uint32_t P1lo = (uint32_t)P1, P1hi = (uint32_t)(P1 >> 32);
    __device__ __forceinline__ void begin(uint64_t acc, uint64_t m0) {
        const uint64_t t = acc + m0;
        tlo = (uint32_t)t; thi = (uint32_t)(t >> 32);
    }
    __device__ __forceinline__ void step(uint64_t m) {
        uint32_t ml, mh, lo, hi, mhv;
        asm("mov.b64 {%0, %1}, %2;" : "=r"(ml), "=r"(mh) : "l"(m));
        const uint32_t rl = __funnelshift_l(thi, tlo, 31);   // (tlo << 31) | (thi >> 1)
        const uint32_t rh = __funnelshift_l(tlo, thi, 31);   // (thi << 31) | (tlo >> 1)
        asm("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(rl), "r"(P1lo));
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(rl), "r"(P1lo));
        asm("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %2, %3, %5;\n\tmad.lo.u32 %0, %1, %4, a;\n\t}"
            : "=r"(mhv) : "r"(rl), "r"(rh), "r"(P1lo), "r"(P1hi), "r"(mh));
        asm("{\n\tadd.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;\n\t}"
            : "=r"(tlo), "=r"(thi) : "r"(lo), "r"(ml), "r"(hi), "r"(mhv));
    }
    __device__ __forceinline__ uint64_t end() const {
        const uint64_t t = (uint64_t)tlo | ((uint64_t)thi << 32);
        return rotl64(t, 31) * P1;
    }
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
__device__ __forceinline__ uint32_t lds_volatile(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile(uint32_t addr, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
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

// x * P2 mod 2^64 on a 2 x 32-bit register pair (1 IMAD.WIDE + 2 IMAD)
__device__ __forceinline__ void mulP2(uint32_t& lo, uint32_t& hi) {
    const uint64_t m = (((uint64_t)hi << 32) | lo) * P2;
    lo = (uint32_t)m; hi = (uint32_t)(m >> 32);
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

__global__ void __launch_bounds__(3 * kWarps * 32, 1)
xxh64_blocks_kernel(const uint8_t* __restrict__ base, const uint64_t* __restrict__ offs,
                    const uint32_t* __restrict__ lens, uint32_t n, uint64_t* __restrict__ hashes,
                    const uint64_t* __restrict__ prior, const uint8_t* __restrict__ prior_valid,
                    uint8_t* __restrict__ changed, uint32_t* __restrict__ work_counter, uint32_t proxy_fence)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t warp_all = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    // Three warps serve each ring (ring r lives on SM sub-partition r):
    //   warp r      TMA warp      : fetches block ids, refills drained stages with bulk copies
    //   warp 4+r    PRE-MULTIPLY  : x -> x*P2 in place as soon as a stage has landed
    //   warp 8+r    CHAIN warp    : the consumer; highest warp id = highest issue priority, so a
    //                               chain never queues behind a burst of pre-multiply IMAD.WIDEs
    // (Tried and rejected, profiles/r01_k1_variants.txt: one producer warp doing both helper jobs by
    // polling could not keep up with a 768-cycle stage; putting the chain warps two-per-scheduler on
    // SMSP 0,1 and all helpers on SMSP 2,3 made two chains fight over one IMAD.WIDE unit.)
    const uint32_t role = warp_all / kWarps;                   // 0 TMA, 1 pre-multiply, 2 chains
    const uint32_t warp = warp_all % kWarps;                   // ring this warp works on
    const uint32_t quad = lane >> 2, k = lane & 3u;
    const bool leader = (k == 0);
    const uint32_t quad_mask = 0xFu << (lane & ~3u);

    uint8_t* wbase = smem + warp * kWarpSmem;
    const uint32_t data_s = smem_u32(wbase);
    uint4* descs = reinterpret_cast<uint4*>(wbase + kWarpData);
    const uint32_t bars_s = smem_u32(wbase + kWarpData + kWarpDesc);
    // Hand-offs.  TMA completion needs an mbarrier (complete_tx); the other two are plain counters in
    // shared memory: an mbarrier arrive / test / wait costs the issuing warp ~150 cycles of blocked
    // issue (profiles/r01_k1_timeline_mbarrier.txt: 3 such ops per 1 KiB stage cost the chains as
    // much as 20 of their 32 rounds), a volatile LDS/STS costs it nothing.
    const uint32_t full_s = bars_s;
    const uint32_t ready_cnt_s = bars_s + 8 * kStages;        // stages pre-multiplied so far (pre-multiply lane 0 writes)
    const uint32_t done_cnt_s  = ready_cnt_s + 4;              // stages drained so far       (chain lane 0 writes)

    if (role == 0 && lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) mbar_init(full_s + 8 * s, 1);   // TMA-warp lane 0 arrives with the stage's tx bytes
        sts_volatile(ready_cnt_s, 0); sts_volatile(done_cnt_s, 0);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();                              // the only CTA-wide barrier
    K1_TRACE_INIT;

    if (role == 0) {
        // =========================== TMA WARP: fetch block ids, refill ===========================
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
        // one chunk per quad into `stage` (quad-uniform state; the leader lane acts)
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
            const uint32_t bar = full_s + 8 * stage;
            uint32_t nb = 0, flags = kFlagNone;
            if (!iss_done) {
                nb = iss_rem < (uint32_t)kChunk ? iss_rem : (uint32_t)kChunk;
                flags = (iss_first ? kFlagFirst : 0u) | (iss_rem == nb ? kFlagLast : 0u);
            }
            const bool slow = __any_sync(0xFFFFFFFFu, !iss_done && (nb != (uint32_t)kChunk || (flags & kFlagLast)));
            const bool endd = __all_sync(0xFFFFFFFFu, iss_done);
            const uint32_t wide = (slow ? kFlagSlow : 0u) | (endd ? kFlagEnd : 0u);
            if (leader) descs[stage * kQuads + quad] = iss_done ? make_uint4(0, 0, kFlagNone | wide, 0) : make_uint4(nb, iss_blk, flags | wide, iss_len);
            // one arrive for the whole stage (sum of the 8 quads' bytes), then the bulk copies
            uint32_t tot;
            if (!slow) {                                   // every active quad moves a full chunk
                tot = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, leader && !iss_done)) * (uint32_t)kChunk;
            } else {
                tot = leader ? nb : 0u;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
            }
            __syncwarp();                                  // the leaders' descriptor stores precede lane 0's (releasing) arrive
            if (lane == 0) { if (tot) mbar_arrive_expect_tx(bar, tot); else mbar_arrive(bar); }
            __syncwarp();                                  // ... and the expect_tx precedes every bulk copy of the stage
            if (leader && nb) bulk_g2s(data_s + stage * kStageStride + quad * kQuadStride, base + iss_off, nb, bar);
            if (!iss_done) {
                iss_off += nb; iss_rem -= nb; iss_first = false;
                if (flags & kFlagLast) iss_need_fetch = true;
            }
        };

#pragma unroll
        for (int s = 0; s < kStages; s++) issue(s);
        for (uint32_t rf = 0;; rf++) {
            if (__all_sync(0xFFFFFFFFu, iss_done)) break;      // an all-NONE stage is out: everyone stops there
            const int stage = rf % kStages;
            while (lds_volatile(done_cnt_s) <= rf) __nanosleep(200);           // the chains drained chunk rf (a poll every ~400 cycles: the ring has 6 stages of slack, the chain warp's scheduler has none)
            K1_TRACE(0, rf + kStages, 0);
            issue(stage);                                                      // chunk rf + kStages
            K1_TRACE(0, rf + kStages, 1);
        }
        return;
    }

    if (role == 1) {
        // ================================= PRE-MULTIPLY WARP =================================
        for (uint32_t pm = 0;; pm++) {
            const int stage = pm % kStages;
            mbar_wait(full_s + 8 * stage, (pm / kStages) & 1u);                // TMA bytes have landed
            K1_TRACE(1, pm, 0);
            // one descriptor read for the whole stage: lane l looks at quad (l & 7); ballots turn the
            // eight byte counts into warp-uniform row masks (a quad slot = two 512-byte rows)
            const uint4 dq = descs[stage * kQuads + (lane & 7u)];
            const bool none = __all_sync(0xFFFFFFFFu, (dq.z & kFlagNone) != 0);
            const uint32_t row0 = __ballot_sync(0xFFFFFFFFu, dq.x > 0u) & 0xFFu;     // quads with >= 1 row
            const uint32_t row1 = __ballot_sync(0xFFFFFFFFu, dq.x > 512u) & 0xFFu;   // quads with 2 rows
            K1_TRACE(1, pm, 2);
            if (!none) {
                // x -> x*P2 in place, 16 bytes per lane per row.  Rows are processed whole (nbytes is a
                // multiple of 32; bytes of a last partial row beyond nbytes are never read by the
                // chains).  Two quad slots per step: up to 4 x LDS.128 issued together, then up to
                // 8 x (IMAD.WIDE + 2 IMAD), then the STS.128 -- every predicate is warp-uniform.
                static_assert(kChunk == 1024 && kQuads == 8, "pre-multiply is written for 8 quads x 2 rows of 512 B");
                uint4* const stage_rows = reinterpret_cast<uint4*>(wbase + stage * kStageStride) + lane;
                const bool ragged = __any_sync(0xFFFFFFFFu, (dq.z & kFlagSlow) != 0);
#pragma unroll
                for (int q = 0; q < kQuads; q += 2) {
                    if (((row0 >> q) & 3u) == 0u) continue;               // both quad slots idle
                    uint4* const r0 = stage_rows + q * (kQuadStride / 16);
                    uint4* const r1 = r0 + kQuadStride / 16;
                    if (!ragged) {
                        // steady state: every busy quad has a full chunk; an idle partner slot is
                        // multiplied too (harmless) so that the step is one branch-free block
                        uint4 va0 = r0[0], va1 = r0[32], vb0 = r1[0], vb1 = r1[32];
                        mulP2(va0.x, va0.y); mulP2(va0.z, va0.w); mulP2(va1.x, va1.y); mulP2(va1.z, va1.w);
                        mulP2(vb0.x, vb0.y); mulP2(vb0.z, vb0.w); mulP2(vb1.x, vb1.y); mulP2(vb1.z, vb1.w);
                        r0[0] = va0; r0[32] = va1; r1[0] = vb0; r1[32] = vb1;
                    } else {
                        const bool a0 = (row0 >> q) & 1u, a1 = (row1 >> q) & 1u, b0 = (row0 >> (q + 1)) & 1u, b1 = (row1 >> (q + 1)) & 1u;
                        uint4 va0 = make_uint4(0, 0, 0, 0), va1 = va0, vb0 = va0, vb1 = va0;
                        if (a0) va0 = r0[0];
                        if (a1) va1 = r0[32];
                        if (b0) vb0 = r1[0];
                        if (b1) vb1 = r1[32];
                        if (a0) { mulP2(va0.x, va0.y); mulP2(va0.z, va0.w); }
                        if (a1) { mulP2(va1.x, va1.y); mulP2(va1.z, va1.w); }
                        if (b0) { mulP2(vb0.x, vb0.y); mulP2(vb0.z, vb0.w); }
                        if (b1) { mulP2(vb1.x, vb1.y); mulP2(vb1.z, vb1.w); }
                        if (a0) r0[0] = va0;
                        if (a1) r0[32] = va1;
                        if (b0) r1[0] = vb0;
                        if (b1) r1[32] = vb1;
                    }
                }
            }
            K1_TRACE(1, pm, 3);
            // No fence.proxy.async by default: the async-proxy (TMA) refill of these bytes is issued only after
            // the chain warp has READ what is stored here and the TMA warp has seen its hand-back, so the
            // stores are long performed.  VMIG_K1_PROXY_FENCE=1 puts the fence in (a warp-uniform kernel argument) so
            // that its cost can be measured and the two variants compared bit for bit: profiles/r02_k1_proxy_fence.txt.
            if (proxy_fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) { __threadfence_block(); sts_volatile(ready_cnt_s, pm + 1); }   // release: stage pm is consumable
            K1_TRACE(1, pm, 1);
            if (none) break;                                                   // the chains exit on the same stage
        }
        return;
    }

    // ============================ CHAIN WARP: the four chains per block ============================
    // Steady state (every quad of the ring is in the middle of a block, or idle): one loop iteration
    // drains one 1 KiB chunk as a single basic block -- 32 chain steps with everything else (the next
    // groups' LDS.64, handing the stage back, fetching the next stage's descriptor and first 8 words)
    // scheduled into the latency shadow of the chain, whose state t (pre-rotation) is carried across
    // chunks.  Measured before this restructuring: 21 cycles per round inside a chunk but ~500 cycles
    // of bookkeeping between chunks (profiles/r01_k1_timeline_counters.txt).
    // Idle quads run the same code on whatever their slot holds; the result is never looked at.
    Chain c; c.tlo = 0; c.thi = 0;
    while (lds_volatile(ready_cnt_s) == 0) { }
    // seed 0: v1 = P1+P2, v2 = P2, v3 = 0, v4 = -P1
    const uint64_t acc0 = (k == 0) ? (P1 + P2) : (k == 1) ? P2 : (k == 2) ? 0ULL : (0ULL - P1);
    const uint32_t sp_lane = data_s + quad * kQuadStride + k * 8;
    uint32_t it = 0, stage = 0, sp = sp_lane;
    uint4 d = descs[quad];
    uint64_t xa[8];
#pragma unroll
    for (int j = 0; j < 8; j++) xa[j] = lds64(sp + 32 * j);
    for (;;) {
        if (d.z & kFlagEnd) break;                                       // nothing left for this ring
        K1_TRACE(2, it, 0);
        const uint32_t nstage = stage + 1 == (uint32_t)kStages ? 0u : stage + 1;
        const uint32_t nsp = sp_lane + nstage * kStageStride;
        uint4 nd;
        uint64_t nxa[8];
        uint32_t next_ready;
        if (!(d.z & kFlagSlow)) {
            static_assert(kChunk == 1024, "the steady-state path is written for 4 groups of 8 stripes");
            uint64_t xb[8], xc[8], xd[8];
#pragma unroll
            for (int j = 0; j < 8; j++) xb[j] = lds64(sp + 256 + 32 * j);
            {
                Chain cb; cb.begin(acc0, xa[0]);
                c.step(xa[0]);
                if (d.z & kFlagFirst) c = cb;
            }
#pragma unroll
            for (int j = 1; j < 8; j++) c.step(xa[j]);
#pragma unroll
            for (int j = 0; j < 8; j++) xc[j] = lds64(sp + 512 + 32 * j);
#pragma unroll
            for (int j = 0; j < 8; j++) c.step(xb[j]);
#pragma unroll
            for (int j = 0; j < 8; j++) xd[j] = lds64(sp + 768 + 32 * j);
            const uint32_t ready_seen = lds_volatile(ready_cnt_s);
#pragma unroll
            for (int j = 0; j < 8; j++) c.step(xc[j]);
            // all of this stage's words are in registers: hand the stage back ...
            __syncwarp();
            if (lane == 0) sts_volatile(done_cnt_s, it + 1);
            K1_TRACE(2, it, 1);
            // ... and fetch the next stage's descriptor and first group speculatively (redone below
            // in the rare case that it had not been pre-multiplied yet)
            nd = descs[nstage * kQuads + quad];
#pragma unroll
            for (int j = 0; j < 8; j++) nxa[j] = lds64(nsp + 32 * j);
#pragma unroll
            for (int j = 0; j < 8; j++) c.step(xd[j]);
            next_ready = ready_seen > it + 1;                            // same value in every lane
        } else {
            // some quad has a ragged chunk (end of a block whose length is not a multiple of the chunk,
            // or a short block) or finishes its block here: per-quad control flow, once per block
            const bool active = !(d.z & kFlagNone), first = (d.z & kFlagFirst) != 0;
            uint32_t rounds = active ? (d.x >> 5) : 0u;
            uint32_t p = sp;
            if (rounds) {
                if (first) c.begin(acc0, lds64(p)); else c.step(lds64(p));
                p += 32; rounds -= 1;
                for (; rounds; rounds--) { c.step(lds64(p)); p += 32; }
            }
            if (active && (d.z & kFlagLast)) {
                // c holds t of the last stripe (or nothing for a block shorter than one stripe)
                const uint64_t acc = (d.w >= 32) ? c.end() : 0ULL;
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
            __syncwarp();                                      // every lane is done reading this stage
            if (lane == 0) sts_volatile(done_cnt_s, it + 1);   // hand it back to the TMA warp
            next_ready = 0;
        }
        if (!next_ready) {                                     // (warp-uniform)
            while (lds_volatile(ready_cnt_s) <= it + 1) { }
            nd = descs[nstage * kQuads + quad];
#pragma unroll
            for (int j = 0; j < 8; j++) nxa[j] = lds64(nsp + 32 * j);
        }
        d = nd;
#pragma unroll
        for (int j = 0; j < 8; j++) xa[j] = nxa[j];
        sp = nsp; stage = nstage; it++;
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
    static std::atomic<uint64_t> attr_set{0};     // bit d: dynamic shared memory opt-in done on device d
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 64 || !(attr_set.load(std::memory_order_acquire) & (1ull << dev))) {
        e = cudaFuncSetAttribute(xxh64_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        if (dev < 64) attr_set.fetch_or(1ull << dev, std::memory_order_release);
    }
    e = cudaMemsetAsync(a.work_counter, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    // persistent grid: one CTA per SM for a large batch.  A small batch (the streaming engine launches once per staging
    // slot: 8 blocks) gets one CTA per FOUR blocks, not per block: the first assignment puts block c + G*(ring + 4*quad) on
    // CTA c, so with G = ceil(n/4) every CTA runs one block on each of its four rings (one per SM sub-partition) -- the
    // chains do not share an issue port, the launch takes exactly as long (one block's chain latency), and it occupies a
    // quarter of the SMs: 16 slots in flight need 32 SMs instead of 128, so several lanes (or a tenant's kernels) fit
    // beside them instead of queueing (profiles/r02_sweep_lanes_1gpu.txt: 55 ms of hash latency with 4 lanes).
    uint32_t grid = (uint32_t)sm_count;
    const uint32_t want = (a.n + 3u) / 4u;
    if (want < grid) grid = want ? want : 1u;
    static const uint32_t proxy_fence = [] { const char* v = getenv("VMIG_K1_PROXY_FENCE"); return (uint32_t)(v && *v == '1'); }();
    xxh64_blocks_kernel<<<grid, 3 * kWarps * 32, kSmemBytes, st>>>(a.base, a.offs, a.lens, a.n, a.hashes, a.prior,
                                                              a.prior_valid, a.changed, a.work_counter, proxy_fence);
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
