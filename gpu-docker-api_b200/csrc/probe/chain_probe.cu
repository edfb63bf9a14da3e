// chain_probe.cu -- one-off microbenchmark (not part of libvmig): dependent-issue latency of the
// integer instructions the XXH64 round is made of, and cycles/round of candidate formulations of
// the round, on one warp per SM sub-partition.  Results: profiles/r01_chain_probe.txt.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_probe chain_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL;
constexpr uint32_t P1lo = (uint32_t)P1, P1hi = (uint32_t)(P1 >> 32), P2lo = (uint32_t)P2, P2hi = (uint32_t)(P2 >> 32);
#define N 4096

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

template <int V> __device__ __forceinline__ void run(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs);

// V0: 32-bit IMAD chain  a = a*K + b
template <> __device__ __forceinline__ void run<0>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t*) {
#pragma unroll 16
    for (int i = 0; i < N; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(P1lo), "r"(b));
}
// V1: IMAD.WIDE chain (lo feeds multiplier)
template <> __device__ __forceinline__ void run<1>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t*) {
    uint64_t w = ((uint64_t)b << 32) | a;
#pragma unroll 16
    for (int i = 0; i < N; i++) { uint32_t lo = (uint32_t)w; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(lo), "r"(P1lo)); }
    a = (uint32_t)w; b = (uint32_t)(w >> 32);
}
// V2: IADD3 chain
template <> __device__ __forceinline__ void run<2>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t*) {
#pragma unroll 16
    for (int i = 0; i < N; i++) asm volatile("{\n\t.reg .u32 t;\n\tadd.u32 t, %0, %1;\n\tadd.u32 %0, t, %2;\n\t}" : "+r"(a) : "r"(b), "r"(c));
}
// V3: SHF chain
template <> __device__ __forceinline__ void run<3>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t*) {
#pragma unroll 16
    for (int i = 0; i < N; i++) asm volatile("shf.l.wrap.b32 %0, %1, %0, 31;" : "+r"(a) : "r"(b));
}
// V4: SHF -> IMAD alternating (cross pipe)
template <> __device__ __forceinline__ void run<4>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t*) {
#pragma unroll 16
    for (int i = 0; i < N / 2; i++) {
        asm volatile("shf.l.wrap.b32 %0, %1, %0, 31;" : "+r"(a) : "r"(b));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(P1lo), "r"(c));
    }
}
// V5: nvcc's own 64-bit round
template <> __device__ __forceinline__ void run<5>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) {
    uint64_t acc = ((uint64_t)b << 32) | a;
#pragma unroll 8
    for (int i = 0; i < N; i++) acc = rotl64(acc + xs[i & 15] * P2, 31) * P1;
    a = (uint32_t)acc; b = (uint32_t)(acc >> 32);
}
// V6: the 4-level PTX round currently in vmig_kernels.cu
template <> __device__ __forceinline__ void run<6>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) {
    uint64_t w = ((uint64_t)b << 32) | a; uint32_t v = c;
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        uint64_t x = xs[i & 15];
        uint32_t xl, xh, mh, tlo, thi, rl, rh;
        asm("mov.b64 {%0, %1}, %2;" : "=r"(xl), "=r"(xh) : "l"(x));
        asm("{\n\t.reg .u32 a;\n\tmul.lo.u32 a, %1, %3;\n\tmad.lo.u32 %0, %2, %4, a;\n\t}" : "=r"(mh) : "r"(xl), "r"(xh), "r"(P2hi), "r"(P2lo));
        asm("{\n\t.reg .u64 t;\n\tmad.wide.u32 t, %2, %3, %4;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(tlo), "=r"(thi) : "r"(xl), "r"(P2lo), "l"(w));
        thi = thi + v + mh;
        rl = __funnelshift_l(thi, tlo, 31); rh = __funnelshift_l(tlo, thi, 31);
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(rl), "r"(P1lo));
        asm("{\n\t.reg .u32 a;\n\tmul.lo.u32 a, %2, %3;\n\tmad.lo.u32 %0, %1, %4, a;\n\t}" : "=r"(v) : "r"(rl), "r"(rh), "r"(P1lo), "r"(P1hi));
    }
    a = (uint32_t)w; b = (uint32_t)(w >> 32); c = v;
}
// V7: 16-bit-split free variant: keep acc as 64-bit, use mul.hi/mul.lo pairs (2 independent muls) + adds
template <> __device__ __forceinline__ void run<7>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) {
    uint32_t tlo = a, thi = b;
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        uint64_t m = xs[i & 15] * P2;                     // off-chain
        uint32_t ml = (uint32_t)m, mhh = (uint32_t)(m >> 32);
        uint32_t rl = __funnelshift_l(thi, tlo, 31), rh = __funnelshift_l(tlo, thi, 31);
        uint32_t lo, hi, u, q;
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(rl), "r"(P1lo));
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(rl), "r"(P1lo));
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(u) : "r"(rh), "r"(P1lo));
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(q) : "r"(rl), "r"(P1hi));
        // t = {lo,hi} + {ml,mhh} + ((u+q)<<32)
        asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(tlo) : "r"(lo), "r"(ml));
        uint32_t s; asm volatile("addc.u32 %0, %1, %2;" : "=r"(s) : "r"(hi), "r"(mhh));
        asm volatile("{\n\t.reg .u32 t;\n\tadd.u32 t, %1, %2;\n\tadd.u32 %0, t, %3;\n\t}" : "=r"(thi) : "r"(s), "r"(u), "r"(q));
    }
    a = tlo; b = thi;
}


// V8: software-pipelined: carried state t (pre-rotation); step consumes m_next = x_next*P2 computed
// one round ahead so that its IMAD.WIDE can issue in the shadow of the on-chain IMAD.WIDE.
__device__ __forceinline__ uint64_t mulP2(uint64_t x) {
    uint32_t xl, xh, lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(xl), "=r"(xh) : "l"(x));
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(xl), "r"(P2lo));
    asm("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %1, %3, %0;\n\tmad.lo.u32 %0, %2, %4, a;\n\t}" : "+r"(hi) : "r"(xl), "r"(xh), "r"(P2hi), "r"(P2lo));
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void step8(uint32_t& tlo, uint32_t& thi, uint64_t m) {
    uint32_t rl = __funnelshift_l(thi, tlo, 31), rh = __funnelshift_l(tlo, thi, 31), wlo, whi, v;
    asm("{\n\t.reg .u64 t;\n\tmad.wide.u32 t, %2, %3, %4;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(wlo), "=r"(whi) : "r"(rl), "r"(P1lo), "l"(m));
    asm("{\n\t.reg .u32 a;\n\tmul.lo.u32 a, %2, %3;\n\tmad.lo.u32 %0, %1, %4, a;\n\t}" : "=r"(v) : "r"(rl), "r"(rh), "r"(P1lo), "r"(P1hi));
    tlo = wlo; thi = whi + v;
}
template <> __device__ __forceinline__ void run<8>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) {
    uint32_t tlo = a, thi = b;
    uint64_t m = mulP2(xs[0]);
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        uint64_t mn = mulP2(xs[(i + 1) & 15]);
        step8(tlo, thi, m);
        m = mn;
    }
    a = tlo; b = thi;
}
// V9: floor: m comes straight from shared memory (as if helper warps had pre-multiplied the chunk)
template <> __device__ __forceinline__ void run<9>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) {
    uint32_t tlo = a, thi = b;
#pragma unroll 8
    for (int i = 0; i < N; i++) step8(tlo, thi, xs[i & 15]);
    a = tlo; b = thi;
}
// V10: like V8 but m computed TWO rounds ahead
template <> __device__ __forceinline__ void run<10>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) {
    uint32_t tlo = a, thi = b;
    uint64_t m0 = mulP2(xs[0]), m1 = mulP2(xs[1]);
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        uint64_t m2 = mulP2(xs[(i + 2) & 15]);
        step8(tlo, thi, m0);
        m0 = m1; m1 = m2;
    }
    a = tlo; b = thi;
}


// V11/V12: the kernel's formulation (IMAD.WIDE addend = wide product of the next input word, cross
// terms joined in the IADD3) with the next word's wide product computed one round ahead.
// V12 adds a fake data dependency (xl ^ (rl & zero)) so that it cannot be scheduled BEFORE the
// on-chain IMAD.WIDE of the current round (the two share one slow unit).
template <int FAKE>
__device__ __forceinline__ void run11(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs, uint32_t zero) {
    uint32_t tlo = a, thi = b;
    uint64_t mw; uint32_t mh;
    {
        uint32_t xl = (uint32_t)xs[0], xh = (uint32_t)(xs[0] >> 32);
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(mw) : "r"(xl), "r"(P2lo));
        asm("{\n\t.reg .u32 a;\n\tmul.lo.u32 a, %1, %3;\n\tmad.lo.u32 %0, %2, %4, a;\n\t}" : "=r"(mh) : "r"(xl), "r"(xh), "r"(P2hi), "r"(P2lo));
    }
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        const uint64_t xn = xs[(i + 1) & 15];
        uint32_t xl = (uint32_t)xn, xh = (uint32_t)(xn >> 32);
        uint32_t rl = __funnelshift_l(thi, tlo, 31), rh = __funnelshift_l(tlo, thi, 31), wlo, whi, v, mhn;
        uint64_t mwn;
        asm("{\n\t.reg .u64 t;\n\tmad.wide.u32 t, %2, %3, %4;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(wlo), "=r"(whi) : "r"(rl), "r"(P1lo), "l"(mw));
        asm("{\n\t.reg .u32 a;\n\tmul.lo.u32 a, %2, %3;\n\tmad.lo.u32 %0, %1, %4, a;\n\t}" : "=r"(v) : "r"(rl), "r"(rh), "r"(P1lo), "r"(P1hi));
        uint32_t xl2 = xl;
        if (FAKE) asm("lop3.b32 %0, %1, %2, %3, 0x78;" : "=r"(xl2) : "r"(xl), "r"(rl), "r"(zero));   // xl ^ (rl & zero)
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(mwn) : "r"(xl2), "r"(P2lo));
        asm("{\n\t.reg .u32 a;\n\tmul.lo.u32 a, %1, %3;\n\tmad.lo.u32 %0, %2, %4, a;\n\t}" : "=r"(mhn) : "r"(xl), "r"(xh), "r"(P2hi), "r"(P2lo));
        tlo = wlo; thi = whi + v + mh;
        mw = mwn; mh = mhn;
    }
    a = tlo; b = thi;
}
template <> __device__ __forceinline__ void run<11>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) { run11<0>(a, b, c, xs, 0); }
__device__ uint32_t g_zero;   // runtime zero the compiler cannot see through
template <> __device__ __forceinline__ void run<12>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) { run11<1>(a, b, c, xs, *(volatile uint32_t*)&g_zero); }
// V13: opaque addend: mw ^= runtime-zero (2 LOP3 off-chain) so ptxas cannot re-associate the two
// wide products; the IADD3 keeps 3 live inputs (w.hi, v, mh) so it cannot absorb the carry add.
template <> __device__ __forceinline__ void run<13>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) {
    const uint32_t z = *(volatile uint32_t*)&g_zero;
    uint32_t tlo = a, thi = b, mwl, mwh, mh;
    auto prep = [&](uint64_t x, uint32_t& ol, uint32_t& oh, uint32_t& omh) {
        uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32), l, h;
        asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(l), "=r"(h) : "r"(xl), "r"(P2lo));
        ol = l ^ z; oh = h ^ z;
        asm("{\n\t.reg .u32 a;\n\tmul.lo.u32 a, %1, %3;\n\tmad.lo.u32 %0, %2, %4, a;\n\t}" : "=r"(omh) : "r"(xl), "r"(xh), "r"(P2hi), "r"(P2lo));
    };
    prep(xs[0], mwl, mwh, mh);
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        uint32_t nl, nh, nmh;
        uint32_t rl = __funnelshift_l(thi, tlo, 31), rh = __funnelshift_l(tlo, thi, 31), wlo, whi, v;
        asm("{\n\t.reg .u64 t, m;\n\tmov.b64 m, {%4, %5};\n\tmad.wide.u32 t, %2, %3, m;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(wlo), "=r"(whi) : "r"(rl), "r"(P1lo), "r"(mwl), "r"(mwh));
        asm("{\n\t.reg .u32 a;\n\tmul.lo.u32 a, %2, %3;\n\tmad.lo.u32 %0, %1, %4, a;\n\t}" : "=r"(v) : "r"(rl), "r"(rh), "r"(P1lo), "r"(P1hi));
        prep(xs[(i + 1) & 15], nl, nh, nmh);
        tlo = wlo; thi = whi + v + mh;
        mwl = nl; mwh = nh; mh = nmh;
    }
    a = tlo; b = thi;
}


// V14: m = x*P2 preloaded (helper-warp design); low half by IMAD (fast), high half by IMAD.HI,
// carry of the low half through a predicate into IADD3.X:
//   tlo' = rl*P1lo + m.lo            IMAD      (4.4)
//   p    = tlo' < m.lo               ISETP     (carry out of the low half)
//   hi   = hi32(rl*P1lo)             IMAD.HI   (slow unit, ~12.5)
//   mhv  = m.hi + rh*P1lo + rl*P1hi  2 x IMAD with addend (hidden under IMAD.HI)
//   thi' = hi + mhv + p              IADD3.X
template <> __device__ __forceinline__ void run<14>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) {
    uint32_t tlo = a, thi = b;
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        const uint64_t m = xs[i & 15];
        const uint32_t ml = (uint32_t)m, mhh = (uint32_t)(m >> 32);
        const uint32_t rl = __funnelshift_l(thi, tlo, 31), rh = __funnelshift_l(tlo, thi, 31);
        uint32_t lo, hi, mhv;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(lo) : "r"(rl), "r"(P1lo), "r"(ml));
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(rl), "r"(P1lo));
        asm("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %2, %3, %5;\n\tmad.lo.u32 %0, %1, %4, a;\n\t}" : "=r"(mhv) : "r"(rl), "r"(rh), "r"(P1lo), "r"(P1hi), "r"(mhh));
        asm("{\n\t.reg .pred p;\n\t.reg .u32 c;\n\tsetp.lt.u32 p, %1, %2;\n\tselp.u32 c, 1, 0, p;\n\tadd.u32 c, c, %3;\n\tadd.u32 %0, c, %4;\n\t}" : "=r"(thi) : "r"(lo), "r"(ml), "r"(hi), "r"(mhv));
        tlo = lo;
    }
    a = tlo; b = thi;
}
// V15: same as V14 but the carry via add.cc/addc on explicit halves (lo product by mul.lo)
template <> __device__ __forceinline__ void run<15>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) {
    uint32_t tlo = a, thi = b;
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        const uint64_t m = xs[i & 15];
        const uint32_t ml = (uint32_t)m, mhh = (uint32_t)(m >> 32);
        const uint32_t rl = __funnelshift_l(thi, tlo, 31), rh = __funnelshift_l(tlo, thi, 31);
        uint32_t lo, hi, mhv;
        asm("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(rl), "r"(P1lo));
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(rl), "r"(P1lo));
        asm("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %2, %3, %5;\n\tmad.lo.u32 %0, %1, %4, a;\n\t}" : "=r"(mhv) : "r"(rl), "r"(rh), "r"(P1lo), "r"(P1hi), "r"(mhh));
        asm("{\n\tadd.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;\n\t}" : "=r"(tlo), "=r"(thi) : "r"(lo), "r"(ml), "r"(hi), "r"(mhv));
    }
    a = tlo; b = thi;
}


// V16: V15's algebra, but the low product uses a REGISTER copy of P1lo that ptxas cannot prove
// equal to the immediate, so {mul.lo, mul.hi} are not fused back into one IMAD.WIDE:
//   hi   = hi32(rl*P1lo)               IMAD.HI (slow unit) -- the only slow op, issued right after SHF
//   lo   = rl*P1lo_reg                 IMAD
//   mhv  = m.hi + rh*P1lo + rl*P1hi    2 x IMAD
//   tlo' = lo + m.lo  (carry)          IADD3
//   thi' = hi + mhv + carry            IADD3.X
template <> __device__ __forceinline__ void run<16>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) {
    const uint32_t p1lo_r = P1lo ^ *(volatile uint32_t*)&g_zero;
    uint32_t tlo = a, thi = b;
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        const uint64_t m = xs[i & 15];
        const uint32_t ml = (uint32_t)m, mhh = (uint32_t)(m >> 32);
        const uint32_t rl = __funnelshift_l(thi, tlo, 31), rh = __funnelshift_l(tlo, thi, 31);
        uint32_t lo, hi, mhv;
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(rl), "r"(P1lo));
        asm("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(rl), "r"(p1lo_r));
        asm("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %2, %3, %5;\n\tmad.lo.u32 %0, %1, %4, a;\n\t}" : "=r"(mhv) : "r"(rl), "r"(rh), "r"(P1lo), "r"(P1hi), "r"(mhh));
        asm("{\n\tadd.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;\n\t}" : "=r"(tlo), "=r"(thi) : "r"(lo), "r"(ml), "r"(hi), "r"(mhv));
    }
    a = tlo; b = thi;
}


// V17/V18: no helper warp -- the chain warp pre-multiplies its own input G rounds ahead (raw x read from shared
// memory with volatile loads so the products cannot be hoisted), V15's chain on the result.  The x*P2 work is
// independent of the chain, so it should fill the chain's stall slots; both share the FMA pipe
// (2 x IMAD.WIDE + 4 x IMAD per round = ~22.8 pipe cycles against a 23.9-cycle dependent path).
__device__ __forceinline__ void premulP2(uint64_t x, uint32_t& ml, uint32_t& mh) {
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
    uint32_t hi;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(ml) : "r"(xl), "r"(P2lo));
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(xl), "r"(P2lo));
    asm("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %1, %3, %5;\n\tmad.lo.u32 %0, %2, %4, a;\n\t}" : "=r"(mh) : "r"(xl), "r"(xh), "r"(P2hi), "r"(P2lo), "r"(hi));
}
__device__ __forceinline__ void round15(uint32_t& tlo, uint32_t& thi, uint32_t ml, uint32_t mhh) {
    const uint32_t rl = __funnelshift_l(thi, tlo, 31), rh = __funnelshift_l(tlo, thi, 31);
    uint32_t lo, hi, mhv;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(rl), "r"(P1lo));
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(rl), "r"(P1lo));
    asm("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %2, %3, %5;\n\tmad.lo.u32 %0, %1, %4, a;\n\t}" : "=r"(mhv) : "r"(rl), "r"(rh), "r"(P1lo), "r"(P1hi), "r"(mhh));
    asm("{\n\tadd.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;\n\t}" : "=r"(tlo), "=r"(thi) : "r"(lo), "r"(ml), "r"(hi), "r"(mhv));
}
template <int G> __device__ __forceinline__ void run_inwarp(uint32_t& a, uint32_t& b, const uint64_t* xs) {
    const volatile uint64_t* vx = xs + (threadIdx.x & 15);   // per-lane addresses: keep it off the uniform datapath
    uint32_t tlo = a, thi = b, ml[G], mh[G];
    uint64_t xr[G];                                   // raw words of the NEXT group, loaded one group early
#pragma unroll
    for (int j = 0; j < G; j++) premulP2(vx[j], ml[j], mh[j]);
#pragma unroll
    for (int j = 0; j < G; j++) xr[j] = vx[(G + j) & 15];
#pragma unroll 1
    for (int i = 0; i < N; i += G) {
        uint32_t nl[G], nh[G]; uint64_t xn[G];
#pragma unroll
        for (int j = 0; j < G; j++) xn[j] = vx[(i + 2 * G + j) & 15];
#pragma unroll
        for (int j = 0; j < G; j++) {
            premulP2(xr[j], nl[j], nh[j]);
            round15(tlo, thi, ml[j], mh[j]);
        }
#pragma unroll
        for (int j = 0; j < G; j++) { ml[j] = nl[j]; mh[j] = nh[j]; xr[j] = xn[j]; }
    }
    a = tlo; b = thi;
}
template <> __device__ __forceinline__ void run<17>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) { run_inwarp<4>(a, b, xs); }
template <> __device__ __forceinline__ void run<18>(uint32_t& a, uint32_t& b, uint32_t& c, const uint64_t* xs) { run_inwarp<8>(a, b, xs); }

template <int V> __global__ void k(uint32_t* out, long long* cyc, uint32_t seed) {
    __shared__ uint64_t xs[32];
    if (threadIdx.x < 32) xs[threadIdx.x] = 0x9E3779B97F4A7C15ULL * (threadIdx.x + seed);
    __syncthreads();
    uint32_t a = threadIdx.x + seed, b = seed * 3 + 1, c = seed * 7 + 5;
    long long t0 = clock64();
    run<V>(a, b, c, xs);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V> void go(const char* name, int per, uint32_t* out, long long* cyc) {
    for (int warps = 1; warps <= 2; warps++) {
        k<V><<<1, 128 * warps>>>(out, cyc, 1); cudaDeviceSynchronize();
        k<V><<<1, 128 * warps>>>(out, cyc, 2); cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s %d warp/SMSP: %.2f cycles per %s\n", name, warps, (double)c / (N / per), per == 1 ? "op" : "pair");
    }
}

int main() {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1 << 12);
    go<0>("IMAD (mad.lo.u32) dependent chain", 1, out, cyc);
    go<1>("IMAD.WIDE.U32 dependent chain", 1, out, cyc);
    go<2>("IADD3 dependent chain", 1, out, cyc);
    go<3>("SHF.L.W dependent chain", 1, out, cyc);
    go<4>("SHF -> IMAD pair", 2, out, cyc);
    go<5>("XXH64 round, nvcc 64-bit lowering", 1, out, cyc);
    go<6>("XXH64 round, 4-level PTX (current kernel)", 1, out, cyc);
    go<7>("XXH64 round, mul.lo/mul.hi split", 1, out, cyc);
    go<8>("XXH64 round, m one round ahead", 1, out, cyc);
    go<9>("XXH64 round, m preloaded (floor)", 1, out, cyc);
    go<10>("XXH64 round, m two rounds ahead", 1, out, cyc);
    go<11>("XXH64 round, kernel form, wide(x) 1 ahead", 1, out, cyc);
    go<12>("XXH64 round, kernel form, 1 ahead + fake dep", 1, out, cyc);
    go<13>("XXH64 round, opaque wide(x) addend, 1 ahead", 1, out, cyc);
    go<14>("XXH64 round, m preloaded, IMAD+IMAD.HI+pred carry", 1, out, cyc);
    go<15>("XXH64 round, m preloaded, mul.lo/hi + add.cc", 1, out, cyc);
    go<16>("XXH64 round, m preloaded, IMAD.HI + IMAD(reg) + add.cc", 1, out, cyc);
    go<17>("XXH64 round, in-warp premul 4 rounds ahead + V15", 1, out, cyc);
    go<18>("XXH64 round, in-warp premul 8 rounds ahead + V15", 1, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
