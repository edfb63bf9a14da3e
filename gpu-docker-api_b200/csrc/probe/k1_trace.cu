// k1_trace.cu -- one-off timeline probe (not part of libvmig): compiles the production kernel with
// VMIG_K1_TRACE and prints, for ring 0 of CTA 0, when each 1 KiB chunk was issued (TMA warp), seen
// landed / made ready (pre-multiply warp) and started / released (chain warp), in SM cycles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DVMIG_K1_TRACE -I.. -o k1_trace k1_trace.cu
#define VMIG_K1_TRACE
#include "../vmig_kernels.cu"
#include <cstdio>
#include <vector>
int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? atoi(argv[1]) : 592; const uint32_t bb = 4u << 20;
    uint8_t* d; cudaMalloc(&d, (size_t)n * bb + 64);
    vmig::launch_splitmix_fill(d, (size_t)n * bb, 7, 0);
    std::vector<uint64_t> offs(n); std::vector<uint32_t> lens(n, bb); for (uint32_t i = 0; i < n; i++) offs[i] = (uint64_t)i * bb;
    uint64_t *doffs, *dh; uint32_t *dlens, *dc; long long* dt;
    cudaMalloc(&doffs, n * 8); cudaMalloc(&dlens, n * 4); cudaMalloc(&dh, n * 8); cudaMalloc(&dc, 256);
    const size_t tn = 3 * vmig::kTraceChunks * 8; cudaMalloc(&dt, tn * 8); cudaMemset(dt, 0, tn * 8);
    cudaMemcpy(doffs, offs.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dlens, lens.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(vmig::g_k1_trace, &dt, sizeof dt);
    vmig::HashLaunch a{d, doffs, dlens, n, dh, nullptr, nullptr, nullptr, dc};
    for (int rep = 0; rep < 2; rep++) { cudaMemset(dt, 0, tn * 8); vmig::launch_xxh64_blocks(a, 148, 0); cudaDeviceSynchronize(); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    std::vector<long long> t(tn); cudaMemcpy(t.data(), dt, tn * 8, cudaMemcpyDeviceToHost);
    auto T = [&](int role, int it, int slot) { return t[((size_t)role * vmig::kTraceChunks + it) * 8 + slot]; };
    const long long t0 = T(2, 0, 0);
    printf("chunk | issue_begin issue_end | landed_seen ready_arrive | chain_start chain_release | chain period | ready->start slack | premul: votes, pairs, sync+fence+flag\n");
    for (int it = 100; it < 124; it++)
        printf("%5d | %10lld %9lld | %11lld %12lld | %11lld %13lld | %5lld | %4lld | %4lld %4lld %4lld\n", it, T(0, it, 0) - t0, T(0, it, 1) - t0, T(1, it, 0) - t0, T(1, it, 1) - t0,
               T(2, it, 0) - t0, T(2, it, 1) - t0, T(2, it, 0) - T(2, it - 1, 0), T(2, it, 0) - T(1, it, 1),
               T(1, it, 2) - T(1, it, 0), T(1, it, 3) - T(1, it, 2), T(1, it, 1) - T(1, it, 3));
    return 0;
}
