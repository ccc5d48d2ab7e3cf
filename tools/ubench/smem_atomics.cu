// Microbenchmark: shared-memory vote throughput on sm_100a (decides K3's vote mechanism).
// Every lane owns a pseudo-random stream of cell indices in a shared-memory tile and applies one "vote" per step.
//   0 atomicAdd u32 (result unused)          1 atomicAdd u32 + atomicOr u32 (two planes: today's k3_scan)
//   2 atomicAdd u64 (result unused)          3 plain RMW u32 (LDS, IADD, STS)
//   4 plain RMW u64 (LDS.64, STS.64)         5 atomicAdd u32, result used
//   6 atomicOr u64 hi | atomicAdd lo via one 64-bit CAS-free trick: red.shared.add.u64 with OR-free payload
//   7 red.global.add.u32 into an L2-resident table   8 atomicAdd u32 sorted (ascending ids per lane group)
// Output: lane-votes per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void bench(uint32_t cells, int iters, uint32_t *gtable, uint32_t gcells, unsigned long long *sink) {
    extern __shared__ __align__(16) uint32_t sm[];
    for (uint32_t i = threadIdx.x; i < cells * 2; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0;
    uint64_t *sm64 = reinterpret_cast<uint64_t *>(sm);
    uint32_t sorted_base = (threadIdx.x & ~7u) * 37u;
    for (int it = 0; it < iters; it++) {
        x = x * 1664525u + 1013904223u;
        uint32_t c = (x >> 8) % cells;
        if (MODE == 8) { sorted_base += 8 * 13; c = (sorted_base + (threadIdx.x & 7) * 11 + ((x >> 20) & 7)) % cells; }
        if (MODE == 0 || MODE == 8) atomicAdd(&sm[c], 0x01000003u);
        if (MODE == 1) { atomicAdd(&sm[c], 0x01000003u); atomicOr(&sm[cells + c], 1u << (x & 31)); }
        if (MODE == 2) atomicAdd(reinterpret_cast<unsigned long long *>(&sm64[c]), 0x0100000300000001ull);
        if (MODE == 3) { uint32_t v = sm[c]; sm[c] = v + 0x01000003u; }
        if (MODE == 4) { uint2 v = *reinterpret_cast<uint2 *>(&sm64[c]); v.x += 0x01000003u; v.y |= 1u << (x & 31); *reinterpret_cast<uint2 *>(&sm64[c]) = v; }
        if (MODE == 5) acc += atomicAdd(&sm[c], 0x01000003u);
        if (MODE == 6) { atomicAdd(&sm[2 * c], 0x01000003u); atomicOr(&sm[2 * c + 1], 1u << (x & 31)); }
        if (MODE == 7) { uint32_t g = (x >> 4) % gcells; atomicAdd(&gtable[g], 1u); }
    }
    __syncthreads();
    unsigned long long s = acc;
    for (uint32_t i = threadIdx.x; i < cells * 2; i += blockDim.x) s += sm[i];
    if (s == 0xdeadbeefcafeull) *sink = s;
}

template <int MODE>
void run(const char *name, int threads, uint32_t cells, int ctas_per_sm, uint32_t *gtable, uint32_t gcells, unsigned long long *sink) {
    const int iters = 4096;
    const size_t smem = (size_t)cells * 8;
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = 148 * ctas_per_sm;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    bench<MODE><<<grid, threads, smem>>>(cells, 64, gtable, gcells, sink);
    cudaEventRecord(a);
    bench<MODE><<<grid, threads, smem>>>(cells, iters, gtable, gcells, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    const double votes = (double)grid * threads * iters;
    const double clk = 1.965e9;
    printf("%-44s thr %4d cells %6u cta/sm %d : %8.3f ms  %7.2f Gvotes/s  %6.3f votes/clk/SM %s\n", name, threads, cells, ctas_per_sm, ms,
           votes / ms * 1e-6, votes / (ms * 1e-3) / clk / 148.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    uint32_t *gtable;
    const uint32_t gcells = 4u << 20; // 16 MB: L2 resident
    cudaMalloc(&gtable, (size_t)gcells * 4);
    cudaMemset(gtable, 0, (size_t)gcells * 4);
    unsigned long long *sink;
    cudaMalloc(&sink, 8);
    for (int threads : {256, 1024}) {
        const int cps = threads == 1024 ? 1 : 4;
        const uint32_t cells = threads == 1024 ? 24576 : 6144;
        run<0>("atomicAdd u32", threads, cells, cps, gtable, gcells, sink);
        run<8>("atomicAdd u32 sorted ids", threads, cells, cps, gtable, gcells, sink);
        run<1>("atomicAdd u32 + atomicOr u32 (2 planes)", threads, cells, cps, gtable, gcells, sink);
        run<6>("atomicAdd + atomicOr (adjacent words)", threads, cells, cps, gtable, gcells, sink);
        run<2>("atomicAdd u64", threads, cells, cps, gtable, gcells, sink);
        run<5>("atomicAdd u32 result used", threads, cells, cps, gtable, gcells, sink);
        run<3>("plain RMW u32", threads, cells, cps, gtable, gcells, sink);
        run<4>("plain RMW u64", threads, cells, cps, gtable, gcells, sink);
        run<7>("red.global u32 (16 MB table)", threads, cells, cps, gtable, gcells, sink);
    }
    return 0;
}
