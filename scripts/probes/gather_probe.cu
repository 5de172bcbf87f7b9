// Random-access probe: how many DRAM bytes does one L2 miss move on this GPU, and is random access bound by the
// number of accesses or by bytes?  Variants: bytes fetched per access (32 / 64 / 128), load flavour (plain __ldg,
// ld.global.nc.L2::64B / L2::128B / L2::256B prefetch hints), and cudaLimitMaxL2FetchGranularity (argv[1]).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

template <int HINT>
__device__ __forceinline__ uint4 ld(const uint4 *p) {
    uint4 v;
    if (HINT == 0) v = __ldg(p);
    else if (HINT == 64) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (HINT == 128) asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (HINT == 256) asm volatile("ld.global.nc.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// BYTES per access: 16 (one LDG.128), 32 (one sector, two loads), 64 (two sectors), 128 (four sectors)
template <int BYTES, int HINT>
__global__ void __launch_bounds__(256) gather(const uint4 *__restrict__ buf, uint64_t mask128, int iters, uint32_t *sink) {
    uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    x = x * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    uint32_t acc = 0;
    for (int it = 0; it < iters; it += 8) {
        uint4 a[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x = x * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t s = (x >> 24) & mask128;            // a random 128-byte line; the access starts at its first byte
            const uint4 *p = buf + 8 * s;
            a[u] = ld<HINT>(p);
            if (BYTES >= 32) { uint4 t = ld<HINT>(p + 1); a[u].x ^= t.w; }
            if (BYTES >= 64) { uint4 t = ld<HINT>(p + 2); a[u].y ^= t.z; t = ld<HINT>(p + 3); a[u].z ^= t.y; }
            if (BYTES >= 128) {
#pragma unroll
                for (int k = 4; k < 8; k++) { uint4 t = ld<HINT>(p + k); a[u].w ^= t.x; }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) acc += a[u].x ^ a[u].y ^ a[u].z ^ a[u].w;
    }
    if (acc == 0x9999999u) sink[0] = acc;
}

template <int BYTES, int HINT>
void run(const char *name, const uint4 *buf, uint64_t lines, uint32_t *sink) {
    const int blocks = 148 * 32, iters = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<BYTES, HINT><<<blocks, 256>>>(buf, lines - 1, 64, sink);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        gather<BYTES, HINT><<<blocks, 256>>>(buf, lines - 1, iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double acc = (double)blocks * 256 * iters;
    printf("%-28s %3d B/access: %7.2f ms  %6.2f G accesses/s  %7.1f GB/s useful\n", name, BYTES, best, acc / best / 1e6, acc * BYTES / best / 1e6);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main(int argc, char **argv) {
    if (argc > 1) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
        printf("cudaLimitMaxL2FetchGranularity <- %s: %s\n", argv[1], cudaGetErrorString(e));
    }
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("L2 fetch granularity limit: %zu\n", g);
    const uint64_t bytes = 16ull << 30, lines = bytes / 128;
    uint4 *buf; uint32_t *sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    run<16, 0>("ldg", buf, lines, sink);
    run<32, 0>("ldg", buf, lines, sink);
    run<64, 0>("ldg", buf, lines, sink);
    run<128, 0>("ldg", buf, lines, sink);
    run<32, 64>("L2::64B", buf, lines, sink);
    run<64, 64>("L2::64B", buf, lines, sink);
    run<32, 128>("L2::128B", buf, lines, sink);
    run<64, 128>("L2::128B", buf, lines, sink);
    run<64, 256>("L2::256B", buf, lines, sink);
    run<32, 1>("L1::no_allocate", buf, lines, sink);
    run<64, 1>("L1::no_allocate", buf, lines, sink);
    return 0;
}
