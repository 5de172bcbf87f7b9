// Roofline denominators measured on the device at hand (SURVEY.md 8d asks for both; nothing is assumed):
//   * random 32-byte-sector gather rate: the ceiling of table-entry and Occ-block fetches, addresses from a
//     per-thread LCG over a buffer far larger than L2
//   * DPX issue rate: independent chains of VIADDMNMX (__viaddmax_s32), the instruction the SW cell is made of
#include "common.cuh"

namespace {

// Random gather of 32-byte sectors, two LDG.128 into each with the L2::64B qualifier -- exactly how sga_kernel reads a
// fused table entry and occ_at() an Occ block.  n_mask + 1 = number of sectors, a power of two: the address is a mask of
// the generator's high bits (the round-1 kernel spent more issue slots on a 64-bit modulo than on the loads, kept four
// sectors in flight for 2 ms and was beaten by production kernels).  Eight independent accesses in flight per thread,
// >= 20 ms per launch.  What it measures is a rate of ACCESSES (scripts/probes/gather_probe.cu: the time per access does
// not depend on the DRAM bytes a miss moves, 69 B here, 128 B without the qualifier, but on the load instructions per
// access): bwt_peak_gather_gbs reports accesses/s x 32 B.
__device__ __forceinline__ uint4 peak_ldg64(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(256) gather_kernel(const uint4 *__restrict__ buf, uint64_t n_mask, int iters,
                                                     uint32_t *__restrict__ sink) {
    uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    x = x * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    uint32_t acc = 0;
    for (int it = 0; it < iters; it += 8) {
        uint4 a[8], b[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x = x * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t s = (x >> 24) & n_mask;
            a[u] = peak_ldg64(buf + 2 * s);
            b[u] = peak_ldg64(buf + 2 * s + 1);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) acc += (a[u].x ^ b[u].w) + (a[u].y ^ b[u].z) + (a[u].z ^ b[u].y) + (a[u].w ^ b[u].x);
    }
    if (acc == 0x9999999u) sink[0] = acc;
}

__global__ void __launch_bounds__(256) dpx_kernel(int iters, int seed, int *__restrict__ sink) {
    int a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = seed + k + threadIdx.x;
    const int b = seed | 1, c = seed - 7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = __viaddmax_s32(a[k], b, c + k);
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = __viaddmax_s32(a[k], -b, c - k);
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= a[k];
    if (s == 0x7FFFFFF1) sink[0] = s;
}

}  // namespace

extern "C" int bwt_peak_gather_gbs(int device, uint64_t bytes, double *gbs) {
    if (!gbs || bytes < (1ull << 20)) { bwt_set_error("bwt_peak_gather_gbs: bad argument"); return BWT_ERR_ARG; }
    BWT_TRY(index_select_device(device));
    uint4 *buf = nullptr;
    uint32_t *sink = nullptr;
    uint64_t n_sectors = 1;
    while (n_sectors * 2 * 32 <= bytes) n_sectors *= 2;                  // largest power of two that fits
    CU_TRY(cudaMalloc(&buf, n_sectors * 32));
    CU_TRY(cudaMalloc(&sink, 4));
    CU_TRY(cudaMemset(buf, 1, n_sectors * 32));
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    const int iters = 512, blocks = 148 * 32;                            // 0.62 G accesses per launch: ~20 ms
    gather_kernel<<<blocks, 256>>>(buf, n_sectors - 1, 64, sink);       // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        CU_TRY(cudaEventRecord(e0));
        gather_kernel<<<blocks, 256>>>(buf, n_sectors - 1, iters, sink);
        CU_TRY(cudaEventRecord(e1));
        CU_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double g = (double)blocks * 256 * iters * 32.0 / (ms * 1e-3) / 1e9;
        if (g > best) best = g;
    }
    *gbs = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(sink);
    return BWT_OK;
}

extern "C" int bwt_peak_dpx_gops(int device, double *gops) {
    if (!gops) { bwt_set_error("bwt_peak_dpx_gops: bad argument"); return BWT_ERR_ARG; }
    BWT_TRY(index_select_device(device));
    int *sink = nullptr;
    CU_TRY(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    const int iters = 4096, blocks = 148 * 8;
    dpx_kernel<<<blocks, 256>>>(iters, 3, sink);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CU_TRY(cudaEventRecord(e0));
        dpx_kernel<<<blocks, 256>>>(iters, 3 + rep, sink);
        CU_TRY(cudaEventRecord(e1));
        CU_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double g = (double)blocks * 256 * iters * 16.0 / (ms * 1e-3) / 1e9;
        if (g > best) best = g;
    }
    *gops = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    return BWT_OK;
}
