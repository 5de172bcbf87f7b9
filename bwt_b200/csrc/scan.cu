// Exclusive prefix sum over uint32 (reduce / recurse / down-sweep), used for per-read hit offsets,
// CIGAR arena offsets and the Occ checkpoints.  Streaming: reads the data twice, writes it once.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if ((threadIdx.x & 31) >= d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread over the block; returns the block total through *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    __shared__ uint32_t tot;
    uint32_t inc = warp_incl_scan(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = l < SCAN_THREADS / 32 ? wsum[l] : 0;
        uint32_t si = warp_incl_scan(s);
        if (l < SCAN_THREADS / 32) wsum[l] = si - s;
        if (l == SCAN_THREADS / 32 - 1) tot = si;
    }
    __syncthreads();
    uint32_t r = inc - v + wsum[w];
    *total = tot;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t *__restrict__ d, size_t n,
                                                                   uint32_t *__restrict__ sums) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) s += d[base + k];
    uint32_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_down_kernel(uint32_t *__restrict__ d, size_t n,
                                                                 const uint32_t *__restrict__ sums,
                                                                 uint32_t *__restrict__ total_out) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? d[base + k] : 0;
        s += v[k];
    }
    uint32_t tot;
    uint32_t off = block_excl_scan(s, &tot) + (sums ? sums[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) d[base + k] = off;
        off += v[k];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *total_out = off;
}

}  // namespace

size_t scan_tmp_words(size_t n) {
    size_t w = 0;
    while (n > SCAN_TILE) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        w += n;
    }
    return w + 1;
}

int exclusive_scan_u32(uint32_t *d_data, size_t n, uint32_t *d_tmp, uint32_t *d_total, cudaStream_t stream,
                       uint32_t *launches) {
    if (n == 0) {
        if (d_total) CU_TRY(cudaMemsetAsync(d_total, 0, sizeof(uint32_t), stream));
        return BWT_OK;
    }
    size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 1) {
        scan_down_kernel<<<1, SCAN_THREADS, 0, stream>>>(d_data, n, nullptr, d_total);
        if (launches) *launches += 1;
        K_CHECK(stream, "scan_down_kernel");
        return BWT_OK;
    }
    scan_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(d_data, n, d_tmp);
    if (launches) *launches += 1;
    CU_TRY(cudaGetLastError());
    BWT_TRY(exclusive_scan_u32(d_tmp, nb, d_tmp + nb, nullptr, stream, launches));
    scan_down_kernel<<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(d_data, n, d_tmp, d_total);
    if (launches) *launches += 1;
    K_CHECK(stream, "scan_down_kernel");
    return BWT_OK;
}
