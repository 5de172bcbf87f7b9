// How fast is plain FP64 on this GPU?  Throughput of independent DADD / DSETP+select chains and the latency of a
// dependent chain, next to the same in int32 -- the double-precision SW kernel (1.6 for 'N', bwt.hrl:15) lives on this.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int CHAINS>
__global__ void __launch_bounds__(256) adds(int iters, T seed, T *sink) {
    T a[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) a[k] = seed + (T)(k + threadIdx.x);
    const T b = seed * (T)0.5 + (T)1;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < CHAINS; k++) a[k] = a[k] + b;
#pragma unroll
        for (int k = 0; k < CHAINS; k++) a[k] = a[k] - seed;
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s += a[k];
    if (s == (T)12345) sink[0] = s;
}

// the SW cell in doubles: three adds, two compare+select pairs, a running maximum
template <typename T>
__global__ void __launch_bounds__(128) cells(int iters, T seed, T *sink) {
    T left = seed, diag = seed + 1, up = seed + 2, mx = seed;
    int dir = 0;
    for (int it = 0; it < iters; it++) {
        const T dg = diag + (T)2, uc = up + (dir == 1 ? (T)-1 : (T)-2);
        const bool uw = uc > dg;
        const T m = uw ? uc : dg;
        const int md = uw ? 1 : 2;
        const T lc = left + (dir == 0 ? (T)-1 : (T)-2);
        const bool lw = lc > m;
        left = lw ? lc : m;
        dir = lw ? 0 : md;
        if (left > mx) mx = left;
        diag = up; up = left + (T)it;
    }
    if (mx == (T)12345 && dir == 7) sink[0] = left;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    double *sd; int *si; cudaMalloc(&sd, 8); cudaMalloc(&si, 4);
    const int iters = 4096, blocks = 148 * 8;
    float ms = timeit([&] { adds<double, 8><<<blocks, 256>>>(iters, 3.0, sd); });
    printf("DADD throughput, 8 chains x 2048 threads/SM: %.1f Gop/s  (%.1f per clk per SM at 1.965 GHz)\n", (double)blocks * 256 * iters * 16 / ms / 1e6, (double)blocks * 256 * iters * 16 / ms / 1e6 / 148 / 1.965);
    ms = timeit([&] { adds<int, 8><<<blocks, 256>>>(iters, 3, si); });
    printf("IADD throughput: %.1f Gop/s\n", (double)blocks * 256 * iters * 16 / ms / 1e6);
    ms = timeit([&] { adds<double, 1><<<148, 32>>>(iters * 8, 3.0, sd); });
    printf("DADD dependent chain, one warp per SM: %.1f ns per add = %.1f cycles\n", ms * 1e6 / (iters * 8 * 2), ms * 1e6 / (iters * 8 * 2) * 1.965);
    ms = timeit([&] { adds<int, 1><<<148, 32>>>(iters * 8, 3, si); });
    printf("IADD dependent chain: %.1f cycles\n", ms * 1e6 / (iters * 8 * 2) * 1.965);
    ms = timeit([&] { cells<double><<<148, 32>>>(iters * 8, 3.0, sd); });
    printf("SW cell in doubles, dependent, one warp per SM: %.1f cycles per cell\n", ms * 1e6 / (iters * 8) * 1.965);
    ms = timeit([&] { cells<int><<<148, 32>>>(iters * 8, 3, si); });
    printf("SW cell in int32: %.1f cycles per cell\n", ms * 1e6 / (iters * 8) * 1.965);
    ms = timeit([&] { cells<double><<<148 * 16, 128>>>(iters, 3.0, sd); });
    printf("SW cell in doubles, 2048 threads/SM: %.1f Gcell/s\n", (double)148 * 16 * 128 * iters / ms / 1e6);
    ms = timeit([&] { cells<int><<<148 * 16, 128>>>(iters, 3, si); });
    printf("SW cell in int32, 2048 threads/SM: %.1f Gcell/s\n", (double)148 * 16 * 128 * iters / ms / 1e6);
    return 0;
}
