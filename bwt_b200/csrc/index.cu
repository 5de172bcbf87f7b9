// Index lifetime: parse / upload / device layout (Occ blocks, 8-mer interval table), save, export.
//
// Replaces fm_index:get_index/2 (fm_index.erl:154-156) and the .ref load + 'N' padding of
// worker_bwt.erl:48-55.  HBM layout per partition (n rows, ref_len bases):
//   occ   : ceil(n/64) x 32 B   checkpoint counts + bit planes of L      (0.5 B / row)
//   sa    : n x 4 B             full suffix array, as the reference keeps it (fmi.erl:89-100)
//   kmer  : 65536 x 8 B         interval after the first 8 characters (L2-resident)
//   ref   : ref_len + 400 B     ASCII .ref padded with 200 'N' each side
#include <stdarg.h>
#include <algorithm>
#include <new>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "etf.hpp"
#include "search.cuh"

// ---- thread-local error text ----------------------------------------------------------------------
static thread_local char g_err[512] = "";

void bwt_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

bool bwt_debug_sync() {
    static const bool on = [] { const char *e = getenv("BWT_DEBUG_SYNC"); return e && *e && *e != '0'; }();
    return on;
}

extern "C" const char *bwt_last_error(void) { return g_err; }
extern "C" int bwt_abi_version(void) { return BWT_B200_ABI_VERSION; }

extern "C" int bwt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int index_select_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        bwt_set_error("no CUDA device available (%s): libbwt_b200 has no CPU path",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return BWT_ERR_NODEVICE;
    }
    if (device < 0 || device >= n) { bwt_set_error("device %d out of range (0..%d)", device, n - 1); return BWT_ERR_ARG; }
    CU_TRY(cudaSetDevice(device));
    return BWT_OK;
}

// ---- layout kernels -------------------------------------------------------------------------------
namespace {

// one thread per 64-row block: bit planes + per-block symbol counts ('$' is stored and counted as A)
__global__ void planes_kernel(const uint8_t *__restrict__ L, uint64_t n, uint4 *__restrict__ occ,
                              uint32_t *__restrict__ cA, uint32_t *__restrict__ cC, uint32_t *__restrict__ cG,
                              uint32_t *__restrict__ cT, uint32_t *__restrict__ dollar, uint32_t *__restrict__ bad) {
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t nb = (n + 63) / 64;
    if (b >= nb) return;
    uint32_t lo[2] = {0, 0}, hi[2] = {0, 0}, cnt[4] = {0, 0, 0, 0};
    uint64_t base = b * 64;
    for (int k = 0; k < 64; k++) {
        uint64_t row = base + k;
        if (row >= n) break;
        uint8_t ch = L[row];
        int c = bwt_code(ch);
        if (c < 0) {
            if (ch == '$') { atomicMin(dollar, (uint32_t)(row + 1)); atomicAdd(dollar + 1, 1u); }
            else atomicAdd(bad, 1u);
            c = 0;
        }
        cnt[c]++;
        lo[k >> 5] |= (uint32_t)(c & 1) << (k & 31);
        hi[k >> 5] |= (uint32_t)(c >> 1) << (k & 31);
    }
    occ[2 * b + 1] = make_uint4(lo[0], lo[1], hi[0], hi[1]);
    cA[b] = cnt[0]; cC[b] = cnt[1]; cG[b] = cnt[2]; cT[b] = cnt[3];
}

__global__ void counts_kernel(uint4 *__restrict__ occ, uint64_t nb, const uint32_t *__restrict__ cA,
                              const uint32_t *__restrict__ cC, const uint32_t *__restrict__ cG,
                              const uint32_t *__restrict__ cT) {
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) occ[2 * b] = make_uint4(cA[b], cC[b], cG[b], cT[b]);
}

// 4^8 threads: the first eight characters of bwa:find_seeds/6 for every ACGT 8-mer; entry format in
// common.cuh (the interval at a not_found stop is kept, so the explicit search API can report it)
__global__ void kmer_kernel(DevIndex ix, uint2 *__restrict__ tab) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 65536u) return;
    uint32_t sp, ep;
    first_interval(ix, (int)(idx & 3), sp, ep);
    int n = 1;
    for (int t = 1; t < 8; t++) {
        if (!lf_step(ix, sp, ep, (int)((idx >> (2 * t)) & 3))) break;
        n++;
    }
    tab[idx] = kmer_pack(sp, ep, n);
}

// depth k -> k+1, in place: entry i (the k characters before the end of a subsequence, last character in bits 0-1)
// has the four children i | c << 2k.  One thread per parent: both Occ sectors are fetched once for the four characters.
__global__ void kmer_expand_kernel(DevIndex ix, uint2 *__restrict__ tab, int k) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t cnt = 1ull << (2 * k);
    if (i >= cnt) return;
    uint2 e = tab[i];
    int ne = (int)(e.y >> 28);
    uint2 child[4] = {e, e, e, e};                      // stopped / overflowed parents propagate unchanged
    if (ne == k) {
        const uint32_t sp0 = e.x, ep0 = e.x + (e.y & KMER_W_MASK);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t sp = sp0, ep = ep0;
            child[c] = lf_step(ix, sp, ep, c) ? kmer_pack(sp, ep, k + 1) : e;   // not_found: stays at n = k
        }
    }
#pragma unroll
    for (int c = 1; c < 4; c++) tab[i + (uint64_t)c * cnt] = child[c];
    tab[i] = child[0];
}

// every suffix start must lie inside the text: SA values index the reference
__global__ void sa_check_kernel(const uint32_t *__restrict__ sa, uint64_t n, uint32_t *__restrict__ bad) {
    uint32_t b = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) b |= sa[i] >= n ? 1u : 0u;
    if (__any_sync(0xFFFFFFFFu, b) && (threadIdx.x & 31) == 0) atomicExch(bad, 1u);
}

__global__ void pad_fill_kernel(uint8_t *p, uint64_t n, uint8_t v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// depth-15 interval table -> fused entries: the interval and the first 14 seed ends' SA values in one 64-byte line
// (sector 0: Sp, width|n, SA x 6; sector 1: SA x 8), so that a whole bwa:find_seeds/6 + fmi:get_seed_ends/4 of an
// N-free 15-character subsequence is ONE DRAM access instead of a table entry and a suffix-array run somewhere else.
__global__ void fused_build_kernel(const uint2 *__restrict__ kdeep, const uint32_t *__restrict__ sa, uint4 *__restrict__ fused,
                                   uint64_t entries) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= entries) return;
    const uint2 e = kdeep[i];
    const uint32_t ne = e.y >> 28, w = e.y & KMER_W_MASK;
    uint32_t cnt = 0;
    if (ne == BWT_MAX_LEN) cnt = w > BWT_MAX_RANGE ? 0u : w + 1;          // bwa.erl:31-32 width check
    else if (ne >= BWT_MIN_LEN) cnt = w + 1;                               // not_found stop: no width check (bwa.erl:37-38)
    uint32_t v[14];
#pragma unroll
    for (int t = 0; t < 14; t++) v[t] = (uint32_t)t < cnt ? sa[(size_t)(e.x - 1) + t] : 0u;
    uint4 *o = fused + 4 * i;
    o[0] = make_uint4(e.x, e.y, v[0], v[1]);
    o[1] = make_uint4(v[2], v[3], v[4], v[5]);
    o[2] = make_uint4(v[6], v[7], v[8], v[9]);
    o[3] = make_uint4(v[10], v[11], v[12], v[13]);
}

// one thread per 16-byte word of an ASCII buffer: 2-bit codes (A0 C1 G2 T3, anything else as A) + a flag per word
__global__ void pack_ref_kernel(const uint8_t *__restrict__ a, uint64_t len, uint64_t nwords, uint32_t *__restrict__ ref2,
                                uint32_t *__restrict__ refx) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t out = 0;
    bool bad = false;
    if (w < nwords) {
        const uint64_t base = w * 16;
#pragma unroll
        for (int t = 0; t < 16; t++) {
            int c = 0;
            if (base + t < len) { c = bwt_code(a[base + t]); if (c < 0) { bad = true; c = 0; } }
            out |= (uint32_t)c << (2 * t);
        }
        ref2[w] = out;
    }
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, bad);
    if ((threadIdx.x & 31) == 0 && (w >> 5) < (nwords + 31) / 32 + 1) refx[w >> 5] = m;
}

}  // namespace

int index_finish_from_device(bwt_index *h, const uint8_t *d_L, uint64_t n, uint32_t *d_sa, uint8_t *d_ref_padded,
                             uint64_t ref_len, const uint32_t pointers[4], int64_t shift) {
    if (n < 2 || n >= 0xFFFFFFFFull - (1u << 20)) {
        bwt_set_error("index of %llu rows is outside the supported range (2 .. 2^32-2^20)", (unsigned long long)n);
        return BWT_ERR_LIMIT;
    }
    if (!(pointers[0] >= 2 && pointers[0] < pointers[1] && pointers[1] < pointers[2] && pointers[2] <= pointers[3] &&
          pointers[3] == n)) {
        bwt_set_error("inconsistent pointers {%u,%u,%u,%u} for %llu rows", pointers[0], pointers[1], pointers[2],
                      pointers[3], (unsigned long long)n);
        return BWT_ERR_FORMAT;
    }
    uint64_t nb = (n + 63) / 64;
    uint4 *d_occ = nullptr;
    uint32_t *d_cnt = nullptr, *d_tmp = nullptr, *d_flags = nullptr;
    uint2 *d_kmer = nullptr;   // (unused: the tables are built by bwt_set_kmer_depth)
    if (cudaMalloc(&d_occ, nb * 32) != cudaSuccess || cudaMalloc(&d_cnt, nb * 4 * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&d_tmp, scan_tmp_words(nb) * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&d_flags, 4 * sizeof(uint32_t)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(d_occ); cudaFree(d_cnt); cudaFree(d_tmp); cudaFree(d_flags);
        bwt_set_error("index layout: device allocation for %llu rows failed", (unsigned long long)n);
        return BWT_ERR_NOMEM;
    }
    uint32_t init[4] = {0xFFFFFFFFu, 0, 0, 0};   // dollar row (min), #dollars, #bad symbols, SA value out of range
    CU_TRY(cudaMemcpy(d_flags, init, sizeof init, cudaMemcpyHostToDevice));
    sa_check_kernel<<<148 * 8, 256>>>(d_sa, n, d_flags + 3);
    uint32_t *cA = d_cnt, *cC = d_cnt + nb, *cG = d_cnt + 2 * nb, *cT = d_cnt + 3 * nb;
    planes_kernel<<<(unsigned)div_up_u64(nb, 256), 256>>>(d_L, n, d_occ, cA, cC, cG, cT, d_flags, d_flags + 2);
    CU_TRY(cudaGetLastError());
    for (int c = 0; c < 4; c++) BWT_TRY(exclusive_scan_u32(d_cnt + c * nb, nb, d_tmp, nullptr, 0, nullptr));
    counts_kernel<<<(unsigned)div_up_u64(nb, 256), 256>>>(d_occ, nb, cA, cC, cG, cT);
    CU_TRY(cudaGetLastError());
    uint32_t flags[4];
    uint32_t tot[4] = {0, 0, 0, 0};                 // symbols of L per letter ('$' counted as A): last block's checkpoint + its own count
    {
        cudaError_t e = cudaMemcpy(flags, d_flags, sizeof flags, cudaMemcpyDeviceToHost);
        uint4 last_cp;
        if (e == cudaSuccess) e = cudaMemcpy(&last_cp, d_occ + 2 * (nb - 1), sizeof last_cp, cudaMemcpyDeviceToHost);
        uint4 last_pl;
        if (e == cudaSuccess) e = cudaMemcpy(&last_pl, d_occ + 2 * (nb - 1) + 1, sizeof last_pl, cudaMemcpyDeviceToHost);
        cudaFree(d_cnt); cudaFree(d_tmp); cudaFree(d_flags);
        if (e != cudaSuccess) { cudaFree(d_occ); bwt_set_error("index layout: %s", cudaGetErrorString(e)); return BWT_ERR_CUDA; }
        const uint32_t cp[4] = {last_cp.x, last_cp.y, last_cp.z, last_cp.w};
        const uint64_t rows_last = n - (nb - 1) * 64;
        for (uint64_t k = 0; k < rows_last; k++) {
            const uint32_t lo = ((k < 32 ? last_pl.x : last_pl.y) >> (k & 31)) & 1u, hi = ((k < 32 ? last_pl.z : last_pl.w) >> (k & 31)) & 1u;
            tot[hi * 2 + lo]++;
        }
        for (int c = 0; c < 4; c++) tot[c] += cp[c];
    }
    if (flags[2] != 0 || flags[1] != 1) {
        cudaFree(d_occ);
        bwt_set_error("L column must hold exactly one '$' and otherwise only ACGT (%u '$', %u other symbols)", flags[1], flags[2]);
        return BWT_ERR_FORMAT;
    }
    if (flags[3]) { cudaFree(d_occ); bwt_set_error("SA column holds a value beyond the %llu rows", (unsigned long long)n); return BWT_ERR_FORMAT; }
    // the LF mapping adds occ counts taken from L to the pointers taken from the file: they must describe the same text
    // (F is L sorted), or a step walks off the index
    if ((uint64_t)pointers[0] != 2ull + (tot[0] - 1) || (uint64_t)pointers[1] != (uint64_t)pointers[0] + tot[1] ||
        (uint64_t)pointers[2] != (uint64_t)pointers[1] + tot[2] || (uint64_t)pointers[2] + tot[3] != n + 1) {
        cudaFree(d_occ);
        bwt_set_error("pointers {%u,%u,%u,%u} disagree with the symbol counts of L (A %u, C %u, G %u, T %u)", pointers[0], pointers[1],
                      pointers[2], pointers[3], tot[0] - 1, tot[1], tot[2], tot[3]);
        return BWT_ERR_FORMAT;
    }
    h->n = n;
    h->ref_len = ref_len;
    h->shift = shift;
    memcpy(h->pointers, pointers, sizeof h->pointers);
    h->d_occ = d_occ; h->d_sa = d_sa; h->d_ref = d_ref_padded;
    DevIndex &dv = h->dv;
    dv.occ = d_occ; dv.sa = d_sa; dv.ref = d_ref_padded;
    dv.ref_len_padded = ref_len + 2 * BWT_REF_EXT;
    dv.n = (uint32_t)n;
    dv.P[0] = 2; dv.P[1] = pointers[0]; dv.P[2] = pointers[1]; dv.P[3] = pointers[2];
    dv.last = pointers[3];
    dv.dollar = flags[0];
    dv.shift = shift;
    dv.ktab = nullptr; dv.kdeep = nullptr; dv.fused = nullptr;
    dv.deep_k = 0;
    cudaFree(d_kmer);
    {   // 2-bit copy of the padded reference for the extension kernels' window fetches
        uint32_t *r2 = nullptr, *rx = nullptr;
        BWT_TRY(pack_ref_device(d_ref_padded, dv.ref_len_padded, &r2, &rx, 0));
        CU_TRY(cudaDeviceSynchronize());
        h->d_ref2 = r2; h->d_refx = rx;
        dv.ref2 = r2; dv.refx = rx;
    }
    h->device_bytes = nb * 32 + n * 4 + dv.ref_len_padded + dv.ref_len_padded / 4 + dv.ref_len_padded / 128;
    // deepest table: the largest depth <= 15 whose 4^k entries do not outnumber 4 rows of the index each
    // (GL000193.1: k = 9; 3.1 Gbp: k = 15 = MAX_LEN, a whole search is one lookup); BWT_KMER_K overrides
    int k = 8;
    while (k < BWT_MAX_LEN && (1ull << (2 * (k + 1))) <= 4ull * n) k++;
    if (const char *e = getenv("BWT_KMER_K")) { int v = atoi(e); if (v == 0 || (v >= 8 && v <= BWT_MAX_LEN)) k = v; }
    return bwt_set_kmer_depth(h, k);
}

int pack_ref_device(const uint8_t *d_ascii, uint64_t len, uint32_t **d_ref2, uint32_t **d_refx, cudaStream_t stream) {
    const uint64_t nwords = len / 16 + 4;                 // ref_word() reads up to two words past the last base
    uint32_t *r2 = nullptr, *rx = nullptr;
    const uint64_t nflags = (nwords + 31) / 32 + 2;
    CU_TRY(cudaMalloc(&r2, nwords * 4));
    if (cudaMalloc(&rx, nflags * 4) != cudaSuccess) { cudaGetLastError(); cudaFree(r2); bwt_set_error("pack_ref_device: allocation failed"); return BWT_ERR_NOMEM; }
    cudaMemsetAsync(rx, 0, nflags * 4, stream);
    const uint64_t threads = (nwords + 31) / 32 * 32;
    pack_ref_kernel<<<(unsigned)div_up_u64(threads, 256), 256, 0, stream>>>(d_ascii, len, nwords, r2, rx);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFree(r2); cudaFree(rx); bwt_set_error("pack_ref_kernel: %s", cudaGetErrorString(e)); return BWT_ERR_CUDA; }
    *d_ref2 = r2; *d_refx = rx;
    return BWT_OK;
}

// Interval tables of EVERY depth 8..k (k <= 15): a subsequence of d < 15 plain characters (the short one at the front of
// a read: 8 + Lq mod 8 characters, any depth once read lengths vary) costs one lookup in the table of exactly its depth
// instead of a shallower lookup plus dependent Occ steps.  Depths 8..14 share one allocation (2.9 GB when complete).
// With k = 15 the deepest table becomes the FUSED table (64 B per 15-mer, 68.7 GB) when BWT_FUSED=1, or -- BWT_FUSED
// unset -- when that leaves 40 GB of the device free; BWT_FUSED=0 keeps the plain 8-byte table (8.6 GB).
extern "C" int bwt_set_kmer_depth(bwt_index *h, int k) {
    if (!h || (k != 0 && (k < 8 || k > BWT_MAX_LEN))) { bwt_set_error("bwt_set_kmer_depth: depth must be 0 or 8..15"); return BWT_ERR_ARG; }
    if (k == 0) k = 8;
    CU_TRY(cudaSetDevice(h->device));
    CU_TRY(cudaDeviceSynchronize());
    if (!h->tail_mu) h->tail_mu = new std::mutex();
    auto drop = [&](void *&p, uint64_t bytes) { if (p) { cudaFree(p); p = nullptr; h->device_bytes -= bytes; } };
    if (h->dv.deep_k) {
        drop(h->d_ktab, ktab_off(std::min(h->dv.deep_k, 14) + 1) * 8);
        drop(h->d_kdeep, 8ull << 30);
        drop(h->d_fused, 64ull << 30);
    }
    h->dv.ktab = nullptr; h->dv.kdeep = nullptr; h->dv.fused = nullptr; h->dv.deep_k = 0;
    const int kt = std::min(k, 14);
    uint2 *ktab = nullptr, *work = nullptr;
    const uint64_t kt_entries = ktab_off(kt + 1), work_entries = 1ull << (2 * k);
    if (cudaMalloc(&ktab, kt_entries * sizeof(uint2)) != cudaSuccess || cudaMalloc(&work, work_entries * sizeof(uint2)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(ktab);
        bwt_set_error("k-mer tables of depth %d (%llu MB) do not fit", k, (unsigned long long)((kt_entries + work_entries) * 8 >> 20));
        return BWT_ERR_NOMEM;
    }
    kmer_kernel<<<256, 256>>>(h->dv, work);
    for (int d = 8; d <= k; d++) {
        if (d <= kt) CU_TRY(cudaMemcpyAsync(ktab + ktab_off(d), work, (sizeof(uint2)) << (2 * d), cudaMemcpyDeviceToDevice, 0));
        if (d < k) kmer_expand_kernel<<<(unsigned)div_up_u64(1ull << (2 * d), 256), 256>>>(h->dv, work, d);
        CU_TRY(cudaGetLastError());
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(ktab); cudaFree(work); bwt_set_error("k-mer table build: %s", cudaGetErrorString(e)); return BWT_ERR_CUDA; }
    h->d_ktab = ktab; h->dv.ktab = ktab; h->dv.deep_k = k;
    h->device_bytes += kt_entries * 8;
    if (k < BWT_MAX_LEN) { cudaFree(work); return BWT_OK; }
    h->d_kdeep = work; h->dv.kdeep = work;
    h->device_bytes += 8ull << 30;
    // fused table?
    const char *ef = getenv("BWT_FUSED");
    int want = ef ? atoi(ef) : -1;
    const uint64_t fused_bytes = 64ull << 30;
    if (want < 0) {
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); fr = 0; }
        want = fr >= fused_bytes + (40ull << 30) ? 1 : 0;
    }
    if (!want) return BWT_OK;
    uint4 *fused = nullptr;
    if (cudaMalloc(&fused, fused_bytes) != cudaSuccess) {
        cudaGetLastError();
        if (ef) { bwt_set_error("fused k-mer table (64 GiB) does not fit"); return BWT_ERR_NOMEM; }
        return BWT_OK;
    }
    fused_build_kernel<<<(unsigned)div_up_u64(work_entries, 256), 256>>>(work, h->dv.sa, fused, work_entries);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(fused); bwt_set_error("fused table build: %s", cudaGetErrorString(e)); return BWT_ERR_CUDA; }
    cudaFree(work);
    h->d_kdeep = nullptr; h->dv.kdeep = nullptr;
    h->d_fused = fused; h->dv.fused = fused;
    h->device_bytes += fused_bytes - (8ull << 30);
    return BWT_OK;
}

// pad: d_out[0..200) = 'N', [200..200+len) = ref, tail = 'N'   (worker_bwt.erl:53-55)
static int upload_padded_ref(const uint8_t *ref, uint64_t len, bool on_device, uint8_t **d_out) {
    uint8_t *d = nullptr;
    uint64_t tot = len + 2 * BWT_REF_EXT;
    CU_TRY(cudaMalloc(&d, tot + 16));            // 16 bytes of slack: the extension kernel reads aligned words
    pad_fill_kernel<<<1, 256>>>(d, BWT_REF_EXT, 'N');
    pad_fill_kernel<<<1, 256>>>(d + BWT_REF_EXT + len, BWT_REF_EXT, 'N');
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpy(d + BWT_REF_EXT, ref, len, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    *d_out = d;
    return BWT_OK;
}

int index_upload_ref_device(const uint8_t *d_ref, uint64_t len, uint8_t **d_out) {
    return upload_padded_ref(d_ref, len, true, d_out);
}

extern "C" int bwt_open_arrays(const uint8_t *L, uint64_t n, const uint32_t *SA, const uint32_t pointers[4],
                               int64_t shift, const uint8_t *ref, uint64_t ref_len, int device, bwt_index **out) {
    if (!L || !SA || !pointers || !ref || !out) { bwt_set_error("bwt_open_arrays: null argument"); return BWT_ERR_ARG; }
    BWT_TRY(index_select_device(device));
    bwt_index *h = new bwt_index();
    h->device = device;
    uint8_t *d_L = nullptr, *d_ref = nullptr;
    uint32_t *d_sa = nullptr;
    int st = BWT_OK;
    do {
        if (cudaMalloc(&d_L, n) != cudaSuccess || cudaMalloc(&d_sa, n * sizeof(uint32_t)) != cudaSuccess) {
            bwt_set_error("device allocation of %llu index rows failed: %s", (unsigned long long)n,
                          cudaGetErrorString(cudaGetLastError()));
            st = BWT_ERR_NOMEM; break;
        }
        if (cudaMemcpy(d_L, L, n, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(d_sa, SA, n * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            bwt_set_error("index upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = BWT_ERR_CUDA; break;
        }
        st = upload_padded_ref(ref, ref_len, false, &d_ref);
        if (st != BWT_OK) break;
        st = index_finish_from_device(h, d_L, n, d_sa, d_ref, ref_len, pointers, shift);
    } while (0);
    if (d_L) cudaFree(d_L);
    if (st != BWT_OK) {
        if (d_sa) cudaFree(d_sa);
        if (d_ref) cudaFree(d_ref);
        delete h;
        return st;
    }
    *out = h;
    return BWT_OK;
}

static int read_file(const char *path, std::vector<uint8_t> &buf) {
    FILE *f = fopen(path, "rb");
    if (!f) { bwt_set_error("cannot open %s", path); return BWT_ERR_IO; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)sz);
    size_t got = sz ? fread(buf.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) { bwt_set_error("short read on %s", path); return BWT_ERR_IO; }
    return BWT_OK;
}

static int bwt_open_impl(const char *fm_path, const char *ref_path, int device, bwt_index **out);

extern "C" int bwt_open(const char *fm_path, const char *ref_path, int device, bwt_index **out) {
    if (!fm_path || !ref_path || !out) { bwt_set_error("bwt_open: null argument"); return BWT_ERR_ARG; }
    try {
        return bwt_open_impl(fm_path, ref_path, device, out);
    } catch (const std::bad_alloc &) {         // never across the extern "C" boundary
        bwt_set_error("bwt_open: out of host memory reading %s", fm_path);
        return BWT_ERR_NOMEM;
    }
}

static int bwt_open_impl(const char *fm_path, const char *ref_path, int device, bwt_index **out) {
    BWT_TRY(index_select_device(device));   // fail before any file work when there is no GPU
    std::vector<uint8_t> fm, ref;
    BWT_TRY(read_file(fm_path, fm));
    BWT_TRY(read_file(ref_path, ref));
    EtfIndex ix;
    std::string err;
    if (!etf_parse_fm(fm.data(), fm.size(), ix, err)) {
        bwt_set_error("%s: %s", fm_path, err.c_str());
        return err.find("out of host memory") != std::string::npos ? BWT_ERR_NOMEM : BWT_ERR_FORMAT;
    }
    fm.clear(); fm.shrink_to_fit();
    if (ix.L.size() != ref.size() + 1) {
        bwt_set_error("%s has %zu rows but %s holds %zu bases (expected rows = bases + 1)", fm_path, ix.L.size(),
                      ref_path, ref.size());
        return BWT_ERR_FORMAT;
    }
    return bwt_open_arrays(ix.L.data(), ix.L.size(), ix.SA.data(), ix.pointers, ix.shift, ref.data(), ref.size(), device, out);
}

// host only: the header of a <Chrom>_pJ.fm without touching a GPU (tooling; the CPU tests of the container reader)
extern "C" int bwt_fm_inspect(const char *fm_path, uint64_t *n_rows, uint32_t pointers[4], int64_t *shift) {
    if (!fm_path) { bwt_set_error("bwt_fm_inspect: null argument"); return BWT_ERR_ARG; }
    try {
        std::vector<uint8_t> fm;
        BWT_TRY(read_file(fm_path, fm));
        EtfIndex ix;
        std::string err;
        if (!etf_parse_fm(fm.data(), fm.size(), ix, err)) {
            bwt_set_error("%s: %s", fm_path, err.c_str());
            return err.find("out of host memory") != std::string::npos ? BWT_ERR_NOMEM : BWT_ERR_FORMAT;
        }
        if (n_rows) *n_rows = ix.L.size();
        if (pointers) memcpy(pointers, ix.pointers, sizeof ix.pointers);
        if (shift) *shift = ix.shift;
        return BWT_OK;
    } catch (const std::bad_alloc &) {
        bwt_set_error("bwt_fm_inspect: out of host memory reading %s", fm_path);
        return BWT_ERR_NOMEM;
    }
}

extern "C" int bwt_close(bwt_index *h) {
    if (!h) return BWT_OK;
    cudaSetDevice(h->device);
    aligner_cache_drop(h);
    delete h->tail_mu;
    cudaFree(h->d_occ); cudaFree(h->d_sa); cudaFree(h->d_ktab); cudaFree(h->d_kdeep); cudaFree(h->d_fused); cudaFree(h->d_ref);
    cudaFree(h->d_ref2); cudaFree(h->d_refx);
    delete h;
    return BWT_OK;
}

extern "C" int bwt_info(const bwt_index *h, bwt_index_info *info) {
    if (!h || !info) { bwt_set_error("bwt_info: null argument"); return BWT_ERR_ARG; }
    info->n_rows = h->n;
    memcpy(info->pointers, h->pointers, sizeof info->pointers);
    info->shift = h->shift;
    info->ref_len = h->ref_len;
    info->dollar_row = h->dv.dollar;
    info->device_bytes = h->device_bytes;
    info->device = h->device;
    info->kmer_k = h->dv.deep_k ? h->dv.deep_k : 8;
    info->fused = h->dv.fused ? 1 : 0;
    return BWT_OK;
}

// ---- export / save ----------------------------------------------------------------------------------
namespace {
__global__ void unpack_L_kernel(const uint4 *__restrict__ occ, uint64_t n, uint32_t dollar, uint8_t *__restrict__ L) {
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    uint4 pl = occ[2 * (row >> 6) + 1];
    uint32_t k = row & 63;
    uint32_t lo = ((k < 32 ? pl.x : pl.y) >> (k & 31)) & 1u, hi = ((k < 32 ? pl.z : pl.w) >> (k & 31)) & 1u;
    const char sym[4] = {'A', 'C', 'G', 'T'};
    L[row] = (row + 1 == dollar) ? '$' : (uint8_t)sym[hi * 2 + lo];
}
}  // namespace

extern "C" int bwt_export(const bwt_index *h, uint8_t *L, uint32_t *SA, uint8_t *ref) {
    if (!h) { bwt_set_error("bwt_export: null handle"); return BWT_ERR_ARG; }
    CU_TRY(cudaSetDevice(h->device));
    if (L) {
        uint8_t *d_L = nullptr;
        CU_TRY(cudaMalloc(&d_L, h->n));
        unpack_L_kernel<<<(unsigned)div_up_u64(h->n, 256), 256>>>(h->dv.occ, h->n, h->dv.dollar, d_L);
        cudaError_t e = cudaMemcpy(L, d_L, h->n, cudaMemcpyDeviceToHost);
        cudaFree(d_L);
        CU_TRY(e);
    }
    if (SA) CU_TRY(cudaMemcpy(SA, h->d_sa, h->n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (ref) CU_TRY(cudaMemcpy(ref, (const uint8_t *)h->d_ref + BWT_REF_EXT, h->ref_len, cudaMemcpyDeviceToHost));
    return BWT_OK;
}

extern "C" int bwt_save(const bwt_index *h, const char *fm_path, const char *ref_path) {
    if (!h || !fm_path || !ref_path) { bwt_set_error("bwt_save: null argument"); return BWT_ERR_ARG; }
    std::vector<uint8_t> L(h->n), ref(h->ref_len), out;
    std::vector<uint32_t> SA(h->n);
    BWT_TRY(bwt_export(h, L.data(), SA.data(), ref.data()));
    etf_write_fm(L.data(), SA.data(), h->n, h->pointers, h->shift, out);
    FILE *f = fopen(fm_path, "wb");
    if (!f || fwrite(out.data(), 1, out.size(), f) != out.size()) { if (f) fclose(f); bwt_set_error("cannot write %s", fm_path); return BWT_ERR_IO; }
    fclose(f);
    f = fopen(ref_path, "wb");
    if (!f || fwrite(ref.data(), 1, ref.size(), f) != ref.size()) { if (f) fclose(f); bwt_set_error("cannot write %s", ref_path); return BWT_ERR_IO; }
    fclose(f);
    return BWT_OK;
}
